"""Profiling driver: the first discriminator layer alone (sg_igemm MODE_PATCH, Conv3d(1->64,k4,s2,p1) on the fp32 volume, B=64, bf16 out).
   Usage: python tools/prof_patch.py [masked] [reps]      (run under ncu --set full -k regex:sg_igemm_kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw
masked = 'masked' in sys.argv
reps = 5
B, r, cout = 64, 32, 64
x = torch.randn((B, r, r, r), device='cuda'); w = torch.randn((cout, 1, 4, 4, 4), device='cuda') * 0.05
img = raw.pack_b(w, 1, cout, 64, 64, 1, 1, s_n0=64, s_tap=1, s_c=0); rows = B * (r // 2) ** 3
y = torch.randn((1, rows, cout), device='cuda').to(torch.bfloat16)
kw = dict(mask=y, mask_act=L.ACT_LRELU) if masked else dict(act=L.ACT_LRELU)
for _ in range(reps):
    raw.igemm(L.MODE_PATCH, 1, x, (B, r, r, r, 1), rows, 64, img, cout, y, cout, **kw)
torch.cuda.synchronize()
print('ok')
