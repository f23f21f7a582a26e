"""Profiling driver: the roofline-probe kernel alone (sg_igemm MODE_CONV, Conv3d(64->128,k4,s2,p1) forward, B=64, bf16).
   Usage: python tools/prof_conv.py [reps]      (run under ncu --set full -k regex:sg_igemm)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
b, r, cin, cout = 64, 16, 64, 128
x = torch.randn((1, b, r, r, r, cin), device='cuda').to(torch.bfloat16)
w = torch.randn((cout, cin, 4, 4, 4), device='cuda') * 0.05
img = raw.pack_conv_fwd(w, 1)
y = torch.empty((1, b, r // 2, r // 2, r // 2, cout), dtype=torch.bfloat16, device='cuda')
rows = b * (r // 2) ** 3
for _ in range(reps):
    raw.igemm(L.MODE_CONV, 1, x, (b, r, r, r, cin), rows, 64 * cin, img, cout, y, cout, act=L.ACT_LRELU)
torch.cuda.synchronize()
print('ok')
