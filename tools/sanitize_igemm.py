"""Small launches of every implicit-GEMM kernel variant for compute-sanitizer (racecheck / synccheck / memcheck):
   compute-sanitizer --tool racecheck python tools/sanitize_igemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw

def bf(shape): return torch.randn((1,) + shape, device='cuda').to(torch.bfloat16)

B = 4
# halo / CTA-pair kernels: Conv3d(64->128) 16^3 -> 8^3 with bias + LeakyReLU, then with a fused mask in place
x = bf((B, 16, 16, 16, 64)); w = torch.randn((128, 64, 4, 4, 4), device='cuda') * 0.05; bias = torch.randn(128, device='cuda')
img = raw.pack_conv_fwd(w, 1); rows = B * 512
y = torch.empty((1, rows, 128), dtype=torch.bfloat16, device='cuda')
raw.igemm(L.MODE_CONV, 1, x, (B, 16, 16, 16, 64), rows, 4096, img, 128, y, 128, bias=bias, act=L.ACT_LRELU)
raw.igemm(L.MODE_CONV, 1, x, (B, 16, 16, 16, 64), rows, 4096, img, 128, y, 128, mask=y, mask_act=L.ACT_LRELU)
os.environ['SG_B200_NO_PAIR'] = '1'
raw.igemm(L.MODE_CONV, 1, x, (B, 16, 16, 16, 64), rows, 4096, img, 128, y, 128, bias=bias, act=L.ACT_LRELU)
os.environ['SG_B200_NO_PAIR'] = '0'
# ConvTranspose3d(128->64) 8^3 -> 16^3 (halo, 8 classes, scattered rows)
xt = bf((B, 8, 8, 8, 128)); wt = torch.randn((128, 64, 4, 4, 4), device='cuda') * 0.05
imgt = raw.pack_convt_fwd(wt, 1)
yt = torch.empty((1, B * 4096, 64), dtype=torch.bfloat16, device='cuda')
raw.igemm(L.MODE_CONVT, 1, xt, (B, 8, 8, 8, 128), B * 512, 1024, imgt, 64, yt, 64, out_dims=(16, 16, 16))
# plain kernel: split-K Conv3d(128->256) 8^3 -> 4^3, dense, Conv3d(1->64) on the fp32 volume
x3 = bf((B, 8, 8, 8, 128)); w3 = torch.randn((256, 128, 4, 4, 4), device='cuda') * 0.05
img3 = raw.pack_conv_fwd(w3, 1)
y3 = torch.empty((1, B * 64, 256), dtype=torch.bfloat16, device='cuda')
raw.igemm(L.MODE_CONV, 1, x3, (B, 8, 8, 8, 128), B * 64, 8192, img3, 256, y3, 256, act=L.ACT_LRELU)
xd = bf((1000, 256)); wd = torch.randn((256, 256), device='cuda') * 0.05
yd = torch.empty((1, 1000, 256), dtype=torch.bfloat16, device='cuda')
raw.igemm(L.MODE_DENSE, 1, xd, (1, 1, 1, 1, 256), 1000, 256, raw.pack_linear(wd, 1), 256, yd, 256, act=L.ACT_RELU)
xp = torch.randn((B, 32, 32, 32), device='cuda'); wp = torch.randn((64, 1, 4, 4, 4), device='cuda') * 0.05
yp = torch.empty((1, B * 4096, 64), dtype=torch.bfloat16, device='cuda')
raw.igemm(L.MODE_PATCH, 1, xp, (B, 32, 32, 32, 1), B * 4096, 64, raw.pack_b(wp, 1, 64, 64, 64, 1, 1, s_n0=64, s_tap=1, s_c=0), 64, yp, 64, act=L.ACT_LRELU)
torch.cuda.synchronize()
print('ok', L.lib().sg_check_device_error())
