"""SDFNet-layer shaped DENSE igemm (rows x 256 x 256, bf16) under different tile configs + wgrad."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw
n = 1 << 20
x = torch.randn((1, n, 256), device='cuda').to(torch.bfloat16)
w = torch.randn((256, 256), device='cuda') * 0.05
b = torch.randn(256, device='cuda')
img = raw.pack_linear(w, 1)
y = torch.empty((1, n, 256), dtype=torch.bfloat16, device='cuda')
fl = 2.0 * n * 256 * 256
def t(fn, reps=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for bn in (0, 64, 128, 256):
    for mt in (0, 1, 2):
        try:
            us = t(lambda: raw.igemm(L.MODE_DENSE, 1, x, (1, 1, 1, 1, 256), n, 256, img, 256, y, 256, bias=b, act=L.ACT_RELU, bn=bn, mt=mt))
            print('dense bn=%3d mt=%d %8.1f us %7.1f TFLOP/s  (HBM floor %.0f us)' % (bn, mt, us, fl / us / 1e6, n * 1024 / 6.4874e6))
        except Exception as e:
            print('dense bn=%3d mt=%d -- %s' % (bn, mt, str(e)[:50]))
g = torch.zeros((256, 256), device='cuda')
for ks in (0, 1, 2):
    us = t(lambda: raw.wgrad(L.MODE_DENSE, 1, y, 256, x, (1, 1, 1, 1, 256), n, g, sm=256, st=0, sc=1, m_valid=256, ksplit=ks))
    print('wgrad ksplit=%d %8.1f us %7.1f TFLOP/s' % (ks, us, fl / us / 1e6))
