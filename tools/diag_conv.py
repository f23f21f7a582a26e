"""Which part of sg_igemm bounds a layer?  Times the kernel with parts switched off (SG_B200_IGEMM_DIAG bit mask:
1 no A gather, 2 no B load, 4 no MMA, 8 no epilogue stores; any bit forces the plain kernel).  Results are garbage in those modes: timing only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw
B = 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

def timeit(fn, reps=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return sum(ts[1:-1]) / (len(ts) - 2) * 1e3

def bf(shape): return torch.randn((1,) + shape, device='cuda').to(torch.bfloat16)

def conv_fwd(r, cin, cout):
    x = bf((B, r, r, r, cin)); w = torch.randn((cout, cin, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_conv_fwd(w, 1); ro = r // 2; rows = B * ro ** 3
    y = torch.empty((1, rows, cout), dtype=torch.bfloat16, device='cuda')
    return (lambda: raw.igemm(L.MODE_CONV, 1, x, (B, r, r, r, cin), rows, 64 * cin, img, cout, y, cout, act=L.ACT_LRELU)), 2.0 * rows * cout * 64 * cin

def convt_fwd(r, cin, cout):
    x = bf((B, r, r, r, cin)); w = torch.randn((cin, cout, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_convt_fwd(w, 1); ro = 2 * r
    y = torch.empty((1, B * ro ** 3, cout), dtype=torch.bfloat16, device='cuda')
    return (lambda: raw.igemm(L.MODE_CONVT, 1, x, (B, r, r, r, cin), B * r ** 3, 8 * cin, img, cout, y, cout, out_dims=(ro, ro, ro))), 2.0 * B * r ** 3 * 8 * cout * 8 * cin

def dense(rows, k, n):
    x = bf((rows, k)); w = torch.randn((n, k), device='cuda') * 0.05
    img = raw.pack_linear(w, 1)
    y = torch.empty((1, rows, n), dtype=torch.bfloat16, device='cuda')
    return (lambda: raw.igemm(L.MODE_DENSE, 1, x, (1, 1, 1, 1, k), rows, k, img, n, y, n, act=L.ACT_RELU)), 2.0 * rows * k * n

layers = [('conv 64->128 16^3', conv_fwd(16, 64, 128)), ('conv 128->256 8^3', conv_fwd(8, 128, 256)),
          ('convT 128->64 8^3', convt_fwd(8, 128, 64)), ('dense 32768x4096x128', dense(262144 // 8, 4096, 128))]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush.zero_(); e0.record(); e1.record(); torch.cuda.synchronize(); print('back-to-back events: %.1f us' % (e0.elapsed_time(e1) * 1e3))
small = torch.zeros(1024, device='cuda')
print('tiny torch kernel after flush: %.1f us' % timeit(lambda: small.add_(1.0)))
for name, (fn, fl) in layers:
    print('== %s  %.2f GFLOP' % (name, fl / 1e9))
    for d, what in ((0, 'full'), (1, 'no A gather'), (2, 'no B load'), (3, 'no loads'), (4, 'no MMA'), (8, 'no stores'), (7, 'empty loop'), (15, 'nothing'), (16, 'prologue only')):
        os.environ['SG_B200_IGEMM_DIAG'] = str(d)
        us = timeit(fn)
        print('   diag=%2d %-12s %8.1f us  %7.1f TFLOP/s-equivalent' % (d, what, us, fl / us / 1e6))
os.environ['SG_B200_IGEMM_DIAG'] = '0'; os.environ['SG_B200_IGEMM_GRID'] = '0'

# (clock traces of CTA 0: tools/trace_igemm.py on a `python -m shapegan_b200.build --force --trace` build)
