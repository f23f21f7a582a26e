"""Time every GEMM shape of the WGAN G+D step (bf16) in its AUTO tile configuration, B=64 and the batched critic's B=128;
extra (bn, mt, ksplit) triples can be appended as arguments "bn,mt,ks" to compare against auto.
   python tools/sweep_layers.py [bn,mt,ks ...] > gpurun_out/sweep.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
CONFIGS = [(0, 0, 0)] + [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]


def timeit(fn, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return sum(ts[1:-1]) / (len(ts) - 2) * 1e3      # us


def bf(shape):
    return torch.randn((1,) + shape, device='cuda').to(torch.bfloat16)


def conv_fwd(B, r, cin, cout):
    x = bf((B, r, r, r, cin)); w = torch.randn((cout, cin, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_conv_fwd(w, 1); ro = r // 2; rows = B * ro ** 3
    y = torch.empty((1, rows, cout), dtype=torch.bfloat16, device='cuda')
    fl = 2.0 * rows * cout * 64 * cin
    return lambda bn, mt, ks: raw.igemm(L.MODE_CONV, 1, x, (B, r, r, r, cin), rows, 64 * cin, img, cout, y, cout, act=L.ACT_LRELU, bn=bn, mt=mt, ksplit=ks), fl


def convt_fwd(B, r, cin, cout):          # also == conv dgrad with roles swapped
    x = bf((B, r, r, r, cin)); w = torch.randn((cin, cout, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_convt_fwd(w, 1); ro = 2 * r
    y = torch.empty((1, B * ro ** 3, cout), dtype=torch.bfloat16, device='cuda')
    fl = 2.0 * B * r ** 3 * 8 * cout * 8 * cin
    return lambda bn, mt, ks: raw.igemm(L.MODE_CONVT, 1, x, (B, r, r, r, cin), B * r ** 3, 8 * cin, img, cout, y, cout, out_dims=(ro, ro, ro), bn=bn, mt=mt, ksplit=ks), fl


def patch_fwd(B, r, cout, masked):       # Conv3d(1 -> cout) on the fp32 input volume (D1 forward; masked + in place = the adjoint sweep's first layer)
    x = torch.randn((B, r, r, r), device='cuda'); w = torch.randn((cout, 1, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_b(w, 1, cout, 64, 64, 1, 1, s_n0=64, s_tap=1, s_c=0); ro = r // 2; rows = B * ro ** 3
    y = bf((rows, cout))
    fl = 2.0 * rows * cout * 64
    kw = dict(mask=y, mask_act=L.ACT_LRELU) if masked else dict(act=L.ACT_LRELU)
    return lambda bn, mt, ks: raw.igemm(L.MODE_PATCH, 1, x, (B, r, r, r, 1), rows, 64, img, cout, y, cout, bn=bn, mt=mt, ksplit=ks, **kw), fl


def wgrad_conv(B, r, cin, cout):
    x = bf((B, r, r, r, cin)); ro = r // 2; dy = bf((B, ro, ro, ro, cout)); rows = B * ro ** 3
    g = torch.zeros((cout, cin, 4, 4, 4), device='cuda')
    fl = 2.0 * rows * cout * 64 * cin
    return lambda bn, mt, ks: raw.wgrad(L.MODE_CONV, 1, dy, cout, x, (B, r, r, r, cin), rows, g, sm=cin * 64, st=1, sc=64, m_valid=cout, merge_n=1, ksplit=ks), fl


print('torch', torch.__version__, torch.cuda.get_device_name(0))
for B in (64, 128, 192):
    layers = [('patch conv 1->64 32^3 (D1 fwd)', patch_fwd(B, 32, 64, False)), ('patch conv 1->64 32^3 masked in place (D1 adjoint)', patch_fwd(B, 32, 64, True)),
              ('conv 64->128 16^3 (D2 fwd / G3 dgrad)', conv_fwd(B, 16, 64, 128)), ('conv 128->256 8^3 (D3 fwd / G2 dgrad)', conv_fwd(B, 8, 128, 256)),
              ('convT 256->128 4^3 (G2 fwd / D3 dgrad)', convt_fwd(B, 4, 256, 128)), ('convT 128->64 8^3 (G3 fwd / D2 dgrad)', convt_fwd(B, 8, 128, 64)),
              ('wgrad conv 64->128 16^3 (D2 / G3)', wgrad_conv(B, 16, 64, 128)), ('wgrad conv 128->256 8^3 (D3 / G2)', wgrad_conv(B, 8, 128, 256))]
    for name, (fn, fl) in layers:
        for bn, mt, ks in CONFIGS:
            if 'wgrad' in name and (bn, mt, ks) != (0, 0, 0):
                continue
            try:
                us = timeit(lambda: fn(bn, mt, ks))
                print('B=%3d %-50s %6.2f GFLOP  bn=%3d mt=%d ks=%d  %8.1f us  %7.1f TFLOP/s' % (B, name, fl / 1e9, bn, mt, ks, us, fl / us / 1e6))
            except Exception as e:
                print('B=%3d %-50s bn=%3d mt=%d ks=%d -- %s' % (B, name, bn, mt, ks, str(e)[:80]))
