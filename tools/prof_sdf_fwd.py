"""Profiling driver: SDFNet forward only (no_grad), N points, repeated.  Usage: python tools/prof_sdf_fwd.py [N] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from model.sdf_net import SDFNet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = SDFNet()
pts = torch.rand((n, 3), device='cuda') * 2 - 1
table = torch.randn((64, 128), device='cuda') * 0.1
idx = (torch.arange(n, device='cuda') % 64).to(torch.int32)
with torch.no_grad():
    for _ in range(reps):
        out = net(pts, table, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(reps):
        out = net(pts, table, idx)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('sdfnet fwd N=%d: %.3f ms  %.1f Mpts/s  %.1f TFLOP/s' % (n, ms, n / ms / 1e3, 0.921e6 * n / ms / 1e9))
