"""What does a kernel boundary cost inside a CUDA graph?  (a) N dependent tiny kernels; (b) N conv-probe GEMMs back to back;
(c) the same with a tiny kernel between consecutive GEMMs; (d) with a weight re-pack between them."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw

N = 50
small = torch.zeros(1024, device='cuda')
b, r, cin, cout = 64, 16, 64, 128
x = torch.randn((1, b, r, r, r, cin), device='cuda').to(torch.bfloat16)
w = torch.randn((cout, cin, 4, 4, 4), device='cuda') * 0.05
img = raw.pack_conv_fwd(w, 1)
y = torch.empty((1, b, r // 2, r // 2, r // 2, cout), dtype=torch.bfloat16, device='cuda')
rows = b * (r // 2) ** 3
img2 = torch.empty_like(img)
gemm = lambda: raw.igemm(L.MODE_CONV, 1, x, (b, r, r, r, cin), rows, 64 * cin, img, cout, y, cout, act=L.ACT_LRELU)
tiny = lambda: small.add_(1.0)
pack = lambda: raw.pack_b(w, 1, cout, 64 * cin, 64, cin, cin, s_n0=cin * 64, s_tap=1, s_c=64, out=img2)

def run(body, name):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N): body()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 / N * 1e3
    print('%-44s %8.2f us per iteration' % (name, us))
    return us

a = run(tiny, 'tiny kernel')
bb = run(gemm, 'conv probe GEMM (L2-warm, back to back)')
c = run(lambda: (gemm(), tiny()), 'GEMM + tiny kernel')
d = run(lambda: (gemm(), pack()), 'GEMM + weight re-pack (64->128)')
e = run(lambda: (gemm(), tiny(), tiny(), tiny()), 'GEMM + 3 tiny kernels')
print('marginal cost of a tiny node between GEMMs: %.2f us; of a re-pack: %.2f us; tiny alone %.2f us' % (c - bb, d - bb, a))
