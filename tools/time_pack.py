"""time the weight re-packs of the 32^3 GAN (fast tap-contiguous kernel vs the generic one)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import raw
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
cases = [('conv_fwd 64->128', raw.pack_conv_fwd, (128, 64, 4, 4, 4)), ('conv_fwd 128->256', raw.pack_conv_fwd, (256, 128, 4, 4, 4)),
         ('conv_dgrad 128->256', raw.pack_conv_dgrad, (256, 128, 4, 4, 4)), ('convt_fwd 256->128', raw.pack_convt_fwd, (256, 128, 4, 4, 4)),
         ('convt_dgrad 256->128', raw.pack_convt_dgrad, (256, 128, 4, 4, 4)), ('convt_fwd 128->64', raw.pack_convt_fwd, (128, 64, 4, 4, 4))]
for name, fn, shape in cases:
    w = torch.randn(shape, device='cuda')
    out = []
    for nf in ('1', '0'):
        os.environ['SG_B200_NO_FAST_PACK'] = nf
        out.append(t(lambda: fn(w, 1)))
    print('%-24s generic %7.1f us   fast %7.1f us' % (name, out[0], out[1]))
