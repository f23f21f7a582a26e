"""clock64 trace of CTA 0 of the halo implicit-GEMM kernels (SG_B200_IGEMM_DIAG=128): where does a tile's time go?
channels (128 slots each): 0 producer got an A-block slot | 1 weight loader got a stage | 2 weight loader signalled `full` |
3 MMA thread got an A block | 4 MMA thread got a weight stage | 5 MMA thread issued a tap | 6 epilogue (accfull seen, tile stored) | 7 misc"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapegan_b200 import _lib as L, raw

def bf(shape): return torch.randn((1,) + shape, device='cuda').to(torch.bfloat16)

def conv_fwd(B, r, cin, cout):
    x = bf((B, r, r, r, cin)); w = torch.randn((cout, cin, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_conv_fwd(w, 1); ro = r // 2; rows = B * ro ** 3
    y = torch.empty((1, rows, cout), dtype=torch.bfloat16, device='cuda')
    return lambda: raw.igemm(L.MODE_CONV, 1, x, (B, r, r, r, cin), rows, 64 * cin, img, cout, y, cout, act=L.ACT_LRELU)

def patch_fwd(B, r, cout, masked):
    x = torch.randn((B, r, r, r), device='cuda'); w = torch.randn((cout, 1, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_b(w, 1, cout, 64, 64, 1, 1, s_n0=64, s_tap=1, s_c=0); rows = B * (r // 2) ** 3
    y = bf((rows, cout))
    kw = dict(mask=y, mask_act=L.ACT_LRELU) if masked else dict(act=L.ACT_LRELU)
    return lambda: raw.igemm(L.MODE_PATCH, 1, x, (B, r, r, r, 1), rows, 64, img, cout, y, cout, **kw)

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
if 'patch' in sys.argv:
    for masked in (False, True):
        fn = patch_fwd(64, 32, 64, masked)
        os.environ['SG_B200_IGEMM_DIAG'] = '128'
        fn(); torch.cuda.synchronize(); flush.zero_(); fn(); torch.cuda.synchronize()
        os.environ['SG_B200_IGEMM_DIAG'] = '0'
        buf = (ctypes.c_longlong * 1024)()
        L.check(L.lib().sg_debug_igemm_trace(buf, 1024), 'trace')
        ch = [list(buf[c * 128:(c + 1) * 128]) for c in range(8)]
        t0 = min(v for v in ch[1][:1] + ch[4][:1] if v)
        rel = lambda v: v - t0 if v else -1
        print('==== patch conv 1->64 32^3 B=64 %s (plain kernel; clocks since the producer got its first stage)' % ('masked, in place' if masked else ''))
        print(' item | producer got stage | producer signalled | MMA got data | MMA issued | acc ready | stored')
        for i in range(12):
            print(' %3d   %8d %8d %8d %8d %8d %8d' % (i, rel(ch[1][i]), rel(ch[2][i]), rel(ch[4][i]), rel(ch[5][i]), rel(ch[6][2 * i]), rel(ch[6][2 * i + 1])))
    sys.exit(0)
def convt_fwd(B, r, cin, cout):
    x = bf((B, r, r, r, cin)); w = torch.randn((cin, cout, 4, 4, 4), device='cuda') * 0.05
    img = raw.pack_convt_fwd(w, 1); ro = 2 * r
    y = torch.empty((1, B * ro ** 3, cout), dtype=torch.bfloat16, device='cuda')
    return lambda: raw.igemm(L.MODE_CONVT, 1, x, (B, r, r, r, cin), B * r ** 3, 8 * cin, img, cout, y, cout, out_dims=(ro, ro, ro))

if 'convt' in sys.argv:
    for k, v in (x.split('=') for x in sys.argv[1:] if '=' in x): os.environ[k] = v
    fn = convt_fwd(64, 8, 128, 64)
    os.environ['SG_B200_IGEMM_DIAG'] = '128'
    fn(); torch.cuda.synchronize(); flush.zero_(); fn(); torch.cuda.synchronize()
    os.environ['SG_B200_IGEMM_DIAG'] = '0'
    buf = (ctypes.c_longlong * 1024)()
    L.check(L.lib().sg_debug_igemm_trace(buf, 1024), 'trace')
    ch = [list(buf[c * 128:(c + 1) * 128]) for c in range(8)]
    t0 = ch[7][0]
    rel = lambda v: v - t0 if v else -1
    print('==== convT 128->64 8^3 B=64 (CTA pair; 16 taps = 4 halo blocks per work item; clocks since kernel entry of CTA 0)')
    print('setup done %d | kernel exit %d' % (rel(ch[7][1]), rel(ch[7][3])))
    print('epilogue per item: ' + ' '.join('[acc ready %d, stored %d]' % (rel(ch[6][2 * i]), rel(ch[6][2 * i + 1])) for i in range(12) if ch[6][2 * i]))
    print('A blocks (producer got slot -> MMA got block): ' + ' '.join('%d->%d' % (rel(ch[0][i]), rel(ch[3][i])) for i in range(40)))
    print(' tap | loader got stage | loader signalled | MMA got weights | MMA issued | d(issued)')
    for i in range(128):
        print(' %3d   %8d %8d %8d %8d   +%d' % (i, rel(ch[1][i]), rel(ch[2][i]), rel(ch[4][i]), rel(ch[5][i]), ch[5][i] - ch[5][i - 1] if i else 0))
    sys.exit(0)
for B in (64, 128):
    for pair in (0, 1):
        os.environ['SG_B200_NO_PAIR'] = '0' if pair else '1'
        for k, v in (x.split('=') for x in sys.argv[1:] if '=' in x): os.environ[k] = v
        fn = conv_fwd(B, 16, 64, 128)
        os.environ['SG_B200_IGEMM_DIAG'] = '128'
        fn(); torch.cuda.synchronize(); flush.zero_(); fn(); torch.cuda.synchronize()
        os.environ['SG_B200_IGEMM_DIAG'] = '0'
        buf = (ctypes.c_longlong * 1024)()
        L.check(L.lib().sg_debug_igemm_trace(buf, 1024), 'trace')
        ch = [list(buf[c * 128:(c + 1) * 128]) for c in range(8)]
        t0 = ch[7][0]
        rel = lambda v: v - t0 if v else -1
        print('==== conv 64->128 16^3 B=%d %s   (clocks since kernel entry of CTA 0)' % (B, 'CTA pair' if pair else 'single CTA'))
        print('setup done %d | MMA thread has accumulator %d | kernel exit %d' % (rel(ch[7][1]), rel(ch[7][2]), rel(ch[7][3])))
        print('epilogue: ' + ' '.join('[acc ready %d, stored %d]' % (rel(ch[6][2 * i]), rel(ch[6][2 * i + 1])) for i in range(4) if ch[6][2 * i]))
        print('A blocks (producer got slot -> MMA got block): ' + ' '.join('%d->%d' % (rel(ch[0][i]), rel(ch[3][i])) for i in range(20)))
        print(' tap | loader got stage | loader signalled | MMA got weights | MMA issued | d(issued)')
        for i in range(128):
            if i < 16 or i % 8 == 0 or 60 <= i < 72:
                print(' %3d   %8d %8d %8d %8d   +%d' % (i, rel(ch[1][i]), rel(ch[2][i]), rel(ch[4][i]), rel(ch[5][i]), ch[5][i] - ch[5][i - 1] if i else 0))
