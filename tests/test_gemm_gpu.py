"""GPU (-m gpu): unit parity of the tcgen05 implicit-GEMM / wgrad kernels through the C ABI, against torch fp64
math on the SAME bf16-rounded operands (so the only difference is fp32 accumulation order: tolerance 2e-5 relative
to the output's max-abs for fp32 outputs; bf16 outputs add one rounding, 2^-8).  fp32x (planes=2) is compared
with fp64 math on the UNROUNDED fp32 operands: the hi/lo split must recover fp32 accuracy (1e-5)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL_F32 = 2e-5
TOL_F32X = 1e-4     # hi/lo bf16 split drops the lo*lo term (2^-18) and the tensor core truncates long fp32 sums
TOL_BF16 = 6e-3


def _imports():
    from shapegan_b200 import _lib as L
    from shapegan_b200 import raw
    return L, raw


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g) * 2 - 1) * scale).cuda()


def q(x, planes):
    """operand as the kernel sees it"""
    return x.to(torch.bfloat16).double() if planes == 1 else x.double()


def report(name, got, ref, tol):
    got, ref = got.double(), ref.double()
    scale = ref.abs().max().item() + 1e-30
    err = (got - ref).abs()
    m = err.max().item() / scale
    if m > tol:
        bad = (err / scale > tol)
        idx = bad.nonzero()
        print('\n[%s] FAIL rel-max-err %.3e tol %.1e; %d/%d bad; first bad %s; got %s ref %s' % (
            name, m, tol, bad.sum().item(), bad.numel(), idx[0].tolist(), got[tuple(idx[0])].item(), ref[tuple(idx[0])].item()))
        # row / column structure of the failure helps localise layout bugs
        if got.dim() == 2:
            print('   bad rows:', bad.any(1).nonzero().flatten()[:16].tolist(), 'bad cols:', bad.any(0).nonzero().flatten()[:16].tolist())
    assert m <= tol, '%s: rel-max-err %.3e > %.1e' % (name, m, tol)
    return m


def check_error_word():
    L, raw = _imports()
    torch.cuda.synchronize()
    assert L.lib().sg_check_device_error() == 0


@pytest.mark.parametrize('planes', [1, 2])
@pytest.mark.parametrize('rows,k,n,bn,mt', [(300, 192, 256, 0, 0), (128, 64, 16, 0, 0), (1000, 256, 256, 256, 2),
                                            (77, 128, 48, 0, 1), (4096, 512, 128, 64, 2), (5, 16384, 128, 0, 0)])
def test_dense(planes, rows, k, n, bn, mt):
    L, raw = _imports()
    x = rnd((rows, k), 1)
    w = rnd((n, k), 2, 0.1)
    bias = rnd((n,), 3)
    xp = raw.to_planes(x, planes)
    img = raw.pack_linear(w, planes)
    out = torch.empty((rows, n), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_DENSE, planes, xp, (1, 1, 1, 1, k), rows, k, img, n, out, n, out_kind=L.OUT_F32, bias=bias, bn=bn, mt=mt)
    ref = q(x, planes) @ q(w, planes).t() + bias.double()
    report('dense p%d %s' % (planes, (rows, k, n)), out, ref, TOL_F32 if planes == 1 else TOL_F32X)
    # bf16 plane output + activation
    outp = torch.empty((planes, rows, n), dtype=torch.bfloat16, device='cuda')
    raw.igemm(L.MODE_DENSE, planes, xp, (1, 1, 1, 1, k), rows, k, img, n, outp, n, bias=bias, act=L.ACT_LRELU, bn=bn, mt=mt)
    report('dense-bf16 p%d' % planes, raw.from_planes(outp), F.leaky_relu(ref, 0.2), TOL_BF16 if planes == 1 else 1e-4)
    check_error_word()


def test_dense_two_sources_and_splitk():
    L, raw = _imports()
    rows = 500
    x1, x2 = rnd((rows, 256), 1), rnd((rows, 192), 2)
    w = rnd((256, 448), 3, 0.1)
    for planes in (1, 2):
        img = raw.pack_linear(w, planes)
        out = torch.empty((rows, 256), dtype=torch.float32, device='cuda')
        raw.igemm(L.MODE_DENSE, planes, raw.to_planes(x1, planes), (1, 1, 1, 1, 256), rows, 448, img, 256, out, 256,
                  out_kind=L.OUT_F32, a2=raw.to_planes(x2, planes), a2_c=192)
        ref = torch.cat((q(x1, planes), q(x2, planes)), 1) @ q(w, planes).t()
        report('dense-2src p%d' % planes, out, ref, TOL_F32)
    # split-K with atomics
    x = rnd((20, 16384), 4)
    w = rnd((128, 16384), 5, 0.05)
    bias = rnd((128,), 6)
    out = torch.zeros((20, 128), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_DENSE, 1, raw.to_planes(x, 1), (1, 1, 1, 1, 16384), 20, 16384, raw.pack_linear(w, 1), 128, out, 128,
              out_kind=L.OUT_F32_ATOMIC, bias=bias, ksplit=64)
    report('dense-splitk', out, q(x, 1) @ q(w, 1).t() + bias.double(), TOL_F32)
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
@pytest.mark.parametrize('b,r,cin,cout', [(2, 8, 64, 128), (3, 16, 32, 64), (2, 8, 24, 48), (1, 4, 128, 256), (2, 8, 96, 16)])
def test_conv_fwd(planes, b, r, cin, cout):
    L, raw = _imports()
    x = rnd((b, r, r, r, cin), 1)                       # NDHWC
    w = rnd((cout, cin, 4, 4, 4), 2, 0.05)
    bias = rnd((cout,), 3)
    ro = r // 2
    rows = b * ro ** 3
    out = torch.empty((rows, cout), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_CONV, planes, raw.to_planes(x, planes), (b, r, r, r, cin), rows, 64 * cin, raw.pack_conv_fwd(w, planes),
              cout, out, cout, out_kind=L.OUT_F32, bias=bias)
    ref = F.conv3d(q(x, planes).permute(0, 4, 1, 2, 3), q(w, planes), bias.double(), stride=2, padding=1)
    report('conv p%d %s' % (planes, (b, r, cin, cout)), out, ref.permute(0, 2, 3, 4, 1).reshape(rows, cout), TOL_F32)
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
@pytest.mark.parametrize('b,r,cin,cout', [(2, 4, 64, 32), (3, 8, 128, 64), (2, 4, 256, 128), (1, 8, 48, 24)])
def test_convt_fwd(planes, b, r, cin, cout):
    L, raw = _imports()
    x = rnd((b, r, r, r, cin), 1)
    w = rnd((cin, cout, 4, 4, 4), 2, 0.05)              # ConvTranspose3d layout
    bias = rnd((cout,), 3)
    ro = 2 * r
    out = torch.zeros((b * ro ** 3, cout), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_CONVT, planes, raw.to_planes(x, planes), (b, r, r, r, cin), b * r ** 3, 8 * cin,
              raw.pack_convt_fwd(w, planes), cout, out, cout, out_kind=L.OUT_F32, bias=bias, out_dims=(ro, ro, ro))
    ref = F.conv_transpose3d(q(x, planes).permute(0, 4, 1, 2, 3), q(w, planes), bias.double(), stride=2, padding=1)
    report('convt p%d %s' % (planes, (b, r, cin, cout)), out, ref.permute(0, 2, 3, 4, 1).reshape(-1, cout), TOL_F32)
    check_error_word()


@pytest.mark.parametrize('mt', [1, 2])
@pytest.mark.parametrize('mode,b,cin,cout', [('conv', 2, 64, 128), ('conv', 3, 128, 64), ('convt', 2, 64, 64), ('convt', 3, 128, 32)])
def test_halo_variant(mode, b, cin, cout, mt, monkeypatch):
    """8 x 8 x 8 row grids, bf16, C % 64 == 0: the halo-reuse kernel (one strided TMA block serves the four taps that differ by +1
    in x / z; shifted, non-1024-aligned UMMA operand views) against torch, and against the plain one-tile-per-tap kernel."""
    L, raw = _imports()
    bias = rnd((cout,), 3)
    outs = []
    for no_halo, no_pair in (('0', '0'), ('0', '1'), ('1', '1')):       # CTA-pair halo kernel | single-CTA halo kernel | plain kernel
        monkeypatch.setenv('SG_B200_NO_HALO', no_halo)
        monkeypatch.setenv('SG_B200_NO_PAIR', no_pair)
        if mode == 'conv':
            r = 16
            x = rnd((b, r, r, r, cin), 1)
            w = rnd((cout, cin, 4, 4, 4), 2, 0.05)
            rows = b * 8 ** 3
            out = torch.empty((rows, cout), dtype=torch.float32, device='cuda')
            raw.igemm(L.MODE_CONV, 1, raw.to_planes(x, 1), (b, r, r, r, cin), rows, 64 * cin, raw.pack_conv_fwd(w, 1), cout, out, cout,
                      out_kind=L.OUT_F32, bias=bias, mt=mt)
            ref = F.conv3d(q(x, 1).permute(0, 4, 1, 2, 3), q(w, 1), bias.double(), stride=2, padding=1).permute(0, 2, 3, 4, 1).reshape(rows, cout)
        else:
            r = 8
            x = rnd((b, r, r, r, cin), 1)
            w = rnd((cin, cout, 4, 4, 4), 2, 0.05)
            out = torch.zeros((b * 16 ** 3, cout), dtype=torch.float32, device='cuda')
            raw.igemm(L.MODE_CONVT, 1, raw.to_planes(x, 1), (b, r, r, r, cin), b * r ** 3, 8 * cin, raw.pack_convt_fwd(w, 1), cout, out, cout,
                      out_kind=L.OUT_F32, bias=bias, out_dims=(16, 16, 16), mt=mt)
            ref = F.conv_transpose3d(q(x, 1).permute(0, 4, 1, 2, 3), q(w, 1), bias.double(), stride=2, padding=1)
            ref = ref.permute(0, 2, 3, 4, 1).reshape(-1, cout)
        report('halo=%s pair=%s %s mt%d %s' % ('off' if no_halo == '1' else 'on', 'off' if no_pair == '1' else 'on', mode, mt, (b, cin, cout)),
               out, ref, TOL_F32)
        outs.append(out.clone())
    for o in outs[:2]:
        assert (o - outs[2]).abs().max().item() < 1e-3 * max(1.0, outs[2].abs().max().item())
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
def test_conv_dgrad_and_convt_dgrad(planes):
    L, raw = _imports()
    # Conv3d input gradient == MODE_CONVT over dY with the conv weight
    b, r, cin, cout = 2, 8, 32, 64
    w = rnd((cout, cin, 4, 4, 4), 1, 0.05)
    dy = rnd((b, r // 2, r // 2, r // 2, cout), 2)
    out = torch.zeros((b * r ** 3, cin), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_CONVT, planes, raw.to_planes(dy, planes), (b, r // 2, r // 2, r // 2, cout), b * (r // 2) ** 3, 8 * cout,
              raw.pack_conv_dgrad(w, planes), cin, out, cin, out_kind=L.OUT_F32, out_dims=(r, r, r))
    ref = F.conv_transpose3d(q(dy, planes).permute(0, 4, 1, 2, 3), q(w, planes), None, stride=2, padding=1)
    report('conv-dgrad p%d' % planes, out, ref.permute(0, 2, 3, 4, 1).reshape(-1, cin), TOL_F32)
    # ConvTranspose3d input gradient == MODE_CONV over dY with the transposed-conv weight
    wt = rnd((cin, cout, 4, 4, 4), 3, 0.05)
    dy2 = rnd((b, r, r, r, cout), 4)
    rows = b * (r // 2) ** 3
    out2 = torch.zeros((rows, cin), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_CONV, planes, raw.to_planes(dy2, planes), (b, r, r, r, cout), rows, 64 * cout,
              raw.pack_convt_dgrad(wt, planes), cin, out2, cin, out_kind=L.OUT_F32)
    ref2 = F.conv3d(q(dy2, planes).permute(0, 4, 1, 2, 3), q(wt, planes), None, stride=2, padding=1)
    report('convt-dgrad p%d' % planes, out2, ref2.permute(0, 2, 3, 4, 1).reshape(rows, cin), TOL_F32 if planes == 1 else TOL_F32X)
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
def test_patch_conv(planes):
    L, raw = _imports()
    b, r, cout = 3, 16, 64
    vol = rnd((b, r, r, r), 1)
    w = rnd((cout, 1, 4, 4, 4), 2, 0.2)
    bias = rnd((cout,), 3)
    rows = b * (r // 2) ** 3
    out = torch.empty((rows, cout), dtype=torch.float32, device='cuda')
    raw.igemm(L.MODE_PATCH, planes, vol, (b, r, r, r, 1), rows, 64, raw.pack_conv_fwd(w, planes), cout, out, cout,
              out_kind=L.OUT_F32, bias=bias)
    ref = F.conv3d(q(vol, planes).unsqueeze(1), q(w, planes), bias.double(), stride=2, padding=1)
    report('patch p%d' % planes, out, ref.permute(0, 2, 3, 4, 1).reshape(rows, cout), TOL_F32)
    check_error_word()


@pytest.mark.parametrize('merge_n', [0, 1])
@pytest.mark.parametrize('planes', [1, 2])
def test_wgrad(planes, merge_n):
    L, raw = _imports()
    # dense: dW[out,in] = dY^T X
    rows, cin, cout = 1000, 192, 256
    x, dy = rnd((rows, cin), 1), rnd((rows, cout), 2)
    g = torch.zeros((cout, cin), dtype=torch.float32, device='cuda')
    raw.wgrad(L.MODE_DENSE, planes, raw.to_planes(dy, planes), cout, raw.to_planes(x, planes), (1, 1, 1, 1, cin), rows, g,
              sm=cin, st=0, sc=1, m_valid=cout, merge_n=merge_n)
    report('wgrad-dense p%d m%d' % (planes, merge_n), g, q(dy, planes).t() @ q(x, planes), TOL_F32)
    # conv: dW[cout,cin,4,4,4]
    for (b, r, cin, cout) in ((2, 8, 64, 128), (2, 8, 24, 48), (1, 8, 32, 200)):
        x = rnd((b, r, r, r, cin), 3)
        dy = rnd((b, r // 2, r // 2, r // 2, cout), 4)
        rows = b * (r // 2) ** 3
        g = torch.zeros((cout, cin, 4, 4, 4), dtype=torch.float32, device='cuda')
        raw.wgrad(L.MODE_CONV, planes, raw.to_planes(dy, planes), cout, raw.to_planes(x, planes), (b, r, r, r, cin), rows, g,
                  sm=cin * 64, st=1, sc=64, m_valid=cout, merge_n=merge_n)
        xx = q(x, planes).permute(0, 4, 1, 2, 3).requires_grad_(False)
        wref = torch.zeros((cout, cin, 4, 4, 4), dtype=torch.float64, device='cuda', requires_grad=True)
        F.conv3d(xx, wref, None, stride=2, padding=1).backward(q(dy, planes).permute(0, 4, 1, 2, 3))
        report('wgrad-conv p%d m%d %s' % (planes, merge_n, (b, r, cin, cout)), g, wref.grad, TOL_F32)
    # patch: dW[cout,1,4,4,4]
    b, r, cout = 2, 16, 64
    vol = rnd((b, r, r, r), 5)
    dy = rnd((b, r // 2, r // 2, r // 2, cout), 6)
    rows = b * (r // 2) ** 3
    g = torch.zeros((cout, 1, 4, 4, 4), dtype=torch.float32, device='cuda')
    raw.wgrad(L.MODE_PATCH, planes, raw.to_planes(dy, planes), cout, vol, (b, r, r, r, 1), rows, g, sm=64, st=0, sc=1,
              m_valid=cout, merge_n=merge_n)
    wref = torch.zeros((cout, 1, 4, 4, 4), dtype=torch.float64, device='cuda', requires_grad=True)
    F.conv3d(q(vol, planes).unsqueeze(1), wref, None, stride=2, padding=1).backward(q(dy, planes).permute(0, 4, 1, 2, 3))
    report('wgrad-patch p%d m%d' % (planes, merge_n), g, wref.grad, TOL_F32)
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
@pytest.mark.parametrize('mode,b,r,cin,cout,ks', [('conv', 8, 8, 128, 256, 0), ('conv', 5, 8, 128, 256, 3), ('convt', 3, 4, 256, 128, 2),
                                                   ('conv', 2, 8, 64, 40, 4),
                                                   ('conv', 192, 8, 128, 256, 0)])     # the batched critic's D3: auto = two sub-tiles x 3 splits
def test_splitk_partial_slabs(mode, b, r, cin, cout, ks, planes, monkeypatch):
    """Split-K through fp32 partial slabs + the finish kernel (bias, LeakyReLU, mask, bf16 hi/lo planes) against the same layer with
    SG_B200_NO_SPLITK=1: the shape class of Conv3d(128->256) 8^3 -> 4^3 (few output tiles x K = 8192).  ks = 0 is the auto policy."""
    L, raw = _imports()
    x = raw.to_planes(rnd((b, r, r, r, cin), 1), planes)
    bias = rnd((cout,), 3)
    if mode == 'conv':
        w = rnd((cout, cin, 4, 4, 4), 2, 0.05)
        img, gmode, k, rows, ro = raw.pack_conv_fwd(w, planes), L.MODE_CONV, 64 * cin, b * (r // 2) ** 3, r // 2
        out_rows, od = rows, (0, 0, 0)
    else:
        w = rnd((cin, cout, 4, 4, 4), 2, 0.05)
        img, gmode, k, rows, ro = raw.pack_convt_fwd(w, planes), L.MODE_CONVT, 8 * cin, b * r ** 3, 2 * r
        out_rows, od = b * ro ** 3, (ro, ro, ro)
    mask = raw.to_planes(rnd((out_rows, cout), 4), planes)
    outs = []
    for split in (True, False):
        if not split:
            monkeypatch.setenv('SG_B200_NO_SPLITK', '1')
        y = torch.zeros((planes, out_rows, cout), dtype=torch.bfloat16, device='cuda')
        launches0 = L.lib().sg_launch_count()
        raw.igemm(gmode, planes, x, (b, r, r, r, cin), rows, k, img, cout, y, cout, bias=bias, act=L.ACT_LRELU, mask=mask,
                  mask_act=L.ACT_LRELU, out_dims=od, ksplit=ks if split else 0)
        outs.append((raw.from_planes(y), L.lib().sg_launch_count() - launches0))
    monkeypatch.delenv('SG_B200_NO_SPLITK')
    assert outs[0][1] == 2 and outs[1][1] == 1, 'expected main + finish kernel vs a single launch, got %s' % ([o[1] for o in outs],)      # SG_B200_NO_SPLITK also switches the fp32x K chunking off
    report('splitk %s p%d ks%d' % (mode, planes, ks), outs[0][0], outs[1][0], 4e-3 if planes == 1 else 5e-5)
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
def test_wgrad_bias_sums_ride_the_gemm(planes):
    """bias_grad = column sums of the A operand (dY) out of the weight-gradient GEMM itself (one N = 16 MMA per K step against a tile of
    ones): dense and conv gathers, several N groups, ragged rows, M not a multiple of 128, accumulation into an existing gradient; the
    weight gradient of the same call must be unchanged."""
    L, raw = _imports()
    for rows, cin, cout in ((1000, 192, 256), (77, 64, 200), (5000, 256, 256)):
        x, dy = rnd((rows, cin), 1), rnd((rows, cout), 2)
        g = torch.zeros((cout, cin), dtype=torch.float32, device='cuda')
        gb = torch.full((cout,), 0.5, dtype=torch.float32, device='cuda')
        raw.wgrad(L.MODE_DENSE, planes, raw.to_planes(dy, planes), cout, raw.to_planes(x, planes), (1, 1, 1, 1, cin), rows, g,
                  sm=cin, st=0, sc=1, m_valid=cout, bias_grad=gb, bias_accumulate=True)
        report('wgrad-bias dense W p%d %s' % (planes, (rows, cin, cout)), g, q(dy, planes).t() @ q(x, planes), TOL_F32)
        report('wgrad-bias dense b p%d %s' % (planes, (rows, cin, cout)), gb, q(dy, planes).sum(0) + 0.5, TOL_F32)
    b, r, cin, cout = 2, 8, 64, 128
    x = rnd((b, r, r, r, cin), 3)
    dy = rnd((b, r // 2, r // 2, r // 2, cout), 4)
    rows = b * (r // 2) ** 3
    g = torch.zeros((cout, cin, 4, 4, 4), dtype=torch.float32, device='cuda')
    gb = torch.zeros((cout,), dtype=torch.float32, device='cuda')
    raw.wgrad(L.MODE_CONV, planes, raw.to_planes(dy, planes), cout, raw.to_planes(x, planes), (b, r, r, r, cin), rows, g,
              sm=cin * 64, st=1, sc=64, m_valid=cout, bias_grad=gb)
    wref = torch.zeros((cout, cin, 4, 4, 4), dtype=torch.float64, device='cuda', requires_grad=True)
    F.conv3d(q(x, planes).permute(0, 4, 1, 2, 3), wref, None, stride=2, padding=1).backward(q(dy, planes).permute(0, 4, 1, 2, 3))
    report('wgrad-bias conv W p%d' % planes, g, wref.grad, TOL_F32)
    report('wgrad-bias conv b p%d' % planes, gb, q(dy, planes).reshape(-1, cout).sum(0), TOL_F32)
    check_error_word()


@pytest.mark.parametrize('planes', [1, 2])
def test_wgrad_cta_pair_multicast(planes, monkeypatch):
    """M = 256 weight gradients on CTA pairs (the two M tiles of a pair share every gathered B tile through TMA multicast, the stage-free
    commit goes to both CTAs) == the single-CTA kernel, dense (SDFNet layer shape, with the bias-sum MMA) and conv (128 -> 256) gathers,
    with more than one work item per pair and a ragged row count."""
    L, raw = _imports()
    res = []
    monkeypatch.setenv('SG_B200_WGRAD_PAIR_MIN_ROWS', '0')       # the launcher keeps pairs for long contractions only
    for no_pair in ('0', '1'):
        monkeypatch.setenv('SG_B200_NO_WGRAD_PAIR', no_pair)
        out = []
        rows, cin, cout = 128 * 300 + 77, 256, 256
        x, dy = rnd((rows, cin), 1), rnd((rows, cout), 2)
        g = torch.zeros((cout, cin), dtype=torch.float32, device='cuda')
        gb = torch.zeros((cout,), dtype=torch.float32, device='cuda')
        raw.wgrad(L.MODE_DENSE, planes, raw.to_planes(dy, planes), cout, raw.to_planes(x, planes), (1, 1, 1, 1, cin), rows, g,
                  sm=cin, st=0, sc=1, m_valid=cout, bias_grad=gb)
        if no_pair == '1':
            report('wgrad-pair ref dense p%d' % planes, g, q(dy, planes).t() @ q(x, planes), TOL_F32)
        out += [g, gb]
        b, r, cin, cout = 3, 8, 128, 256
        x = rnd((b, r, r, r, cin), 3)
        dy = rnd((b, r // 2, r // 2, r // 2, cout), 4)
        g2 = torch.zeros((cout, cin, 4, 4, 4), dtype=torch.float32, device='cuda')
        raw.wgrad(L.MODE_CONV, planes, raw.to_planes(dy, planes), cout, raw.to_planes(x, planes), (b, r, r, r, cin), b * (r // 2) ** 3, g2,
                  sm=cin * 64, st=1, sc=64, m_valid=cout)
        out.append(g2)
        res.append(out)
    monkeypatch.delenv('SG_B200_NO_WGRAD_PAIR')
    for a, c in zip(res[0], res[1]):
        assert torch.equal(a, c), 'the pair kernel must reproduce the single-CTA partial sums bit for bit (same MMAs, same order)'
    check_error_word()


@pytest.mark.gpu
@pytest.mark.parametrize('planes', [1, 2])
def test_fast_weight_pack_is_bit_identical(planes, monkeypatch):
    """sg_pack_b's shared-memory-transposing kernel for the conv layouts (64 contiguous taps per (n, c) pair) writes exactly the bytes
    of the generic strided kernel: Conv3d fwd / dgrad, ConvTranspose3d fwd / dgrad images, ragged N (cout = 72 -> n_pad = 80)."""
    L, raw = _imports()
    packers = [('conv_fwd', raw.pack_conv_fwd, (72, 128, 4, 4, 4)), ('conv_dgrad', raw.pack_conv_dgrad, (128, 64, 4, 4, 4)),
               ('convt_fwd', raw.pack_convt_fwd, (128, 72, 4, 4, 4)), ('convt_dgrad', raw.pack_convt_dgrad, (64, 128, 4, 4, 4))]
    for name, fn, shape in packers:
        w = rnd(shape, 11)
        monkeypatch.setenv('SG_B200_NO_FAST_PACK', '1')
        ref = fn(w, planes).clone()
        monkeypatch.setenv('SG_B200_NO_FAST_PACK', '0')
        got = fn(w, planes)
        assert ref.numel() == got.numel() and torch.equal(ref, got), name
    monkeypatch.delenv('SG_B200_NO_FAST_PACK')
    check_error_word()


@pytest.mark.gpu
def test_tiled_col2im_is_bit_identical(monkeypatch):
    """second stage of ConvTranspose3d(C -> 1): the shared-memory tiled kernel sums the same 8 taps in the same order as the direct one
    (16^3 -> 32^3 with bias + tanh = the generator's last layer; 8 x 4 x 12 grid without bias = an input-gradient shape)"""
    L, raw = _imports()
    for (n, d, h, w, act, with_bias) in ((3, 16, 16, 16, L.ACT_TANH, True), (2, 12, 4, 8, L.ACT_NONE, False)):
        pm = rnd((n * d * h * w, 64), 5).to(torch.bfloat16).reshape(1, -1, 64)
        bias = rnd((1,), 6) if with_bias else None
        monkeypatch.setenv('SG_B200_NO_TILED_COL2IM', '1')
        ref = raw.col2im_c1(pm, n, d, h, w, bias, act).clone()
        monkeypatch.setenv('SG_B200_NO_TILED_COL2IM', '0')
        got = raw.col2im_c1(pm, n, d, h, w, bias, act)
        assert torch.equal(ref, got)
    monkeypatch.delenv('SG_B200_NO_TILED_COL2IM')
    check_error_word()
