"""GPU (-m gpu): every LinOp configuration the drop-in modules instantiate (exact channel counts), checked primitive by
primitive (fwd / tr / wgrad) against torch fp64 autograd of the equivalent torch.nn.functional op on the same GPU.
This is the localisation layer between tests/test_gemm_gpu.py (raw kernels) and tests/test_parity_gpu.py (modules).
Runs in fp32x so a layout bug (O(1) error) is clearly separated from rounding (<= 2e-4 of the output scale)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-4


@pytest.fixture(autouse=True)
def _fp32x():
    from shapegan_b200 import config
    old = config.precision()
    config.set_precision('fp32x')
    yield
    config.set_precision(old)


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand(shape, generator=g) * 2 - 1) * scale).cuda()


def close(name, got, ref, tol=TOL):
    got, ref = got.double(), ref.double()
    scale = ref.abs().max().item() + 1e-30
    err = ((got - ref).abs().max().item()) / scale
    if err > tol:
        bad = ((got - ref).abs() / scale > tol)
        print('\n[%s] rel-max-err %.3e; %d/%d bad; got-norm %.4e ref-norm %.4e; first bad idx %s' % (
            name, err, int(bad.sum()), bad.numel(), got.norm().item(), ref.norm().item(), bad.nonzero()[0].tolist()))
    assert err <= tol, '%s: %.3e' % (name, err)


def P(x):
    from shapegan_b200 import raw
    return raw.to_planes(x, 2)


def V(t):
    from shapegan_b200 import raw
    return raw.from_planes(t)


def ndhwc(x):            # [B,C,D,H,W] -> [B,D,H,W,C]
    return x.permute(0, 2, 3, 4, 1).contiguous()


def ncdhw(x):
    return x.permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize('cin,cout,r,b', [(64, 128, 8, 2), (128, 256, 8, 2), (32, 64, 8, 2), (24, 48, 8, 3), (48, 96, 8, 3)])
def test_conv_op(cin, cout, r, b):
    from shapegan_b200 import ops
    op = ops.ConvOp(cin, cout)
    x = rnd((b, cin, r, r, r), 1)
    w = rnd((cout, cin, 4, 4, 4), 2, 0.05)
    bias = rnd((cout,), 3)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, bias.double(), stride=2, padding=1)
    gy = rnd(ref.shape, 4)
    ref.backward(gy.double())
    y = op.fwd(P(ndhwc(x)), w, bias, ops.ACT_NONE)
    close('conv fwd', ncdhw(V(y)), ref)
    close('conv tr', ncdhw(V(op.tr(P(ndhwc(gy)), w))), xd.grad)
    close('conv wgrad', op.wgrad(P(ndhwc(x)), P(ndhwc(gy)), tuple(w.shape)), wd.grad)


@pytest.mark.parametrize('cout,w_cin,r,b', [(64, 1, 16, 2), (32, 1, 16, 2), (24, 1, 32, 3), (64, 32, 16, 2), (128, 64, 8, 2), (256, 128, 8, 2)])
def test_conv1_op(cout, w_cin, r, b):
    from shapegan_b200 import ops
    op = ops.Conv1Op(cout, w_cin=w_cin)
    vol = rnd((b, r, r, r), 1)
    w = rnd((cout, w_cin, 4, 4, 4), 2, 0.2)
    bias = rnd((cout,), 3)
    x_full = torch.zeros((b, w_cin, r, r, r), device='cuda', dtype=torch.float64)
    x_full[:, 0] = vol.double()
    x_full.requires_grad_(True)
    wd = w.double().requires_grad_(True)
    ref = F.conv3d(x_full, wd, bias.double(), stride=2, padding=1)
    gy = rnd(ref.shape, 4)
    ref.backward(gy.double())
    close('conv1 fwd', ncdhw(V(op.fwd(vol, w, bias, ops.ACT_NONE))), ref)
    close('conv1 tr', op.tr(P(ndhwc(gy)), w), x_full.grad[:, 0])
    gw = op.wgrad(vol, P(ndhwc(gy)), tuple(w.shape))
    close('conv1 wgrad', gw, wd.grad)       # channels >= 1 of the reference gradient are exactly zero (zero inputs)


@pytest.mark.parametrize('cin,cout,r,b', [(256, 128, 4, 2), (128, 64, 8, 2), (96, 48, 4, 4), (48, 24, 8, 4)])
def test_convt_op(cin, cout, r, b):
    from shapegan_b200 import ops
    op = ops.ConvTOp(cin, cout)
    x = rnd((b, cin, r, r, r), 1)
    w = rnd((cin, cout, 4, 4, 4), 2, 0.05)
    bias = rnd((cout,), 3)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv_transpose3d(xd, wd, bias.double(), stride=2, padding=1)
    gy = rnd(ref.shape, 4)
    ref.backward(gy.double())
    close('convt fwd', ncdhw(V(op.fwd(P(ndhwc(x)), w, bias, ops.ACT_NONE))), ref)
    close('convt tr', ncdhw(V(op.tr(P(ndhwc(gy)), w))), xd.grad)
    close('convt wgrad', op.wgrad(P(ndhwc(x)), P(ndhwc(gy)), tuple(w.shape)), wd.grad)


@pytest.mark.parametrize('cin,r,b', [(64, 16, 2), (24, 16, 4)])
def test_convt1_op(cin, r, b):
    from shapegan_b200 import ops
    op = ops.ConvT1Op(cin)
    x = rnd((b, cin, r, r, r), 1)
    w = rnd((cin, 1, 4, 4, 4), 2, 0.1)
    bias = rnd((1,), 3)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv_transpose3d(xd, wd, bias.double(), stride=2, padding=1)
    gy = rnd(ref.shape, 4)
    ref.backward(gy.double())
    close('convt1 fwd', op.fwd(P(ndhwc(x)), w, bias, ops.ACT_NONE), ref[:, 0])
    close('convt1 tr', ncdhw(V(op.tr(gy[:, 0].contiguous(), w))), xd.grad)
    close('convt1 wgrad', op.wgrad(P(ndhwc(x)), gy[:, 0].contiguous(), tuple(w.shape)), wd.grad)


def test_dense_ops():
    from shapegan_b200 import ops
    b = 5
    # (a) ConvTranspose3d(cin->cout,k4,s1) on a 1^3 grid: generator layer 0 / decoder.4
    for cin, cout, tag in ((128, 256, 'g0'), (256, 96, 'd1')):
        op = ops.DenseOp(64, cout, 1, 64, 1, cin, 0, cout * 64, 't_' + tag)
        x, w, bias = rnd((b, cin), 1), rnd((cin, cout, 4, 4, 4), 2, 0.05), rnd((cout,), 3)
        xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
        ref = F.conv_transpose3d(xd.reshape(b, cin, 1, 1, 1), wd, bias.double(), stride=1)       # [b,cout,4,4,4]
        gy = rnd(ref.shape, 4)
        ref.backward(gy.double())
        y = op.fwd(P(x), w, bias, ops.ACT_NONE)                                                  # [P,b,64*cout] (pos, co)
        close(tag + ' fwd', V(y).reshape(b, 4, 4, 4, cout).permute(0, 4, 1, 2, 3), ref)
        gyp = P(ndhwc(gy).reshape(b, 64 * cout))
        close(tag + ' tr', V(op.tr(gyp, w)), xd.grad)
        close(tag + ' wgrad', op.wgrad(P(x), gyp, tuple(w.shape)), wd.grad)
    # (b) Conv3d(96->256,k4,s1) on a 4^3 grid: encoder.9
    cin, cout = 96, 256
    op = ops.DenseOp(1, cout, 0, cin * 64, 64, cin, 1, 64, 't_e3')
    x, w, bias = rnd((b, cin, 4, 4, 4), 1), rnd((cout, cin, 4, 4, 4), 2, 0.05), rnd((cout,), 3)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, bias.double(), stride=1).reshape(b, cout)
    gy = rnd(ref.shape, 4)
    ref.backward(gy.double())
    xp = P(ndhwc(x).reshape(b, 64 * cin))
    close('e3 fwd', V(op.fwd(xp, w, bias, ops.ACT_NONE)), ref)
    close('e3 tr', V(op.tr(P(gy), w)).reshape(b, 4, 4, 4, cin).permute(0, 4, 1, 2, 3), xd.grad)
    close('e3 wgrad', op.wgrad(xp, P(gy), tuple(w.shape)), wd.grad)
    # (c) NCDHW flatten + Linear(16384->128): progressive head
    op = ops.DenseOp(1, 128, 0, 16384, 64, 256, 1, 64, 't_head')
    x, w, bias = rnd((b, 256, 4, 4, 4), 1), rnd((128, 16384), 2, 0.02), rnd((128,), 3)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.linear(xd.reshape(b, 16384), wd, bias.double())
    gy = rnd(ref.shape, 4)
    ref.backward(gy.double())
    xp = P(ndhwc(x).reshape(b, 16384))
    close('head fwd', V(op.fwd(xp, w, bias, ops.ACT_NONE)), ref)
    close('head tr', V(op.tr(P(gy), w)).reshape(b, 4, 4, 4, 256).permute(0, 4, 1, 2, 3), xd.grad)
    close('head wgrad', op.wgrad(xp, P(gy), tuple(w.shape)), wd.grad)
    # (d) plain nn.Linear
    for fin, fout in ((128, 256), (256, 128), (128, 128)):
        op = ops.linear_op(fin, fout, 't')
        x, w, bias = rnd((b, fin), 1), rnd((fout, fin), 2, 0.1), rnd((fout,), 3)
        xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
        ref = F.linear(xd, wd, bias.double())
        gy = rnd(ref.shape, 4)
        ref.backward(gy.double())
        close('lin fwd', V(op.fwd(P(x), w, bias, ops.ACT_NONE)), ref)
        close('lin tr', V(op.tr(P(gy), w)), xd.grad)
        close('lin wgrad', op.wgrad(P(x), P(gy), tuple(w.shape)), wd.grad)


def test_rowdot_and_batchnorm_and_fade():
    from shapegan_b200 import ops, raw
    b = 6
    # Conv3d(256->1,k4,s1) as a strided row dot
    x, w, bias = rnd((b, 256, 4, 4, 4), 1), rnd((1, 256, 4, 4, 4), 2, 0.02), rnd((1,), 3)
    xd, wd = x.double().requires_grad_(True), w.double().requires_grad_(True)
    ref = F.conv3d(xd, wd, bias.double(), stride=1).reshape(b)
    gy = rnd((b,), 4)
    ref.backward(gy.double())
    xp = P(ndhwc(x).reshape(b, 16384)).requires_grad_(True)
    wp = w.clone().requires_grad_(True)
    bp = bias.clone().requires_grad_(True)
    y = ops.rowdot(xp, wp, bp, ops.ACT_NONE, 256, 1, 64)
    close('rowdot fwd', y, ref)
    y.backward(gy)
    close('rowdot gw', wp.grad, wd.grad)
    close('rowdot gb', bp.grad, gy.sum().reshape(1))
    close('rowdot gx', V(xp.grad).reshape(b, 4, 4, 4, 256).permute(0, 4, 1, 2, 3), xd.grad)
    # train-mode BatchNorm3d + LeakyReLU for the channel counts on the path
    for c, rows_shape in ((24, (3, 8, 8, 8)), (96, (4, 4, 4, 4)), (256, (4, 1, 1, 1)), (64, (2, 16, 16, 16))):
        bn = torch.nn.BatchNorm3d(c).cuda()
        bn.weight.data = rnd((c,), 5) + 1.5
        bn.bias.data = rnd((c,), 6)
        x = rnd((rows_shape[0], c) + rows_shape[1:], 7) * 2 + 0.3
        ref_bn = torch.nn.BatchNorm3d(c).cuda().double()
        ref_bn.load_state_dict({k: v.double() if v.dtype.is_floating_point else v for k, v in bn.state_dict().items()})
        xd = x.double().requires_grad_(True)
        ref = F.leaky_relu(ref_bn(xd), 0.2)
        gy = rnd(ref.shape, 8)
        ref.backward(gy.double())
        xp = P(ndhwc(x)).requires_grad_(True)
        y = ops.batchnorm_act(xp, bn, ops.ACT_LRELU, c)
        close('bn%d fwd' % c, ncdhw(V(y)), ref)
        y.backward(P(ndhwc(gy)))
        close('bn%d gx' % c, ncdhw(V(xp.grad)), xd.grad, 5e-4)
        close('bn%d ggamma' % c, bn.weight.grad, ref_bn.weight.grad, 5e-4)
        close('bn%d gbeta' % c, bn.bias.grad, ref_bn.bias.grad, 5e-4)
        close('bn%d rmean' % c, bn.running_mean, ref_bn.running_mean)
        close('bn%d rvar' % c, bn.running_var, ref_bn.running_var)
    # eval-mode BatchNorm (running statistics are constants): forward and BACKWARD (e.g. fine-tuning through a frozen generator)
    c = 48
    bn = torch.nn.BatchNorm3d(c).cuda()
    bn.weight.data = rnd((c,), 21) + 1.5
    bn.bias.data = rnd((c,), 22)
    bn.running_mean.data = rnd((c,), 23) * 0.3
    bn.running_var.data = rnd((c,), 24) * 0.4 + 1.0
    bn.eval()
    ref_bn = torch.nn.BatchNorm3d(c).cuda().double()
    ref_bn.load_state_dict({k: v.double() if v.dtype.is_floating_point else v for k, v in bn.state_dict().items()})
    ref_bn.eval()
    x = rnd((3, c, 4, 4, 4), 25) * 2
    xd = x.double().requires_grad_(True)
    ref = F.leaky_relu(ref_bn(xd), 0.2)
    gy = rnd(ref.shape, 26)
    ref.backward(gy.double())
    xp = P(ndhwc(x)).requires_grad_(True)
    y = ops.batchnorm_act(xp, bn, ops.ACT_LRELU, c)
    close('bn-eval fwd', ncdhw(V(y)), ref)
    y.backward(P(ndhwc(gy)))
    close('bn-eval gx', ncdhw(V(xp.grad)), xd.grad, 5e-4)
    close('bn-eval ggamma', bn.weight.grad, ref_bn.weight.grad, 5e-4)
    close('bn-eval gbeta', bn.bias.grad, ref_bn.bias.grad, 5e-4)
    # fade-in blend
    bsz, r, c, f = 2, 8, 64, 0.3
    h = rnd((bsz, r, r, r, c), 9)
    vol = rnd((bsz, 2 * r, 2 * r, 2 * r), 10)
    y = ops.fade(P(h), vol, f)
    ref = f * h.double()
    ref[..., 0] += (1 - f) * vol.double()[:, ::2, ::2, ::2]
    close('fade', V(y), ref)
    # plane <-> fp32 conversion + column sums
    x = rnd((1000, 131), 11)
    xp = raw.f32_to_planes(x, 2, 192)
    close('to_planes', V(xp)[:, :131], x)
    assert float(V(xp)[:, 131:].abs().max()) == 0.0
    close('from_planes', raw.planes_to_f32(xp, 192, 131), x)
