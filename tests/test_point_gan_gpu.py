"""GPU (-m gpu): the point-set GAN path -- SDFGenerator / PointNet of model/point_sdf_net.py and the step bodies of
train_point_gan.py:52-87 (critic loss + gradient penalty w.r.t. the interpolated distances, generator loss) -- against the golden
of the unmodified reference module (tests/golden/point_gan.npz, oracle/gen_golden_points.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import ref_torch as R
from test_parity_gpu import Checker, check_dev, cu, prec, seeded_load  # noqa: F401  (prec: fixture)

pytestmark = pytest.mark.gpu


def test_point_gan_steps(prec):
    from model.point_sdf_net import PointNet, SDFGenerator
    g = load_golden('point_gan')
    c = Checker(prec, 'point_gan', g)
    gen = SDFGenerator(128, 256, 8, True, dropout=0.0).cuda()
    dis = PointNet(out_channels=1).cuda()
    seeded_load(gen, int(g['seed_gen']))
    seeded_load(dis, int(g['seed_dis']))
    pos, dist, alpha = cu(g['pos']), cu(g['dist']), cu(g['alpha'])
    # ---- critic step (train_point_gan.py:52-76)
    fake = gen(pos, cu(g['z_dis']))
    assert tuple(fake.shape) == (3, 256, 1)
    c.out('fake', fake)
    out_real, out_fake = dis(pos, dist), dis(pos, fake)
    assert tuple(out_real.shape) == (3, 1)
    c.out('out_real', out_real)
    c.out('out_fake', out_fake)
    d_loss = out_fake.mean() - out_real.mean()
    inter = (alpha * dist + (1 - alpha) * fake).detach().requires_grad_(True)
    out = dis(pos, inter)
    grad = torch.autograd.grad(out, inter, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True, only_inputs=True)[0]
    c.grad('gp_input_grad', grad)
    gp = 10 * ((grad.view(grad.size(0), -1).norm(dim=-1, p=2) - 1).pow(2).mean())
    c.scalar('gp', gp.item())
    c.scalar('d_loss', d_loss.item())
    dis.zero_grad(); gen.zero_grad()
    (d_loss + gp).backward()
    c.params('dis_grad.', dis)
    # ---- generator step (:80-86)
    dis.zero_grad(); gen.zero_grad()
    loss = -dis(pos, gen(pos, cu(g['z_gen']))).mean()
    c.scalar('g_loss', loss.item())
    loss.backward()
    c.params('gen_grad.', gen)
    assert gen.norms[7].weight.grad is None            # constructed but unused, like the reference (point_sdf_net.py:110)
    c.done()
    check_dev()


def test_pointnet_pooling_kernels():
    """segment max (+ first arg-max), its scatter / gather pair, the per-shape vector add and its segmented column sums, LayerNorm+ReLU
    forward / backward against plain torch on the same bf16-rounded values"""
    from shapegan_b200 import ops, point_ops, raw
    from shapegan_b200 import config
    old = config.precision()
    config.set_precision('bf16')
    try:
        gen = torch.Generator().manual_seed(9)
        segs, n, ch = 5, 333, 136
        x = torch.randn((segs * n, ch), generator=gen).cuda()
        xp = raw.to_planes(x, 1)
        xr = raw.from_planes(xp)
        # max pooling
        out, arg = point_ops._SegMax.apply(xp, n)
        ref, ridx = xr.reshape(segs, n, ch).max(dim=1)
        assert torch.equal(raw.from_planes(out), ref) and torch.equal(arg.long(), ridx)
        small = raw.to_planes(torch.randn((segs, ch), generator=gen).cuda(), 1)
        big = point_ops._SegScatter.apply(small, arg, n, segs * n)
        dense = torch.zeros((segs, n, ch), device='cuda').scatter_(1, ridx.unsqueeze(1), raw.from_planes(small).unsqueeze(1))
        assert torch.equal(raw.from_planes(big).reshape(segs, n, ch), dense)
        assert torch.equal(raw.from_planes(point_ops._SegGather.apply(big, arg, n)), raw.from_planes(small))
        # per-shape vector add + segmented sums
        v = torch.randn((segs, ch), generator=gen).cuda()
        y = point_ops._RowsAddVec.apply(xp, v, n)
        want = (xr.reshape(segs, n, ch) + v.unsqueeze(1)).reshape(-1, ch)
        assert (raw.from_planes(y) - want).abs().max().item() < 2e-2
        assert rel_l2(raw.segment_colsum(xp, segs, n), xr.reshape(segs, n, ch).sum(1)) < 1e-5
        long = raw.to_planes(torch.randn((2 * 40000, 64), generator=gen).cuda(), 1)          # few long segments: split over rows + fold
        assert rel_l2(raw.segment_colsum(long, 2, 40000), raw.from_planes(long).reshape(2, 40000, 64).sum(1)) < 1e-5
        # LayerNorm + ReLU
        ln = torch.nn.LayerNorm(ch).cuda()
        with torch.no_grad():
            ln.weight.uniform_(0.5, 1.5); ln.bias.uniform_(-0.2, 0.2)
        xq = xp.clone().requires_grad_(True)
        yq = point_ops.layernorm_act(xq, ln, ops.ACT_RELU)
        xt = xr.clone().requires_grad_(True)
        yt = torch.relu(torch.nn.functional.layer_norm(xt, (ch,), ln.weight, ln.bias, ln.eps))
        assert rel_l2(raw.from_planes(yq), yt) < 5e-3
        gy = torch.randn_like(xr)
        gw0, gb0 = torch.autograd.grad(yt, (ln.weight, ln.bias), gy, retain_graph=True)
        gx0 = torch.autograd.grad(yt, xt, gy)[0]
        ln.zero_grad()
        yq.backward(raw.to_planes(gy, 1))
        assert rel_l2(raw.from_planes(xq.grad), gx0) < 2e-2
        assert rel_l2(ln.weight.grad, gw0) < 2e-2 and rel_l2(ln.bias.grad, gb0) < 2e-2
        check_dev()
    finally:
        config.set_precision(old)
