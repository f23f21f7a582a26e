"""CPU (-m "not gpu"): pin oracle/ref_torch.py against the golden vectors that
oracle/gen_golden.py produced by running the UNMODIFIED reference modules.
Tolerance: the oracle and the reference call the same torch CPU kernels in the same
order, so agreement is ~1e-6 relative (fp32 summation-order noise only)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import check_digest, load_golden, rel_l2
from oracle import ref_torch as R

TOL = 2e-5


def sd_from(shapes_module_sd, seed):
    return R.seeded_state_dict(shapes_module_sd, seed)


from oracle.shapes import ae_shapes, disc_shapes, gen_shapes, prog_shapes, sdf_shapes  # noqa: E402,F401


def ordered(shapes, order):
    """seeded_state_dict draws in dict order; the golden used the reference's state_dict() order."""
    return {k: shapes[k] for k in order}


def test_voxel_coordinates_bit_exact():
    g = load_golden('voxel_coordinates')
    for r in (8, 16):
        assert np.array_equal(R.voxel_coordinates(r).numpy(), g['coords_%d' % r])
    for r in (32, 64):
        assert hashlib.sha256(R.voxel_coordinates(r).numpy().tobytes()).hexdigest() == str(g['coords_%d_sha256' % r])
    # row index <-> (i,j,k) with z fastest  (util.py:60-74)
    c = R.voxel_coordinates(8)
    lin = np.linspace(-1, 1, 8).astype(np.float32)
    for (i, j, k) in ((0, 0, 1), (3, 5, 7), (7, 0, 2)):
        assert tuple(c[i * 64 + j * 8 + k].tolist()) == (lin[i], lin[j], lin[k])


def test_sdfnet_seeded():
    g = load_golden('sdfnet_seeded')
    sd = R.seeded_state_dict(sdf_shapes(), int(g['seed_weights']))
    for v in sd.values():
        v.requires_grad_(True)
    pts = torch.from_numpy(g['points']).requires_grad_(True)
    table = torch.from_numpy(g['latent_table']).requires_grad_(True)
    idx = torch.from_numpy(g['shape_index'])
    out = R.sdfnet_forward(sd, pts, table[idx])
    assert rel_l2(out, g['out']) < TOL
    loss = R.sdfnet_autodecoder_loss(sd, pts, table, idx, torch.from_numpy(g['target']))
    assert abs(loss.item() - float(g['loss'])) < 1e-6
    loss.backward()
    assert rel_l2(table.grad, g['grad_latent_table']) < TOL
    assert rel_l2(pts.grad, g['grad_points']) < TOL
    for k, v in sd.items():
        check_digest(g, 'grad.' + k, v.grad, TOL, 'sdfnet')
    assert list(R.sdfnet_forward(sd, pts[:1], table[idx][:1]).shape) == list(g['out_n1_shape']) == []
    assert list(R.sdfnet_forward(sd, pts[:0], table[idx][:0]).shape) == list(g['out_n0_shape']) == [0]


def test_sdfnet_latent0_and_chairs():
    g = load_golden('sdfnet_latent0')
    sd = R.seeded_state_dict(sdf_shapes(0), int(g['seed_weights']))
    pts = torch.from_numpy(g['points'])
    assert rel_l2(R.sdfnet_forward(sd, pts, torch.zeros((pts.shape[0], 0))), g['out']) < TOL
    g = load_golden('sdfnet_chairs')
    sd = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('w.')}
    grid = R.voxel_coordinates(32)
    out = R.sdfnet_forward(sd, grid, torch.from_numpy(g['z']).repeat(grid.shape[0], 1))
    assert rel_l2(out, g['out']) < TOL
    # SURVEY §8c known-answer on the chairs checkpoint
    assert abs(out.sum().item() - 2386.4048) < 0.05 and abs(out.min().item() + 0.1375) < 1e-3


def test_generator():
    g = load_golden('gan_generator')
    sd = R.seeded_state_dict(gen_shapes(), int(g['seed_weights']))
    params = [k for k in sd if 'running' not in k and 'num_batches' not in k]
    for k in params:
        sd[k].requires_grad_(True)
    z = torch.from_numpy(g['z'])
    stats = {}
    out = R.generator_forward(sd, z, training=True, stats_out=stats)
    assert rel_l2(out, g['out_train']) < TOL
    wout = torch.rand((4, 1, 32, 32, 32), generator=torch.Generator().manual_seed(int(g['seed_wout']))) * 2 - 1
    (out * wout).sum().backward()
    for k in params:
        check_digest(g, 'grad.' + k, sd[k].grad, 5e-5, 'generator', atol=1e-2)
    for k, v in stats.items():
        assert rel_l2(v, g['after.' + k]) < TOL, k
    sd.update(stats)      # the reference's eval pass ran after the train-mode pass had updated the running stats
    with torch.no_grad():
        assert rel_l2(R.generator_forward(sd, z, training=False), g['out_eval']) < TOL


def test_discriminator_and_gp():
    g = load_golden('gan_discriminator')
    sd = R.seeded_state_dict(disc_shapes(), int(g['seed_weights']))
    for v in sd.values():
        v.requires_grad_(True)
    real, fake = torch.from_numpy(g['real']), torch.from_numpy(g['fake'])
    assert rel_l2(R.discriminator_forward(sd, real, True), g['out_sigmoid']) < TOL
    assert list(R.discriminator_forward(sd, real[:1], True).shape) == list(g['out_b1_shape']) == []
    fake_g = fake.clone().requires_grad_(True)
    of, orl = R.discriminator_forward(sd, fake_g, False), R.discriminator_forward(sd, real, False)
    assert rel_l2(of, g['out_fake']) < TOL and rel_l2(orl, g['out_real']) < TOL
    (of.mean() - orl.mean()).backward()
    for k, v in sd.items():
        check_digest(g, 'grad.' + k, v.grad, 5e-5, 'disc')
        v.grad = None
    check_digest(g, 'grad_fake', fake_g.grad, 5e-5)
    gp = R.gradient_penalty(lambda x: R.discriminator_forward(sd, x, False), real, fake.squeeze(1),
                            torch.from_numpy(g['alpha']))
    assert abs(gp.item() - float(g['gp'])) / float(g['gp']) < 1e-5
    gp.backward()
    for k, v in sd.items():
        if 'bias' in k:
            # the GP contributes exactly zero gradient to every bias
            assert v.grad is None or float(v.grad.abs().max()) == 0.0
        else:
            check_digest(g, 'gp_grad.' + k, v.grad, 5e-5, 'gp')


@pytest.mark.parametrize('name', ['progressive_disc_it0_f100', 'progressive_disc_it1_f030', 'progressive_disc_it2_f060',
                                  'progressive_disc_it2_f100', 'progressive_disc_it3_f025'])
def test_progressive_discriminator(name):
    g = load_golden(name)
    it, fade = int(g['iteration']), float(g['fade'])
    sd = R.seeded_state_dict(prog_shapes(), int(g['seed_weights']))
    # the reference state_dict lists the four conv blocks twice (alias, progressive_gan.py:41-42);
    # load_state_dict applies optional_layer_{i} after optional_layers.{i}: the alias copy wins.
    for i in range(4):
        for n in ('weight', 'bias'):
            sd['optional_layers.%d.0.%s' % (i, n)] = sd['optional_layer_%d.0.%s' % (i, n)]
    keys = ['head.1.weight', 'head.1.bias', 'head.3.weight', 'head.3.bias'] + \
           ['optional_layers.%d.0.%s' % (i, n) for i in range(4) for n in ('weight', 'bias')]
    for k in keys:
        sd[k].requires_grad_(True)
    real, fake = torch.from_numpy(g['real']), torch.from_numpy(g['fake'])
    fake_g = fake.clone().requires_grad_(True)
    fn = lambda x: R.progressive_discriminator_forward(sd, x, it, fade)   # noqa: E731
    of, orl = fn(fake_g), fn(real)
    assert rel_l2(of, g['out_fake']) < TOL and rel_l2(orl, g['out_real']) < TOL
    gp = R.gradient_penalty(fn, real, fake, torch.from_numpy(g['alpha']))
    assert abs(gp.item() - float(g['gp'])) / float(g['gp']) < 1e-5
    (of.mean() - orl.mean() + gp).backward()
    for k in keys:
        gk = 'grad.' + k
        if gk + '@sub' in g:
            check_digest(g, gk, sd[k].grad, 1e-4, name)
    check_digest(g, 'grad_fake', fake_g.grad, 1e-4)


@pytest.mark.parametrize('variational', [True, False])
def test_autoencoder(variational):
    g = load_golden('autoencoder_vae' if variational else 'autoencoder_classic')
    sd = R.seeded_state_dict(ae_shapes(variational), int(g['seed_weights']))
    params = [k for k in sd if 'running' not in k and 'num_batches' not in k]
    for k in params:
        sd[k].requires_grad_(True)
    x, eps = torch.from_numpy(g['x']), torch.from_numpy(g['eps'])
    stats = {}
    if variational:
        out, mean, logvar = R.autoencoder_forward(sd, x, True, True, eps, stats)
        assert rel_l2(mean, g['mean']) < TOL and rel_l2(logvar, g['log_variance']) < TOL
        loss = R.reconstruction_loss(out, x) + R.kld_loss(mean, logvar)
    else:
        out = R.autoencoder_forward(sd, x, False, True, None, stats)
        loss = R.reconstruction_loss(out, x)
    assert rel_l2(out, g['out_train']) < TOL
    assert abs(loss.item() - float(g['loss'])) / abs(float(g['loss'])) < 1e-5
    loss.backward()
    for k in params:
        check_digest(g, 'grad.' + k, sd[k].grad, 2e-4, 'ae', atol=1e-5)
    for k, v in stats.items():
        if 'running' in k:
            check_digest(g, 'after.' + k, v, TOL)
    sd.update(stats)
    with torch.no_grad():
        o = R.autoencoder_forward(sd, x, variational, False)
        assert rel_l2(o[0] if variational else o, g['out_eval']) < TOL


def test_render_oracle_pinned_to_chairs_golden():
    """oracle/ref_render.py: voxelise() reproduces the reference's chairs grid evaluation; march() / normals() are consistent with it
    (a marched hit point has |sdf| below the threshold band, normals are unit length and point along increasing sdf)."""
    from oracle import ref_render as RR
    g = load_golden('sdfnet_chairs')
    sd = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('w.')}
    z = torch.from_numpy(g['z'])
    full = RR.voxelise(sd, z, 32, sphere_only=False, pad=False)
    assert rel_l2(torch.from_numpy(full).reshape(-1), g['out']) < TOL
    vox = RR.voxelise(sd, z, 32, sphere_only=True)
    mask, _ = RR.sphere_mask(32)
    m3 = mask.reshape(32, 32, 32)
    assert int(mask.sum()) == 20360 and np.all(vox[~m3] == 1.0) and np.allclose(vox[m3], full[m3], atol=1e-6)
    pts = torch.tensor([[0.9, 0.0, 0.0], [0.0, 0.9, 0.1], [-0.8, 0.2, 0.3]])
    sdf, nrm = RR.normals(sd, z, pts)
    assert torch.allclose(nrm.norm(dim=1), torch.ones(3), atol=1e-5)
    eps = 1e-3
    sdf2 = R.sdfnet_forward(sd, pts + eps * nrm, z.reshape(1, -1).repeat(3, 1))
    assert torch.all(sdf2 > sdf)                                   # the normal is the direction of increasing distance
    # ingest == datasets.py:16-23
    raw = np.array([0.3, -0.2, 0.05, 0.1], dtype=np.float32)
    assert torch.equal(RR.ingest(raw), torch.tensor([1.0, -1.0, np.float32(0.05) / np.float32(0.1), 1.0]))


def test_point_gan_oracle_pinned_to_reference_golden():
    """oracle/ref_points.py against the golden of the unmodified model/point_sdf_net.py + train_point_gan.py:52-87"""
    from oracle import ref_points as RP
    g = load_golden('point_gan')
    shapes_g = {'lins.%d.weight' % i: s for i, s in enumerate([(256, 3), (256, 256), (256, 256), (256, 256), (256, 259), (256, 256), (256, 256), (1, 256)])}
    gsd, dsd = {}, {}
    # rebuild the seeded state dicts in the reference modules' state_dict() ORDER
    order_g = []
    for i in range(8):
        order_g += [('lins.%d.weight' % i, shapes_g['lins.%d.weight' % i]), ('lins.%d.bias' % i, (shapes_g['lins.%d.weight' % i][0],))]
    for i in range(8):
        c = 1 if i == 7 else 256
        order_g += [('norms.%d.weight' % i, (c,)), ('norms.%d.bias' % i, (c,))]
    order_g += [('z_lin1.weight', (256, 128)), ('z_lin1.bias', (256,)), ('z_lin2.weight', (256, 128)), ('z_lin2.bias', (256,))]
    gsd = R.seeded_state_dict(dict(order_g), int(g['seed_gen']))
    order_d = []
    for i, (a, b) in zip((0, 2, 4, 6), ((4, 64), (64, 128), (128, 256), (256, 512))):
        order_d += [('nn1.%d.weight' % i, (b, a)), ('nn1.%d.bias' % i, (b,))]
    for i, (a, b) in zip((0, 2, 4), ((512, 256), (256, 128), (128, 1))):
        order_d += [('nn2.%d.weight' % i, (b, a)), ('nn2.%d.bias' % i, (b,))]
    dsd = R.seeded_state_dict(dict(order_d), int(g['seed_dis']))
    pos, dist, alpha = (torch.from_numpy(g[k]) for k in ('pos', 'dist', 'alpha'))
    fake = RP.sdf_generator_forward(gsd, pos, torch.from_numpy(g['z_dis']))
    assert rel_l2(fake, g['fake']) < TOL
    assert rel_l2(RP.pointnet_forward(dsd, pos, dist), g['out_real']) < TOL
    for v in dsd.values():
        v.requires_grad_(True)
    total, d_loss, gp = RP.critic_loss_with_gp(dsd, pos, dist, fake.detach(), alpha)
    assert abs(d_loss.item() - float(g['d_loss'])) < 1e-5 and abs(gp.item() - float(g['gp'])) < 1e-4 * max(1.0, float(g['gp']))
    total.backward()
    for k, v in dsd.items():
        check_digest(g, 'dis_grad.' + k, v.grad, 2e-4, k, atol=1e-7)
