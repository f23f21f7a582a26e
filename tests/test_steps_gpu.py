"""GPU (-m gpu): the step objects of shapegan_b200.train (flat arenas + fused optimizer kernels) against the oracle's CPU port of
the reference's script bodies (oracle/ref_steps.py), same seeded weights and inputs, fp32x mode.

After ONE optimizer step parameters agree to the absolute bound of the update itself: RMSprop/Adam's first step moves every weight
by ~lr*(sign-like) whatever |g| is, so a near-zero gradient whose sign flips under rounding moves a weight by up to 2x the step."""
import pytest
import torch

from oracle import ref_steps as S
from oracle import ref_torch as R
from oracle import shapes as TS

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32x():
    from shapegan_b200 import config
    old = config.precision()
    config.set_precision('fp32x')
    yield
    config.set_precision(old)


def load(module, shapes, seed):
    sd = R.seeded_state_dict(shapes, seed)
    module.load_state_dict(sd, strict=True)


def rnd(shape, seed, lo=-1.0, hi=1.0):
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed)) * (hi - lo) + lo


def voxels(b, r, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(torch.randn((b, r, r, r), generator=g) * 0.05, -0.1, 0.1) / 0.1


def params_close(module, sd, atol, what):
    worst = 0.0
    for k, v in module.state_dict().items():
        if 'num_batches' in k or k.startswith('optional_layer_'):
            continue
        ref = sd[k].detach()
        worst = max(worst, (v.detach().cpu() - ref).abs().max().item())
    assert worst <= atol, '%s: max parameter deviation %.3e > %.1e' % (what, worst, atol)


def test_autodecoder_step():
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    n, shapes = 4096 + 77, 5
    pts = rnd((n, 3), 1)
    sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    idx = (torch.arange(n) * shapes) // n
    table = rnd((shapes, 128), 2) * 0.3
    net = SDFNet()
    load(net, TS.sdf_shapes(), 11)
    ref = S.AutodecoderStepCPU(S.make_params(TS.sdf_shapes(), 11), table)
    step = train.AutodecoderStep(net, table.cuda())
    for _ in range(2):
        l_ref = ref(pts, sdf, idx)
        l = step(pts.cuda(), sdf.cuda(), idx.to(torch.int32).cuda())
        assert abs(l.item() - l_ref.item()) < 2e-4 * max(1.0, abs(l_ref.item()))
    params_close(net, ref.sd, 4.5e-5, 'sdfnet after 2 Adam steps (lr 1e-5)')
    assert (step.table.detach().cpu() - ref.table.detach()).abs().max().item() < 4.5e-5


def test_gan_step():
    from model.gan import Discriminator, Generator
    from shapegan_b200 import train
    b = 4
    gen, dis = Generator(), Discriminator()
    load(gen, TS.gen_shapes(), 21)
    load(dis, TS.disc_shapes(), 22)
    ref = S.GANStepCPU(S.make_params(TS.gen_shapes(), 21), S.make_params(TS.disc_shapes(), 22))
    step = train.GANStep(gen, dis)
    real, z1, z2 = voxels(b, 32, 3), rnd((b, 128), 4, -2, 2), rnd((b, 128), 5, -2, 2)
    want = ref(real, z1, z2)
    got = step(real.cuda(), z1.cuda(), z2.cuda())
    for a, c in zip(got, want):
        assert abs(a.item() - c.item()) < 2e-3 * max(1.0, abs(c.item()))
    params_close(gen, ref.g, 2.2e-3, 'generator after Adam lr 1e-3')
    params_close(dis, ref.d, 4.4e-5, 'discriminator after two Adam steps lr 1e-5')


@pytest.mark.parametrize('variational', [True, False])
def test_vae_step(variational):
    import shapegan_b200.nn.autoencoder as impl
    from model.autoencoder import Autoencoder
    from shapegan_b200 import train
    b = 4
    m = Autoencoder(is_variational=variational)
    load(m, TS.ae_shapes(variational), 31)
    ref = S.VAEStepCPU(S.make_params(TS.ae_shapes(variational), 31), variational)
    step = train.VAEStep(m)
    x = voxels(b, 32, 6)
    eps = torch.randn((b, 128), generator=torch.Generator().manual_seed(7))

    class _Fixed:
        def sample(self, shape):
            return eps.reshape(shape)
    old = impl.standard_normal_distribution
    impl.standard_normal_distribution = _Fixed()
    try:
        got = step(x.cuda())
    finally:
        impl.standard_normal_distribution = old
    want = ref(x, eps)
    assert abs(got.item() - want.item()) < 2e-3 * max(1.0, abs(want.item()))
    params_close(m, ref.sd, 1.1e-4, 'autoencoder after Adam lr 5e-5')


@pytest.mark.parametrize('it,fade', [(1, 0.4), (2, 1.0)])
def test_hybrid_progressive_step(it, fade):
    from model.progressive_gan import Discriminator
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    b = 2
    r = R.RESOLUTIONS[it]
    gen, dis = SDFNet(), Discriminator().cuda()
    load(gen, TS.sdf_shapes(), 41)
    load(dis, TS.prog_shapes(), 42)
    dis.fade_in_progress = fade
    gsd = S.make_params(TS.sdf_shapes(), 41)
    dsd = S.make_params(TS.prog_shapes(), 42)
    for i in range(4):      # load_state_dict lets the alias entries win (progressive_gan.py:41-42)
        for n in ('weight', 'bias'):
            dsd['optional_layers.%d.0.%s' % (i, n)] = dsd['optional_layer_%d.0.%s' % (i, n)]
    ref = S.HybridProgressiveStepCPU(gsd, dsd, it, fade)
    step = train.HybridProgressiveStep(gen, dis, it)
    z1, z2 = rnd((b, 128), 8, -1, 1), rnd((b, 128), 9, -1, 1)
    valid = voxels(b, r, 10) * 0.1
    alpha = rnd((b, 1, 1, 1), 11, 0, 1)
    # discriminator update (:153-166) then generator update (:136-146)
    l_ref, gp_ref = ref.discriminator_update(valid, z1, alpha)
    l, gp = step.discriminator_update(valid.cuda(), z1.cuda(), alpha.cuda())
    assert abs(gp.item() - gp_ref.item()) < 5e-3 * max(1.0, abs(gp_ref.item()))
    assert abs(l.item() - l_ref.item()) < 5e-3 * max(1.0, abs(l_ref.item()))
    g_ref = ref.generator_update(z2)
    g = step.generator_update(z2.cuda())
    assert abs(g.item() - g_ref.item()) < 5e-3 * max(1.0, abs(g_ref.item()))
    params_close(dis, ref.d, 2.2e-3, 'progressive discriminator after RMSprop lr 1e-4')     # first RMSprop step = 10*lr
    params_close(gen, ref.g, 2.2e-3, 'SDFNet generator after RMSprop lr 1e-4')


def _wgan_setup(seed_g=51, seed_c=52):
    from model.gan import Discriminator, Generator
    from shapegan_b200 import train
    gen, cri = Generator(), Discriminator()
    load(gen, TS.gen_shapes(), seed_g)
    load(cri, TS.disc_shapes(), seed_c)
    return gen, cri, train.WGANStep(gen, cri)


def test_wgan_step_graph_replay_matches_eager():
    """A captured WGANStep must re-pack the weight images INSIDE the graph: N replays == N eager steps, and a replay after the
    critic's weights were zeroed behind the graph's back must see the zeros (a cached pre-capture image would not)."""
    from shapegan_b200 import config
    config.set_precision('bf16')
    b = 4
    real, z1, z2 = voxels(b, 32, 61).cuda(), rnd((b, 128), 62, -2, 2).cuda(), rnd((b, 128), 63, -2, 2).cuda()
    gen_e, cri_e, step_e = _wgan_setup()
    gen_g, cri_g, step_g = _wgan_setup()
    losses = torch.zeros(2, device='cuda')

    def body():
        cl, gl = step_g(real, z1, z2)
        losses[0].copy_(cl); losses[1].copy_(gl)
    body()                                                    # eager warm-up step (also what bench.py does before capturing)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    for _ in range(2):
        step_e(real, z1, z2)
    for _ in range(3):
        graph.replay()
        want = step_e(real, z1, z2)
        torch.cuda.synchronize()
        assert abs(losses[0].item() - want[0].item()) < 1e-5 + 1e-4 * abs(want[0].item())
        assert abs(losses[1].item() - want[1].item()) < 1e-5 + 1e-4 * abs(want[1].item())
    for m_g, m_e in ((gen_g, gen_e), (cri_g, cri_e)):
        for (k, v), (_, w) in zip(m_g.state_dict().items(), m_e.state_dict().items()):
            if 'num_batches' in k:
                assert int(v) == int(w), k
            else:
                assert (v - w).abs().max().item() <= 1e-5, k
    with torch.no_grad():
        step_g.copt.flat.zero_()                              # every critic weight and bias := 0, behind the graph's back
    graph.replay()
    torch.cuda.synchronize()
    assert losses[0].item() == 0.0, 'the replayed critic pass ran on a weight image packed before the capture'


def test_autodecoder_step_fresh_index_tensors():
    """The per-shape point counts of the 0.01*mean(z^2) term must follow the index tensor of THIS call, also when a data
    loader hands over fresh tensors that the caching allocator places at a recycled address (train_sdf_autodecoder.py:77-80)."""
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    n, shapes = 2048, 4
    pts = rnd((n, 3), 71)
    sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    table = rnd((shapes, 128), 72) * 0.5
    net = SDFNet()
    load(net, TS.sdf_shapes(), 73)
    ref = S.AutodecoderStepCPU(S.make_params(TS.sdf_shapes(), 73), table)
    step = train.AutodecoderStep(net, table.cuda())
    for rep in range(3):
        idx = ((torch.arange(n) * shapes) // n) if rep != 1 else torch.full((n,), 3, dtype=torch.int64)    # all points of shape 3
        l_ref = ref(pts, sdf, idx)
        idx_d = idx.to(torch.int32).cuda()                    # fresh device tensor each call; the previous one is freed
        l = step(pts.cuda(), sdf.cuda(), idx_d)
        del idx_d
        assert abs(l.item() - l_ref.item()) < 2e-4 * max(1.0, abs(l_ref.item())), rep
    assert (step.table.detach().cpu() - ref.table.detach()).abs().max().item() < 7e-5


# ------------------------------------------------------------------------------------------------- step objects vs the reference's step goldens
def _after_close(g, prefix, module, atol):
    import numpy as np
    from conftest import digest_errors
    worst, n = 0.0, 0
    for k, v in module.state_dict().items():
        if prefix + k + '@sub' not in g:
            continue
        _, _, maxabs = digest_errors(g, prefix + k, v)
        worst, n = max(worst, maxabs), n + 1
    assert n > 0 and worst <= atol, '%s: max deviation %.3e > %.1e over %d tensors' % (prefix, worst, atol, n)
    return np.float64(worst)


def test_gan_step_vs_reference_golden():
    """train_gan.py:58-86 golden from the unmodified reference (oracle/gen_golden_steps.py)."""
    from conftest import load_golden
    from model.gan import Discriminator, Generator
    from shapegan_b200 import train
    g = load_golden('gan_step')
    gen, dis = Generator(), Discriminator()
    load(gen, TS.gen_shapes(), 701)
    load(dis, TS.disc_shapes(), 702)
    got = train.GANStep(gen, dis)(torch.from_numpy(g['real']).cuda(), torch.from_numpy(g['z_gen']).cuda(), torch.from_numpy(g['z_dis']).cuda())
    for a, c in zip(got, g['losses']):
        assert abs(a.item() - float(c)) < 2e-3 * max(1.0, abs(float(c)))
    _after_close(g, 'gen_after.', gen, 2.2e-3)               # Adam lr 1e-3: the first step moves every weight by ~lr
    _after_close(g, 'dis_after.', dis, 4.4e-5)               # two Adam steps of lr 1e-5


@pytest.mark.parametrize('variational', [True, False])
def test_vae_step_vs_reference_golden(variational):
    import shapegan_b200.nn.autoencoder as impl
    from conftest import load_golden
    from model.autoencoder import Autoencoder
    from shapegan_b200 import train
    g = load_golden('vae_step_%s' % ('vae' if variational else 'classic'))
    m = Autoencoder(is_variational=variational)
    load(m, TS.ae_shapes(variational), 711 + int(variational))
    eps = torch.from_numpy(g['eps'])

    class _Fixed:
        def sample(self, shape):
            return eps.reshape(shape)
    old = impl.standard_normal_distribution
    impl.standard_normal_distribution = _Fixed()
    try:
        got = train.VAEStep(m)(torch.from_numpy(g['x']).cuda())
    finally:
        impl.standard_normal_distribution = old
    assert abs(got.item() - float(g['loss'])) < 2e-3 * max(1.0, abs(float(g['loss'])))
    _after_close(g, 'after.', m, 1.1e-4)                     # Adam lr 5e-5


def test_autodecoder_step_vs_reference_golden():
    from conftest import load_golden
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    g = load_golden('autodecoder_step')
    net = SDFNet()
    load(net, TS.sdf_shapes(), 721)
    step = train.AutodecoderStep(net, torch.from_numpy(g['latent_table']).cuda())
    pts, sdf, idx = (torch.from_numpy(g[k]).cuda() for k in ('points', 'sdf', 'shape_index'))
    for want in g['losses']:
        got = step(pts, sdf, idx.to(torch.int32))
        assert abs(got.item() - float(want)) < 2e-4 * max(1.0, abs(float(want)))
    _after_close(g, 'after.', net, 4.5e-5)                   # two Adam steps of lr 1e-5
    assert (step.table.detach().cpu() - torch.from_numpy(g['latent_table_after'])).abs().max().item() < 4.5e-5


def test_hybrid_step_vs_reference_golden():
    from conftest import load_golden
    from model.progressive_gan import Discriminator
    from model.sdf_net import SDFNet
    from shapegan_b200 import train
    g = load_golden('hybrid_step_it1')
    gen, dis = SDFNet(), Discriminator().cuda()
    load(gen, TS.sdf_shapes(), 731)
    load(dis, TS.prog_shapes(), 732)
    dis.fade_in_progress = float(g['fade'])
    step = train.HybridProgressiveStep(gen, dis, int(g['iteration']))
    dl, gp = step.discriminator_update(torch.from_numpy(g['valid']).cuda(), torch.from_numpy(g['z_dis']).cuda(), torch.from_numpy(g['alpha']).cuda())
    gl = step.generator_update(torch.from_numpy(g['z_gen']).cuda())
    for a, c in zip((dl, gp, gl), g['losses']):
        assert abs(a.item() - float(c)) < 5e-3 * max(1.0, abs(float(c)))
    _after_close(g, 'dis_after.', dis, 2.2e-3)               # first RMSprop step = 10 lr = 1e-3 per weight
    _after_close(g, 'gen_after.', gen, 2.2e-3)


def test_autodecoder_uniform_segments_equal_general_path():
    """points_per_shape = P (per-shape latent gradients: segmented sums of g_1 / g_5 + [S, .] matrix products) against the general path
    (K = 131 input-gradient GEMMs + scatter by index) on the same batch, bf16 fused kernels: loss, every parameter and the table."""
    from model.sdf_net import SDFNet
    from shapegan_b200 import config, train
    config.set_precision('bf16')
    shapes, per = 6, 1024
    n = shapes * per
    pts = rnd((n, 3), 81).cuda()
    sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    idx = (torch.arange(n) // per).to(torch.int32).cuda()
    table = (rnd((shapes, 128), 82) * 0.3).cuda()
    res = []
    for pps in (per, 0):
        net = SDFNet()
        load(net, TS.sdf_shapes(), 83)
        step = train.AutodecoderStep(net, table, points_per_shape=pps)
        loss = step(pts, sdf, idx)
        grads = {k: p.grad.clone() for k, p in net.named_parameters()}
        res.append((loss.item(), grads, step.lopt.grad.clone()))
    assert abs(res[0][0] - res[1][0]) < 1e-6
    for k in res[0][1]:
        a, b = res[0][1][k], res[1][1][k]
        err = (a - b).norm().item() / max(b.norm().item(), 1e-30)
        assert err < 2e-2, (k, err)                   # bf16 g planes: the general path rounds g W to bf16 per point before summing
    lat_err = (res[0][2] - res[1][2]).norm().item() / res[1][2].norm().item()
    assert lat_err < 2e-2, lat_err
