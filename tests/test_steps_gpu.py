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
