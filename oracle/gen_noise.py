"""ORACLE fixture generator — test infrastructure, NOT product code.

The reference's OWN fp32 round-off, per tensor: every golden case of oracle/gen_golden.py is run twice through the UNMODIFIED
reference modules — once in float32 (what the goldens hold) and once in float64 on the same weights and inputs — and the relative
L2 distance fp32-vs-fp64 of every output / gradient tensor is written to tests/golden/fp32_self_noise.json.

Why: BASELINE.json asks for "within 1e-3 relative fp32".  A gradient tensor whose fp32 value in the reference is itself only
accurate to, say, 4e-3 (long reductions behind BatchNorms, the cancellation in mean D(fake) - mean D(real), the (|g|-1) factor of
the gradient penalty) cannot be matched to 1e-3 by ANY other fp32-accumulating implementation; tests/test_parity_gpu.py therefore
gates a tensor at max(1e-3, 3 x its self-noise) and fails loudly otherwise.

Run:  python oracle/gen_noise.py   (needs /root/reference; the JSON is committed)
"""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import gen_golden as G  # noqa: E402
from oracle import ref_torch as R  # noqa: E402
from oracle.gen_golden import rnd, synth_voxels  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def load(module, seed, dtype):
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    sd = R.seeded_state_dict(shapes, seed)
    module.load_state_dict(sd, strict=True)
    return module.to(dtype)


def grads(module, prefix, out):
    seen = set()
    for k, p in module.named_parameters():
        if p.grad is not None and id(p) not in seen:
            seen.add(id(p))
            out[prefix + k] = p.grad.detach().clone()


def run_cases(mods, dt):
    gan, pg, ae, sn, util = mods
    res = {}
    c = lambda t: t.to(dt)      # noqa: E731
    # ---- sdfnet_seeded
    o = {}
    net = load(sn.SDFNet(device='cpu'), 101, dt)
    n, shapes = 1000, 4
    pts = c(rnd((n, 3), 102)).requires_grad_(True)
    table = c(rnd((shapes, 128), 103) * 0.5).requires_grad_(True)
    idx = (torch.arange(n) * shapes) // n
    target = torch.clamp(pts.detach().norm(dim=1) - 0.5, -0.1, 0.1)
    z = table[idx, :]
    out = net(pts, z)
    (torch.mean(torch.abs(out - target)) + 0.01 * torch.mean(torch.pow(z, 2))).backward()
    o['out'], o['grad_latent_table'], o['grad_points'] = out, table.grad, pts.grad
    grads(net, 'grad.', o)
    res['sdfnet_seeded'] = o
    # ---- gan_generator
    o = {}
    gen = load(gan.Generator(), 201, dt)
    b = 4
    gen.train()
    out = gen(c(rnd((b, 128), 202, -2, 2)))
    (out * c(rnd((b, 1, 32, 32, 32), 203))).sum().backward()
    o['out_train'] = out
    grads(gen, 'grad.', o)
    res['gan_generator'] = o
    # ---- gan_discriminator
    o = {}
    dis = load(gan.Discriminator(), 301, dt)
    real = c(synth_voxels(b, 32, 302))
    fake = c(torch.tanh(rnd((b, 1, 32, 32, 32), 303, -1.5, 1.5)))
    dis.use_sigmoid = False
    fake_g = fake.clone().requires_grad_(True)
    of, orl = dis(fake_g), dis(real)
    (torch.mean(of) - torch.mean(orl)).backward()
    o['out_fake'], o['out_real'], o['grad_fake'] = of, orl, fake_g.grad
    grads(dis, 'grad.', o)
    dis.zero_grad()
    dis.use_sigmoid = True
    bce = torch.nn.functional.binary_cross_entropy
    (bce(dis(fake), torch.zeros(b, dtype=dt)) + bce(dis(real), torch.ones(b, dtype=dt))).backward()
    grads(dis, 'bce_grad.', o)
    dis.zero_grad()
    dis.use_sigmoid = False
    alpha = c(rnd((b, 1, 1, 1), 304, 0, 1))
    a = alpha.expand(real.shape)
    xi = (a * real + (1 - a) * fake.squeeze(1)).requires_grad_(True)
    oo = dis(xi)
    g = torch.autograd.grad(outputs=oo, inputs=xi, grad_outputs=torch.ones(oo.shape, dtype=dt), create_graph=True, retain_graph=True,
                            only_inputs=True)[0]
    gp = ((g.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * 10
    gp.backward()
    o['gp'], o['gp_input_grad'] = gp.reshape(1), g
    grads(dis, 'gp_grad.', o)
    res['gan_discriminator'] = o
    # ---- progressive discriminators
    for it, bb, fade in ((0, 3, 1.0), (1, 3, 0.3), (2, 3, 0.6), (3, 2, 0.25), (2, 2, 1.0)):
        o = {}
        d = load(pg.Discriminator(), 400 + it, dt)
        d.set_iteration(it)
        d.fade_in_progress = fade
        r = pg.RESOLUTIONS[it]
        real = c(synth_voxels(bb, r, 410 + it))
        fake = c(torch.clamp(rnd((bb, r, r, r), 420 + it, -0.12, 0.12), -0.1, 0.1))
        alpha = c(rnd((bb, 1, 1, 1), 430 + it, 0, 1))
        fake_g = fake.clone().requires_grad_(True)
        of, orl = d(fake_g), d(real)
        a = alpha.expand(real.shape)
        xi = (a * real + (1 - a) * fake).detach().requires_grad_(True)
        oo = d(xi)
        g = torch.autograd.grad(outputs=oo, inputs=xi, grad_outputs=torch.ones(oo.shape, dtype=dt), create_graph=True, retain_graph=True,
                                only_inputs=True)[0]
        gp = ((g.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * 10
        (of.mean() - orl.mean() + gp).backward()
        o['out_fake'], o['out_real'], o['gp'], o['grad_fake'] = of, orl, gp.reshape(1), fake_g.grad
        grads(d, 'grad.', o)
        res['progressive_disc_it%d_f%03d' % (it, int(fade * 100))] = o
    # ---- autoencoders
    for variational in (True, False):
        o = {}
        m = load(ae.Autoencoder(is_variational=variational), 500 + int(variational), dt)
        x = c(synth_voxels(b, 32, 510))
        eps = c(torch.randn((b, 128), generator=torch.Generator().manual_seed(511)))
        m.train()
        if variational:
            class _Fixed:
                def sample(self, shape):
                    return eps.reshape(shape)
            ae.standard_normal_distribution = _Fixed()
            out, mean, logvar = m(x)
            o['mean'], o['log_variance'] = mean, logvar
            kld = -0.5 * torch.sum(1 + logvar - mean.pow(2) - logvar.exp()) / mean.nelement()
        else:
            out, kld = m(x), 0
        diff = out - x
        diff = torch.where(x < 0, diff * 32, diff)
        loss = torch.mean(torch.abs(diff)) + kld
        loss.backward()
        o['out_train'], o['loss'] = out, loss.reshape(1)
        grads(m, 'grad.', o)
        res['autoencoder_%s' % ('vae' if variational else 'classic')] = o
    # ---- wgan_step_b64: critic pass of train_wgan.py:65-69 at BASELINE configs[1]'s batch (keys of tests/golden/wgan_step_b64.npz)
    o = {}
    bq = 64
    gen, cri = load(gan.Generator(), 601, dt), load(gan.Discriminator(), 602, dt)
    cri.use_sigmoid = False
    z1, batch = c(rnd((bq, 128), 743, -2, 2)), c(synth_voxels(bq, 32, 745))
    gen.train()
    fake = gen(z1).detach()
    closs = torch.mean(cri(fake)) - torch.mean(cri(batch))
    closs.backward()
    o['critic_loss'] = closs.reshape(1)
    grads(cri, 'critic_grad.', o)
    res['wgan_step_b64'] = o
    return res


def main():
    torch.manual_seed(0)
    mods = G.import_reference()
    r32 = run_cases(mods, torch.float32)
    r64 = run_cases(mods, torch.float64)
    table = {'_meta': {'what': 'rel-L2 of the reference modules in float32 vs float64, same weights and inputs (oracle/gen_noise.py)',
                       'torch': torch.__version__, 'threads': torch.get_num_threads()}}
    for case in r32:
        table[case] = {k: rel(r32[case][k], r64[case][k]) for k in r32[case]}
    path = os.path.join(REPO, 'tests', 'golden', 'fp32_self_noise.json')
    with open(path, 'w') as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print('wrote', path)
    for case, d in table.items():
        if case != '_meta':
            worst = sorted(d.items(), key=lambda kv: -kv[1])[:3]
            print('%-28s worst: %s' % (case, ', '.join('%s %.1e' % kv for kv in worst)))


if __name__ == '__main__':
    main()
