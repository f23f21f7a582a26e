"""ORACLE fixture generator for the training-step bodies — test infrastructure, NOT product code.

Runs the UNMODIFIED reference modules (/root/reference, imported with the shims of oracle/gen_golden.py) through the step bodies of
the reference's scripts, restated here line by line because the scripts execute on import (datasets, viewers, argv):

  gan_step.npz            train_gan.py:58-86               Adam 1e-3 / 1e-5, BCE                  B=4
  vae_step_{vae,classic}  train_autoencoder.py:98-117      Adam 5e-5                              B=4
  autodecoder_step.npz    train_sdf_autodecoder.py:77-91   two Adams 1e-5 (`//` at :78) 2 steps, 2048 points x 4 shapes
  hybrid_step_it1.npz     train_hybrid_progressive_gan.py:134-166   RMSprop 1e-4, GP (alpha injected)  it=1 (16^3), B=2, fade 0.4
  wgan_step_b64.npz       train_wgan.py:62-84 at BASELINE configs[1]'s batch (B=64), digests only

They pin oracle/ref_steps.py (tests/test_oracle_golden.py, CPU) and the CUDA step objects (tests/test_steps_gpu.py).
Run:  python oracle/gen_golden_steps.py      (needs /root/reference; outputs are committed)
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import gen_golden as G  # noqa: E402
from oracle.gen_golden import load_seeded, put, rnd, save, synth_voxels  # noqa: E402


def after(store, prefix, module):
    seen = set()
    for k, v in module.state_dict().items():
        if k.startswith('optional_layer_'):          # alias entries of progressive_gan.py:41-42
            continue
        if id(v) in seen:
            continue
        seen.add(id(v))
        put(store, prefix + k, v)


def main():
    torch.manual_seed(0)
    gan, pg, ae, sn, util = G.import_reference()
    b = 4

    # ---------------------------------------------------------------- train_gan.py:58-86
    gen, dis = gan.Generator(), gan.Discriminator()
    load_seeded(gen, 701)
    load_seeded(dis, 702)
    gopt = torch.optim.Adam(gen.parameters(), lr=0.001)                                   # :28
    dopt = torch.optim.Adam(dis.parameters(), lr=0.00001)                                 # :31
    bce = torch.nn.functional.binary_cross_entropy                                        # :30
    real, z1, z2 = synth_voxels(b, 32, 703), rnd((b, 128), 704, -2, 2), rnd((b, 128), 705, -2, 2)
    s = {'real': real.numpy(), 'z_gen': z1.numpy(), 'z_dis': z2.numpy()}
    gopt.zero_grad()                                                                      # :61
    fake = gen(z1)                                                                        # :63 generate() == self(z), gan.py:33-34
    gl = -torch.mean(torch.log(dis(fake)))                                                # :67-68
    gl.backward(); gopt.step()                                                            # :69-70
    dopt.zero_grad()                                                                      # :78
    fake = gen(z2).detach()                                                               # :79
    fl = bce(dis(fake), torch.zeros(b))                                                   # :80-81
    fl.backward(); dopt.step()                                                            # :82-83
    dopt.zero_grad()                                                                      # :85
    vl = bce(dis(real), torch.ones(b))                                                    # :86-87
    vl.backward(); dopt.step()                                                            # :88-89
    s['losses'] = np.array([gl.item(), fl.item(), vl.item()], dtype=np.float64)
    after(s, 'gen_after.', gen)
    after(s, 'dis_after.', dis)
    save('gan_step', s)

    # ---------------------------------------------------------------- train_autoencoder.py:98-117
    for variational in (True, False):
        m = ae.Autoencoder(is_variational=variational)
        load_seeded(m, 711 + int(variational))
        opt = torch.optim.Adam(m.parameters(), lr=0.00005)                                # :35
        x = synth_voxels(b, 32, 713)
        eps = torch.randn((b, 128), generator=torch.Generator().manual_seed(714))
        s = {'x': x.numpy(), 'eps': eps.numpy()}

        class _Fixed:
            def sample(self, shape):
                return eps.reshape(shape)
        ae.standard_normal_distribution = _Fixed()
        m.zero_grad(); m.train()                                                          # :102-103
        if variational:
            out, mean, logvar = m(x)                                                      # :105
            kld = -0.5 * torch.sum(1 + logvar - mean.pow(2) - logvar.exp()) / mean.nelement()      # :54-55
        else:
            out, kld = m(x), 0
        diff = out - x                                                                    # :57-62
        diff[x < 0] *= 32
        loss = torch.mean(torch.abs(diff)) + kld                                          # :113
        loss.backward(); opt.step()                                                       # :118-119
        s['loss'] = np.float64(loss.item())
        after(s, 'after.', m)
        save('vae_step_%s' % ('vae' if variational else 'classic'), s)

    # ---------------------------------------------------------------- train_sdf_autodecoder.py:77-91 (two steps)
    net = sn.SDFNet(device='cpu')
    load_seeded(net, 721)
    n, shapes = 2048, 4
    pts = rnd((n, 3), 722)
    sdf = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    table = (rnd((shapes, 128), 723) * 0.3).requires_grad_(True)
    nopt = torch.optim.Adam(net.parameters(), lr=1e-5)                                    # :44
    lopt = torch.optim.Adam([table], lr=1e-5)                                             # :45
    idx = (torch.arange(n) * shapes) // n                                                 # :78 with // (D6)
    s = {'points': pts.numpy(), 'sdf': sdf.numpy(), 'latent_table': table.detach().numpy().copy(), 'shape_index': idx.numpy()}
    losses = []
    for _ in range(2):
        z = table[idx, :]                                                                 # :80
        net.zero_grad()                                                                   # :84
        if table.grad is not None:
            table.grad.data.zero_()                                                       # :85-86
        out = net.forward(pts, z)                                                         # :87
        loss = torch.mean(torch.abs(out - sdf)) + 0.01 * torch.mean(torch.pow(z, 2))      # :88
        loss.backward(); nopt.step(); lopt.step()                                         # :89-91
        losses.append(loss.item())
    s['losses'] = np.array(losses, dtype=np.float64)
    s['latent_table_after'] = table.detach().numpy()
    after(s, 'after.', net)
    save('autodecoder_step', s)

    # ---------------------------------------------------------------- train_hybrid_progressive_gan.py:134-166, it=1
    it, fade, bb = 1, 0.4, 2
    r = pg.RESOLUTIONS[it]
    gen = sn.SDFNet(device='cpu')
    load_seeded(gen, 731)
    dis = pg.Discriminator()
    load_seeded(dis, 732)
    dis.set_iteration(it)
    dis.fade_in_progress = fade
    gopt = torch.optim.RMSprop(gen.parameters(), lr=0.0001)                               # :81
    dopt = torch.optim.RMSprop(dis.parameters(), lr=0.0001)                               # :82
    grid = util.get_voxel_coordinates(r, return_torch_tensor=True).cpu()                  # :95
    pts_b = grid.repeat((bb, 1))                                                          # :96
    z1, z2 = rnd((bb, 128), 733, -1, 1), rnd((bb, 128), 734, -1, 1)
    valid = synth_voxels(bb, r, 735) * 0.1
    alpha = rnd((bb, 1, 1, 1), 736, 0, 1)
    s = {'z_dis': z1.numpy(), 'z_gen': z2.numpy(), 'valid': valid.numpy(), 'alpha': alpha.numpy(), 'iteration': np.int64(it), 'fade': np.float64(fade)}

    def latent(z):
        return z.repeat((1, 1, grid.shape[0])).reshape(-1, 128)                            # :92
    # discriminator update :153-166
    dopt.zero_grad()                                                                      # :153
    fake = gen(pts_b, latent(z1)).reshape(-1, r, r, r)                                    # :154-156
    of = dis(fake)                                                                        # :157
    ov = dis(valid)                                                                       # :160
    a = alpha.expand(valid.shape)                                                         # :103 (device RNG in the reference: injected)
    xi = a * valid.detach() + ((1 - a) * fake.detach())                                   # :105
    xi.requires_grad = True                                                               # :106
    o = dis(xi)                                                                           # :108
    grads = torch.autograd.grad(outputs=o, inputs=xi, grad_outputs=torch.ones(o.shape), create_graph=True, retain_graph=True,
                                only_inputs=True)[0]                                      # :110
    gp = ((grads.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * 10                            # :111
    dl = of.mean() - ov.mean() + gp                                                       # :163
    dl.backward(); dopt.step()                                                            # :164-166
    # generator update :136-146
    gopt.zero_grad()                                                                      # :137
    fake = gen(pts_b, latent(z2)).reshape(-1, r, r, r)                                    # :139-141
    gl = -dis(fake).mean()                                                                # :147-148
    gl.backward(); gopt.step()                                                            # :149-150
    s['losses'] = np.array([dl.item(), gp.item(), gl.item()], dtype=np.float64)
    after(s, 'gen_after.', gen)
    after(s, 'dis_after.', dis)
    save('hybrid_step_it1', s)

    # ---------------------------------------------------------------- train_wgan.py:62-84 at B=64 (BASELINE configs[1]), digests
    bq = 64
    gen, cri = gan.Generator(), gan.Discriminator()
    load_seeded(gen, 601)
    load_seeded(cri, 602)
    cri.use_sigmoid = False                                                               # train_wgan.py:31
    gopt = torch.optim.RMSprop(gen.parameters(), lr=0.00005)
    copt = torch.optim.RMSprop(cri.parameters(), lr=0.00005)
    z1, z2, batch = rnd((bq, 128), 743, -2, 2), rnd((bq, 128), 744, -2, 2), synth_voxels(bq, 32, 745)
    s = {'seed_z_critic': np.int64(743), 'seed_z_gen': np.int64(744), 'seed_batch': np.int64(745), 'batch_size': np.int64(bq)}
    gen.zero_grad(); cri.zero_grad()
    fake = gen(z1).detach()
    closs = torch.mean(cri(fake)) - torch.mean(cri(batch))
    closs.backward()
    for k, p in cri.named_parameters():
        put(s, 'critic_grad.' + k, p.grad)
    copt.step(); cri.clip_weights(0.01)
    gen.zero_grad(); cri.zero_grad()
    gloss = -torch.mean(cri(gen(z2)))
    gloss.backward()
    for k, p in gen.named_parameters():
        put(s, 'gen_grad.' + k, p.grad)
    gopt.step()
    s['critic_loss'], s['generator_loss'] = np.float64(closs.item()), np.float64(gloss.item())
    after(s, 'gen_after.', gen)
    after(s, 'critic_after.', cri)
    save('wgan_step_b64', s)


if __name__ == '__main__':
    main()
