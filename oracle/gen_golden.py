"""ORACLE fixture generator — test infrastructure, NOT product code.

Imports the UNMODIFIED reference modules from /root/reference (read-only) in this
container, feeds them seeded weights/inputs and writes golden input/output
vectors to tests/golden/*.npz.  The reference has no tests or golden vectors of
its own, so these fixtures — "reference modules x torch CPU fp32" —
are the pin for oracle/ref_torch.py and for the CUDA path.

Run (from anywhere):  python oracle/gen_golden.py
Cannot run on the GPU box (/root/reference is absent there); the outputs are committed.

Shims (reference files untouched):
  * sys.modules stubs for trimesh / skimage (imported at model/sdf_net.py:2-3, unused on this path)
  * nn.Module.cuda -> no-op (model/gan.py:25,59 and model/autoencoder.py:65 hard-code self.cuda())
  * scratch cwd (util.py:11-13 creates plots/ models/ data/ on import)
"""
import hashlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
OUT = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, REPO)

from oracle import ref_torch as R  # noqa: E402


def import_reference():
    for name in ('trimesh', 'skimage', 'skimage.measure'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['skimage'].measure = sys.modules['skimage.measure']
    torch.nn.Module.cuda = lambda self, device=None: self
    scratch = tempfile.mkdtemp(prefix='sg_ref_')
    os.chdir(scratch)
    sys.path.insert(0, REF)
    import model.gan as gan
    import model.progressive_gan as pg
    import model.autoencoder as ae
    import model.sdf_net as sn
    import util
    return gan, pg, ae, sn, util


def digest(t, max_len=4096):
    """Compact fingerprint of a tensor: strided subsample + norms."""
    f = t.detach().double().flatten()
    stride = max(1, (f.numel() + max_len - 1) // max_len)
    return {
        'sub': f[::stride].float().numpy(),
        'stride': np.int64(stride),
        'sum': np.float64(f.sum().item()),
        'l2': np.float64(f.norm().item()),
        'shape': np.array(t.shape, dtype=np.int64),
    }


def put(store, key, t, full=False):
    if full:
        store[key] = t.detach().cpu().numpy()
    else:
        for k, v in digest(t).items():
            store['%s@%s' % (key, k)] = v


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def load_seeded(module, seed):
    sd = R.seeded_state_dict(shapes_of(module), seed)
    module.load_state_dict(sd, strict=True)
    return sd


def rnd(shape, seed, lo=-1.0, hi=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(shape, generator=g) * (hi - lo) + lo


def synth_voxels(b, r, seed):
    """clamp(randn*0.05, -0.1, 0.1)/0.1, mimicking datasets.py:20-22."""
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(torch.randn((b, r, r, r), generator=g) * 0.05, -0.1, 0.1) / 0.1


def save(name, store):
    store['torch_version'] = np.array(torch.__version__)
    store['threads'] = np.int64(torch.get_num_threads())
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **store)
    print('wrote', path, '%.1f KB' % (os.path.getsize(path) / 1024))


def main():
    torch.manual_seed(0)
    gan, pg, ae, sn, util = import_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---------------------------------------------------------------- grid coordinates (bit-exact)
    s = {}
    for r in (8, 16):
        s['coords_%d' % r] = util.get_voxel_coordinates(r)
    for r in (32, 64):
        s['coords_%d_sha256' % r] = np.array(hashlib.sha256(util.get_voxel_coordinates(r).tobytes()).hexdigest())
    save('voxel_coordinates', s)

    # ---------------------------------------------------------------- SDFNet fwd + autodecoder loss/backward
    net = sn.SDFNet(device='cpu')
    load_seeded(net, 101)
    n, shapes = 1000, 4
    pts = rnd((n, 3), 102)
    table = (rnd((shapes, 128), 103) * 0.5).requires_grad_(True)
    idx = (torch.arange(n) * shapes) // n
    target = torch.clamp(pts.norm(dim=1) - 0.5, -0.1, 0.1)
    s = {'seed_weights': np.int64(101), 'points': pts.numpy(), 'latent_table': table.detach().numpy(),
         'shape_index': idx.numpy(), 'target': target.numpy()}
    pts_g = pts.clone().requires_grad_(True)
    z = table[idx, :]
    out = net(pts_g, z)
    put(s, 'out', out, full=True)
    loss = torch.mean(torch.abs(out - target)) + 0.01 * torch.mean(torch.pow(z, 2))      # train_sdf_autodecoder.py:88
    loss.backward()
    s['loss'] = np.float64(loss.item())
    put(s, 'grad_latent_table', table.grad, full=True)
    put(s, 'grad_points', pts_g.grad, full=True)
    for k, p in net.named_parameters():
        put(s, 'grad.' + k, p.grad)
    # N=1 -> 0-d output; N=0 -> empty (sdf_net.py:61,73-74); latent_code_size=0 (demo_training.py:30)
    s['out_n1_shape'] = np.array(net(pts[:1], z[:1].detach()).shape, dtype=np.int64)
    s['out_n0_shape'] = np.array(net(pts[:0], z[:0].detach()).shape, dtype=np.int64)
    save('sdfnet_seeded', s)

    net0 = sn.SDFNet(latent_code_size=0, device='cpu')
    load_seeded(net0, 111)
    s = {'seed_weights': np.int64(111), 'points': pts.numpy()}
    put(s, 'out', net0(pts, torch.zeros((n, 0))), full=True)
    save('sdfnet_latent0', s)

    # ---------------------------------------------------------------- SDFNet chairs checkpoint on the 32^3 grid (examples/*.to)
    ck = torch.load(os.path.join(REF, 'examples', 'gan_generator_voxels_chairs.to'), weights_only=True, map_location='cpu')
    net.load_state_dict(ck, strict=True)
    g = torch.Generator().manual_seed(1234)
    zc = torch.randn((128,), generator=g)
    grid = util.get_voxel_coordinates(32, return_torch_tensor=True).cpu()
    with torch.no_grad():
        out = net(grid, zc.repeat(grid.shape[0], 1))
        ev = net.evaluate_in_batches(grid, zc, batch_size=10000)                        # sdf_net.py:63-75
    assert torch.allclose(out, ev, atol=1e-6)      # batched evaluation differs only by GEMM blocking
    s = {'z': zc.numpy()}
    for k, v in ck.items():
        s['w.' + k] = v.numpy()
    put(s, 'out', out, full=True)
    save('sdfnet_chairs', s)

    # ---------------------------------------------------------------- gan.Generator
    gen = gan.Generator()
    load_seeded(gen, 201)
    b = 4
    zg = rnd((b, 128), 202, -2, 2)
    wout = rnd((b, 1, 32, 32, 32), 203)
    s = {'seed_weights': np.int64(201), 'z': zg.numpy()}
    gen.train()
    out = gen(zg)
    put(s, 'out_train', out, full=True)
    (out * wout).sum().backward()
    s['seed_wout'] = np.int64(203)
    for k, p in gen.named_parameters():
        put(s, 'grad.' + k, p.grad)
    for k, v in gen.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            put(s, 'after.' + k, v, full=True)
    gen.eval()
    with torch.no_grad():
        put(s, 'out_eval', gen(zg), full=True)
    save('gan_generator', s)

    # ---------------------------------------------------------------- gan.Discriminator
    dis = gan.Discriminator()
    load_seeded(dis, 301)
    real = synth_voxels(b, 32, 302)
    fake = torch.tanh(rnd((b, 1, 32, 32, 32), 303, -1.5, 1.5))
    s = {'seed_weights': np.int64(301), 'real': real.numpy(), 'fake': fake.numpy()}
    dis.use_sigmoid = True
    with torch.no_grad():
        put(s, 'out_sigmoid', dis(real), full=True)
        s['out_b1_shape'] = np.array(dis(real[:1]).shape, dtype=np.int64)
    dis.use_sigmoid = False
    fake_g = fake.clone().requires_grad_(True)
    of, orl = dis(fake_g), dis(real)
    put(s, 'out_fake', of, full=True)
    put(s, 'out_real', orl, full=True)
    (torch.mean(of) - torch.mean(orl)).backward()                                        # train_wgan.py:68
    for k, p in dis.named_parameters():
        put(s, 'grad.' + k, p.grad)
    put(s, 'grad_fake', fake_g.grad)
    # BCE losses of train_gan.py:64,78,84
    dis.zero_grad()
    dis.use_sigmoid = True
    o = dis(fake)
    lg = -torch.mean(torch.log(o))
    lf = torch.nn.functional.binary_cross_entropy(o, torch.zeros(b))
    lv = torch.nn.functional.binary_cross_entropy(dis(real), torch.ones(b))
    s['bce_gen_loss'], s['bce_fake_loss'], s['bce_valid_loss'] = (np.float64(x.item()) for x in (lg, lf, lv))
    (lf + lv).backward()
    for k, p in dis.named_parameters():
        put(s, 'bce_grad.' + k, p.grad)
    # gradient penalty on gan.Discriminator (GP function of train_hybrid_progressive_gan.py:102-111)
    dis.zero_grad()
    dis.use_sigmoid = False
    alpha = rnd((b, 1, 1, 1), 304, 0, 1)
    s['alpha'] = alpha.numpy()
    fk = fake.squeeze(1)
    a = alpha.expand(real.shape)
    xi = (a * real + (1 - a) * fk).requires_grad_(True)
    o = dis(xi)
    grads = torch.autograd.grad(outputs=o, inputs=xi, grad_outputs=torch.ones(o.shape), create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
    gp = ((grads.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * 10
    s['gp'] = np.float64(gp.item())
    put(s, 'gp_input_grad', grads)
    gp.backward()
    for k, p in dis.named_parameters():
        if p.grad is not None:
            put(s, 'gp_grad.' + k, p.grad)
    save('gan_discriminator', s)

    # ---------------------------------------------------------------- progressive discriminator, it = 0..3
    for it, bb, fade in ((0, 3, 1.0), (1, 3, 0.3), (2, 3, 0.6), (3, 2, 0.25), (2, 2, 1.0)):
        d = pg.Discriminator()
        load_seeded(d, 400 + it)
        d.set_iteration(it)
        d.fade_in_progress = fade
        r = pg.RESOLUTIONS[it]
        real = synth_voxels(bb, r, 410 + it)
        fake = torch.clamp(rnd((bb, r, r, r), 420 + it, -0.12, 0.12), -0.1, 0.1)
        alpha = rnd((bb, 1, 1, 1), 430 + it, 0, 1)
        s = {'seed_weights': np.int64(400 + it), 'iteration': np.int64(it), 'fade': np.float64(fade),
             'real': real.numpy(), 'fake': fake.numpy(), 'alpha': alpha.numpy()}
        fake_g = fake.clone().requires_grad_(True)
        of, orl = d(fake_g), d(real)
        put(s, 'out_fake', of, full=True)
        put(s, 'out_real', orl, full=True)
        a = alpha.expand(real.shape)
        xi = (a * real + (1 - a) * fake).detach().requires_grad_(True)
        o = d(xi)
        grads = torch.autograd.grad(outputs=o, inputs=xi, grad_outputs=torch.ones(o.shape), create_graph=True,
                                    retain_graph=True, only_inputs=True)[0]
        gp = ((grads.norm(2, dim=(1, 2, 3)) - 1) ** 2).mean() * 10
        s['gp'] = np.float64(gp.item())
        loss = of.mean() - orl.mean() + gp                                               # train_hybrid_progressive_gan.py:163
        loss.backward()
        seen = set()
        for k, p in d.named_parameters():
            if p.grad is not None and id(p) not in seen:
                seen.add(id(p))
                put(s, 'grad.' + k, p.grad)
        put(s, 'grad_fake', fake_g.grad)
        save('progressive_disc_it%d_f%03d' % (it, int(fade * 100)), s)

    # ---------------------------------------------------------------- Autoencoder (VAE + classic)
    for variational in (True, False):
        m = ae.Autoencoder(is_variational=variational)
        load_seeded(m, 500 + int(variational))
        x = synth_voxels(b, 32, 510)
        eps = torch.randn((b, 128), generator=torch.Generator().manual_seed(511))
        s = {'seed_weights': np.int64(500 + int(variational)), 'x': x.numpy(), 'eps': eps.numpy()}
        m.train()
        if variational:
            # inject eps: util.standard_normal_distribution is a module global read at autoencoder.py:79
            class _Fixed:
                def sample(self, shape):
                    return eps.reshape(shape)
            ae.standard_normal_distribution = _Fixed()
            out, mean, logvar = m(x)
            put(s, 'mean', mean, full=True)
            put(s, 'log_variance', logvar, full=True)
            kld = -0.5 * torch.sum(1 + logvar - mean.pow(2) - logvar.exp()) / mean.nelement()   # train_autoencoder.py:54-55
        else:
            out = m(x)
            kld = 0
        put(s, 'out_train', out, full=True)
        diff = out - x
        diff = torch.where(x < 0, diff * 32, diff)                                       # train_autoencoder.py:57-62
        rec = torch.mean(torch.abs(diff))
        loss = rec + kld
        s['loss'] = np.float64(loss.item())
        loss.backward()
        for k, p in m.named_parameters():
            put(s, 'grad.' + k, p.grad)
        for k, v in m.state_dict().items():
            if 'running' in k:
                put(s, 'after.' + k, v)
        m.eval()
        with torch.no_grad():
            o = m(x)
            put(s, 'out_eval', o[0] if variational else o, full=True)
        save('autoencoder_%s' % ('vae' if variational else 'classic'), s)

    # ---------------------------------------------------------------- one train_wgan.py step (RMSprop + clip), B=4
    gen = gan.Generator()
    load_seeded(gen, 601)
    cri = gan.Discriminator()
    load_seeded(cri, 602)
    cri.use_sigmoid = False
    gopt = torch.optim.RMSprop(gen.parameters(), lr=0.00005)                            # train_wgan.py:37,45-46
    copt = torch.optim.RMSprop(cri.parameters(), lr=0.00005)
    z1, z2 = rnd((b, 128), 603, -2, 2), rnd((b, 128), 604, -2, 2)
    batch = synth_voxels(b, 32, 605)
    s = {'z_critic': z1.numpy(), 'z_gen': z2.numpy(), 'batch': batch.numpy()}
    gen.zero_grad(); cri.zero_grad()                                                     # :62-63
    fake = gen(z1).detach()                                                              # :65
    closs = torch.mean(cri(fake)) - torch.mean(cri(batch))                               # :66-68
    closs.backward(); copt.step(); cri.clip_weights(0.01)                                # :69-71
    gen.zero_grad(); cri.zero_grad()                                                     # :75-76
    gloss = -torch.mean(cri(gen(z2)))                                                    # :78-82
    gloss.backward(); gopt.step()                                                        # :83-84
    s['critic_loss'], s['generator_loss'] = np.float64(closs.item()), np.float64(gloss.item())
    for k, v in gen.state_dict().items():
        put(s, 'gen_after.' + k, v)
    for k, v in cri.state_dict().items():
        put(s, 'critic_after.' + k, v)
    save('wgan_step', s)


if __name__ == '__main__':
    main()
