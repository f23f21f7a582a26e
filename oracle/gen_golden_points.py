"""ORACLE fixture generator — test infrastructure, NOT product code.

Goldens of the point-set GAN path from the UNMODIFIED reference module model/point_sdf_net.py (imported from /root/reference; its
optional torch_scatter import is already guarded there) driven through the step bodies of train_point_gan.py:52-87 with the
device-RNG draws (z, alpha) injected:  tests/golden/point_gan.npz.
Run:  python oracle/gen_golden_points.py   (outputs are committed)"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import gen_golden as G  # noqa: E402
from oracle.gen_golden import load_seeded, put, rnd, save  # noqa: E402


def main():
    torch.manual_seed(0)
    G.import_reference()
    import model.point_sdf_net as P
    b, n = 3, 256
    gen = P.SDFGenerator(128, 256, 8, True, dropout=0.0)                                   # train_point_gan.py:21
    dis = P.PointNet(out_channels=1)                                                       # :22
    load_seeded(gen, 801)
    load_seeded(dis, 802)
    pos = rnd((b, n, 3), 803)
    dist = (pos.norm(dim=-1, keepdim=True) - 0.5).clamp(-0.1, 0.1)                         # "uniform[..., 3:]"
    z1, z2 = rnd((b, 128), 804, -2, 2), rnd((b, 128), 805, -2, 2)
    alpha = rnd((b, 1, 1), 806, 0, 1)
    s = {'pos': pos.numpy(), 'dist': dist.numpy(), 'z_dis': z1.numpy(), 'z_gen': z2.numpy(), 'alpha': alpha.numpy(),
         'seed_gen': np.int64(801), 'seed_dis': np.int64(802)}
    # ---- critic step (:52-76)
    fake = gen(pos, z1)                                                                    # :57
    put(s, 'fake', fake, full=True)
    out_real = dis(pos, dist)                                                              # :58
    out_fake = dis(pos, fake)                                                              # :59
    put(s, 'out_real', out_real, full=True)
    put(s, 'out_fake', out_fake, full=True)
    d_loss = out_fake.mean() - out_real.mean()                                             # :60
    inter = alpha * dist + (1 - alpha) * fake                                              # :62-63
    inter = inter.detach().requires_grad_(True)                                            # :64 (on a leaf copy: the reference's in-place flag needs a leaf)
    out = dis(pos, inter)                                                                  # :65
    grad = torch.autograd.grad(out, inter, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True, only_inputs=True)[0]   # :67-70
    gn = grad.view(grad.size(0), -1).norm(dim=-1, p=2)                                     # :71
    gp = 10 * ((gn - 1).pow(2).mean())                                                     # :72
    put(s, 'gp_input_grad', grad, full=True)
    dis.zero_grad(); gen.zero_grad()
    (d_loss + gp).backward()                                                               # :74-75
    s['d_loss'], s['gp'] = np.float64(d_loss.item()), np.float64(gp.item())
    for k, p in dis.named_parameters():
        put(s, 'dis_grad.' + k, p.grad)
    # ---- generator step (:80-86)
    dis.zero_grad(); gen.zero_grad()
    fake = gen(pos, z2)
    loss = -dis(pos, fake).mean()
    loss.backward()
    s['g_loss'] = np.float64(loss.item())
    for k, p in gen.named_parameters():
        if p.grad is not None:                         # norms.7 is constructed but never used (point_sdf_net.py:110: i < num_layers - 1)
            put(s, 'gen_grad.' + k, p.grad)
    save('point_gan', s)


if __name__ == '__main__':
    main()
