"""ORACLE — test infrastructure, NOT product code.

CPU fp32 functional restatement of the point-set GAN models and losses of the reference:
  pointnet_forward      PointNet.forward        model/point_sdf_net.py:32-46 (batch=None branch: x.max(dim=-2)[0])
  sdf_generator_forward SDFGenerator.forward    model/point_sdf_net.py:87-116 (norm=True, dropout=0)
  critic_loss_with_gp / generator_loss          train_point_gan.py:52-78, :80-87 (alpha injected instead of torch.rand on the device)
Pinned against goldens produced by the unmodified reference module (oracle/gen_golden_points.py) in tests/test_oracle_golden.py."""
import torch
import torch.nn.functional as F


def pointnet_forward(sd, pos, dist):
    dist = dist.unsqueeze(-1) if dist.size(-1) != 1 else dist
    x = torch.cat([pos, dist], dim=-1)
    for i in (0, 2, 4, 6):
        x = F.linear(x, sd['nn1.%d.weight' % i], sd['nn1.%d.bias' % i])
        if i < 6:
            x = F.relu(x)
    x = x.max(dim=-2)[0]
    for i in (0, 2, 4):
        x = F.linear(x, sd['nn2.%d.weight' % i], sd['nn2.%d.bias' % i])
        if i < 4:
            x = F.relu(x)
    return x


def sdf_generator_forward(sd, pos, z, num_layers=8):
    pos = pos.unsqueeze(0) if pos.dim() == 2 else pos
    z = z.unsqueeze(0) if z.dim() == 1 else z
    x = pos
    for i in range(num_layers):
        if i == num_layers // 2:
            x = torch.cat([x, pos], dim=-1)
        x = F.linear(x, sd['lins.%d.weight' % i], sd['lins.%d.bias' % i])
        if i == 0:
            x = F.linear(z, sd['z_lin1.weight'], sd['z_lin1.bias']).unsqueeze(1) + x
        if i == num_layers // 2:
            x = F.linear(z, sd['z_lin2.weight'], sd['z_lin2.bias']).unsqueeze(1) + x
        if i < num_layers - 1:
            x = F.layer_norm(x, (x.shape[-1],), sd['norms.%d.weight' % i], sd['norms.%d.bias' % i], 1e-5)
            x = F.relu(x)
    return x


def critic_loss_with_gp(dsd, pos, dist, fake, alpha, weight=10.0):
    """train_point_gan.py:57-74: D_loss + gp (fake detached: only D is updated here)"""
    out_real = pointnet_forward(dsd, pos, dist)
    out_fake = pointnet_forward(dsd, pos, fake)
    d_loss = out_fake.mean() - out_real.mean()
    inter = (alpha * dist + (1 - alpha) * fake).detach().requires_grad_(True)
    out = pointnet_forward(dsd, pos, inter)
    grad = torch.autograd.grad(out, inter, grad_outputs=torch.ones_like(out), create_graph=True, retain_graph=True, only_inputs=True)[0]
    gn = grad.view(grad.size(0), -1).norm(dim=-1, p=2)
    gp = weight * ((gn - 1).pow(2).mean())
    return d_loss + gp, d_loss, gp
