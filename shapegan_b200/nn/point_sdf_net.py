"""Point-set GAN models -- drop-in for model/point_sdf_net.py: `PointNet(out_channels)` (the critic of
train_point_gan.py) and `SDFGenerator(latent_channels, hidden_channels, num_layers, norm=True, dropout=0.0)`, same constructor
arguments, attribute names (`nn1`, `nn2`, `lins`, `norms`, `z_lin1`, `z_lin2`), state_dict keys and forward signatures.

Every Linear runs on the tcgen05 implicit-GEMM kernels (ops.DenseOp / rowdot), LayerNorm+ReLU, the per-shape latent add and the
max pooling on csrc/sg_pointnet.cu.  PointNet is piecewise linear (Linear + ReLU + max), so the gradient penalty of
train_point_gan.py:61-71 -- autograd.grad(create_graph=True) w.r.t. the interpolated distances, then backward -- composes out of the
same twice-differentiable Functions as the voxel critics.  The torch.nn modules below are parameter containers only."""
import torch
import torch.nn as nn
from torch.nn import LayerNorm, Linear, ReLU, Sequential

from .. import ops, point_ops
from . import _require_cuda

try:
    from torch_scatter import scatter_max
except ImportError:
    scatter_max = None


def _dense(in_f, out_f, tag):
    """nn.Linear(in_f, out_f) as a DenseOp whose K is physically padded to a multiple of 64 (the K chunk of sg_igemm and the column atom
    of sg_wgrad): the input rows carry zero columns"""
    k = (in_f + 63) // 64 * 64
    return ops.DenseOp(1, out_f, 0, in_f, 1, k, 0, 1, '%s_%d_%d' % (tag, in_f, out_f), c_valid=in_f)


class PointNet(torch.nn.Module):
    def __init__(self, out_channels):
        super(PointNet, self).__init__()
        self.nn1 = Sequential(Linear(4, 64), ReLU(), Linear(64, 128), ReLU(), Linear(128, 256), ReLU(), Linear(256, 512))
        self.nn2 = Sequential(Linear(512, 256), ReLU(), Linear(256, 128), ReLU(), Linear(128, out_channels))
        self.out_channels = out_channels
        dims1, dims2 = (4, 64, 128, 256, 512), (512, 256, 128)
        self._ops1 = [_dense(dims1[i], dims1[i + 1], 'pn1') for i in range(4)]
        self._ops2 = [_dense(dims2[i], dims2[i + 1], 'pn2') for i in range(2)]
        self._op_out = _dense(128, out_channels, 'pn_out') if out_channels % 8 == 0 else None

    def forward(self, pos, dist, batch=None):
        _require_cuda(pos, 'PointNet.forward')
        if batch is not None:
            raise NotImplementedError('PointNet.forward: ragged `batch` vectors (torch_scatter) are not supported; pass [B, N, 3] / [B, N] tensors')
        dist = dist.unsqueeze(-1) if dist.size(-1) != 1 else dist
        x = torch.cat([pos, dist], dim=-1)                                   # [..., N, 4]   (point_sdf_net.py:36)
        lead = x.shape[:-2]
        n = x.shape[-2]
        rows = x.reshape(-1, 4).float()
        h = ops.to_planes(rows, 64)                                          # [P, rows, 64], channels 4..63 zero
        l1 = self.nn1
        for i, op in enumerate(self._ops1):
            lin = l1[2 * i]
            h = ops.linear_layer(op, h, lin.weight, lin.bias, ops.ACT_RELU if i < 3 else ops.ACT_NONE)
        h = point_ops.segment_max(h, n)                                      # [P, B, 512]   (:40)
        l2 = self.nn2
        for i, op in enumerate(self._ops2):
            lin = l2[2 * i]
            h = ops.linear_layer(op, h, lin.weight, lin.bias, ops.ACT_RELU)
        last = l2[4]
        if self.out_channels == 1:
            y = ops.rowdot(h, last.weight, last.bias).reshape(-1, 1)
        elif self._op_out is not None:
            y = ops.from_planes(ops.linear_layer(self._op_out, h, last.weight, last.bias))
        else:
            raise NotImplementedError('PointNet: out_channels must be 1 or a multiple of 8')
        return y.reshape(tuple(lead) + (self.out_channels,))


class SDFGenerator(torch.nn.Module):
    def __init__(self, latent_channels, hidden_channels, num_layers, norm=True, dropout=0.0):
        super(SDFGenerator, self).__init__()
        assert num_layers % 2 == 0
        self.layers1 = None
        self.layers2 = None
        self.latent_channels = latent_channels
        self.hidden_channels = hidden_channels
        self.num_layers = num_layers
        self.norm = norm
        self.dropout = dropout
        in_channels, out_channels = 3, hidden_channels
        self.lins = torch.nn.ModuleList()
        self.norms = torch.nn.ModuleList()
        self._ops = []
        for i in range(num_layers):                                          # point_sdf_net.py:63-80
            self.lins.append(Linear(in_channels, out_channels))
            self.norms.append(LayerNorm(out_channels))
            self._ops.append(_dense(in_channels, out_channels, 'sg%d' % i) if out_channels % 8 == 0 else None)
            in_channels = hidden_channels + 3 if i == (num_layers // 2) - 1 else hidden_channels
            if i == num_layers - 2:
                out_channels = 1
        self.z_lin1 = Linear(latent_channels, hidden_channels)
        self.z_lin2 = Linear(latent_channels, hidden_channels)
        self._op_z = _dense(latent_channels, hidden_channels, 'sgz')

    def _z(self, lin, z):
        return ops.from_planes(ops.linear_layer(self._op_z, ops.to_planes(z.float(), (z.shape[1] + 63) // 64 * 64), lin.weight, lin.bias))   # fp32 [B, hidden]

    def forward(self, pos, z):
        _require_cuda(pos, 'SDFGenerator.forward')
        if self.dropout != 0.0 and self.training:
            raise NotImplementedError('SDFGenerator: dropout > 0 is not implemented (train_point_gan.py uses 0.0)')
        if not self.norm:
            raise NotImplementedError('SDFGenerator: norm=False is not implemented (train_point_gan.py uses NORM = True)')
        pos = pos.unsqueeze(0) if pos.dim() == 2 else pos
        assert pos.dim() == 3 and pos.size(-1) == 3
        z = z.unsqueeze(0) if z.dim() == 1 else z
        assert z.dim() == 2 and z.size(-1) == self.latent_channels and pos.size(0) == z.size(0)
        b, n = pos.shape[0], pos.shape[1]
        p8 = ops.to_planes(pos.reshape(-1, 3).float(), 64)                   # [P, B*N, 64], channels 3..63 zero
        x = p8
        half = self.num_layers // 2
        for i, (lin, norm) in enumerate(zip(self.lins, self.norms)):
            if i == half:
                x = torch.cat((x, p8), dim=2)                                # cat([x, pos]) (:101), zero-padded to hidden + 64 columns
            if i == self.num_layers - 1:
                y = ops.rowdot(x, lin.weight, lin.bias)                      # Linear(hidden -> 1), no norm / activation on the last layer
                return y.reshape(b, n, 1)
            x = ops.linear_layer(self._ops[i], x, lin.weight, lin.bias)
            if i == 0:
                x = point_ops.rows_add_vec(x, self._z(self.z_lin1, z), n)    # :105-106
            if i == half:
                x = point_ops.rows_add_vec(x, self._z(self.z_lin2, z), n)    # :108-109
            x = point_ops.layernorm_act(x, norm, ops.ACT_RELU)               # :111-113
        return x
