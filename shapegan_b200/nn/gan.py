"""Voxel GAN generator / discriminator — drop-in for model/gan.py (class names, ctor, attributes, state_dict keys
`layers.{0,1,3,4,6,7,9}.*` / `layers.{0,2,4,6}.*`, forward signatures), with the whole forward/backward on libsg_b200.

The torch.nn layer objects below are parameter CONTAINERS only (they give the reference's state_dict layout and
default initialisation); their forward() is never called."""
import torch
import torch.nn as nn

from .. import ops, raw
from . import LATENT_CODE_SIZE, Lambda, SavableModule, _require_cuda

standard_normal_distribution = torch.distributions.normal.Normal(0, 1)      # CPU RNG, as util.py:3

_G_CHANNELS = (LATENT_CODE_SIZE, 256, 128, 64, 1)       # model/gan.py:9-21
_D_CHANNELS = (1, 64, 128, 256, 1)                      # model/gan.py:49-55


def _stack(kinds):
    return nn.Sequential(*[k() for k in kinds])


class Generator(SavableModule):
    """z [B,128] -> SDF voxels [B,1,32,32,32]  (model/gan.py:4-34)."""

    def __init__(self):
        super().__init__(filename="generator.to")
        c = _G_CHANNELS
        mods = [nn.ConvTranspose3d(c[0], c[1], kernel_size=4, stride=1), nn.BatchNorm3d(c[1]), nn.LeakyReLU(0.2)]
        for i in (1, 2):
            mods += [nn.ConvTranspose3d(c[i], c[i + 1], kernel_size=4, stride=2, padding=1), nn.BatchNorm3d(c[i + 1]),
                     nn.LeakyReLU(0.2)]
        mods += [nn.ConvTranspose3d(c[3], c[4], kernel_size=4, stride=2, padding=1), nn.Tanh()]
        self.layers = nn.Sequential(*mods)
        # ConvTranspose3d(128->256,k4,s1) on a 1^3 grid is a GEMM [B,128] x [128, 64 positions x 256]
        self._op0 = ops.DenseOp(64, c[1], 1, 64, 1, c[0], 0, c[1] * 64, 'g0')
        self._op1 = ops.ConvTOp(c[1], c[2])
        self._op2 = ops.ConvTOp(c[2], c[3])
        self._op3 = ops.ConvT1Op(c[3])
        self._to_default_device()

    def forward(self, x):
        x = x.reshape((-1, LATENT_CODE_SIZE))
        _require_cuda(x, 'Generator.forward')
        l, c = self.layers, _G_CHANNELS
        b = x.shape[0]
        h = ops.to_planes(x.float())
        h = ops.linear_layer(self._op0, h, l[0].weight, l[0].bias)
        h = h.reshape(h.shape[0], b, 4, 4, 4, c[1])
        h = ops.batchnorm_act(h, l[1], ops.ACT_LRELU, c[1])
        h = ops.linear_layer(self._op1, h, l[3].weight, l[3].bias)
        h = ops.batchnorm_act(h, l[4], ops.ACT_LRELU, c[2])
        h = ops.linear_layer(self._op2, h, l[6].weight, l[6].bias)
        h = ops.batchnorm_act(h, l[7], ops.ACT_LRELU, c[3])
        out = ops.convt1_act(self._op3, h, l[9].weight, l[9].bias, ops.ACT_TANH)
        return out.unsqueeze(1)

    def generate(self, sample_size=1):
        z = standard_normal_distribution.sample(torch.Size((sample_size, LATENT_CODE_SIZE))).to(self.device)
        return self(z)

    def copy_autoencoder_weights(self, autoencoder):
        raise Exception("Not implemented.")          # as the reference (model/gan.py:36-40)


class Discriminator(SavableModule):
    """voxels [B,32,32,32] or [B,1,32,32,32] -> score [B] (0-d for B == 1)  (model/gan.py:43-69)."""

    def __init__(self):
        super().__init__(filename="discriminator.to")
        self.use_sigmoid = True
        c = _D_CHANNELS
        mods = []
        for i in range(3):
            mods += [nn.Conv3d(c[i], c[i + 1], kernel_size=4, stride=2, padding=1), nn.LeakyReLU(0.2)]
        mods += [nn.Conv3d(c[3], c[4], kernel_size=4, stride=1),
                 Lambda(lambda t: ops.unary_f32(t, ops.ACT_SIGMOID) if self.use_sigmoid else t)]
        self.layers = nn.Sequential(*mods)
        self._op0 = ops.Conv1Op(c[1])
        self._op1 = ops.ConvOp(c[1], c[2])
        self._op2 = ops.ConvOp(c[2], c[3])
        self._to_default_device()

    def forward(self, x):
        if len(x.shape) < 5:
            x = x.unsqueeze(dim=1)
        _require_cuda(x, 'Discriminator.forward')
        l = self.layers
        b = x.shape[0]
        vol = x.reshape(b, x.shape[2], x.shape[3], x.shape[4]).float()
        h = ops.linear_layer(self._op0, vol, l[0].weight, l[0].bias, ops.ACT_LRELU)
        h = ops.linear_layer(self._op1, h, l[2].weight, l[2].bias, ops.ACT_LRELU)
        h = ops.linear_layer(self._op2, h, l[4].weight, l[4].bias, ops.ACT_LRELU)
        # Conv3d(256->1,k4,s1) on the 4^3 grid is one dot product per sample; weight [1,256,4,4,4] read in place
        h = h.reshape(h.shape[0], b, 64 * _D_CHANNELS[3])
        y = ops.rowdot(h, l[6].weight, l[6].bias, ops.ACT_NONE, _D_CHANNELS[3], 1, 64)
        y = l[7](y.reshape(b, 1, 1, 1, 1))
        return y.squeeze()

    def clip_weights(self, value):
        for parameter in self.parameters():
            raw.clamp_(parameter.data, -value, value)
        ops.invalidate_weight_cache()
