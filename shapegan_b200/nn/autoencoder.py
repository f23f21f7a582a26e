"""Voxel (variational) autoencoder — drop-in for model/autoencoder.py (state_dict keys `encoder.{0,1,3,4,6,7,9,10,13}`,
`encoder.vae-bn`, `encode_mean`, `encode_log_variance`, `decoder.{0,1,4,5,7,8,10,11,13}`; encode/decode/forward)."""
import torch
import torch.nn as nn

from .. import ops
from . import LATENT_CODE_SIZE, Lambda, SavableModule, _require_cuda

standard_normal_distribution = torch.distributions.normal.Normal(0, 1)      # CPU RNG (util.py:3); eps drawn at autoencoder.py:79

AUTOENCODER_MODEL_COMPLEXITY_MULTIPLIER = 24
amcm = AUTOENCODER_MODEL_COMPLEXITY_MULTIPLIER
_ENC = (1, 1 * amcm, 2 * amcm, 4 * amcm, LATENT_CODE_SIZE * 2)              # model/autoencoder.py:16-28
_DEC = (LATENT_CODE_SIZE * 2, 4 * amcm, 2 * amcm, 1 * amcm, 1)              # :51-63


class Autoencoder(SavableModule):
    def __init__(self, is_variational=True):
        super().__init__(filename="autoencoder-{:d}.to".format(LATENT_CODE_SIZE))
        self.is_variational = is_variational
        if is_variational:
            self.filename = 'variational-' + self.filename
        e, d = _ENC, _DEC
        enc = []
        for i in range(3):
            enc += [nn.Conv3d(e[i], e[i + 1], kernel_size=4, stride=2, padding=1), nn.BatchNorm3d(e[i + 1]),
                    nn.LeakyReLU(negative_slope=0.2, inplace=True)]
        enc += [nn.Conv3d(e[3], e[4], kernel_size=4, stride=1), nn.BatchNorm3d(e[4]), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                Lambda(lambda x: x), nn.Linear(e[4], LATENT_CODE_SIZE)]
        self.encoder = nn.Sequential(*enc)
        if is_variational:
            self.encoder.add_module('vae-bn', nn.BatchNorm1d(LATENT_CODE_SIZE))
            self.encoder.add_module('vae-lr', nn.LeakyReLU(negative_slope=0.2, inplace=True))
            self.encode_mean = nn.Linear(LATENT_CODE_SIZE, LATENT_CODE_SIZE)
            self.encode_log_variance = nn.Linear(LATENT_CODE_SIZE, LATENT_CODE_SIZE)
        dec = [nn.Linear(LATENT_CODE_SIZE, d[0]), nn.BatchNorm1d(d[0]), nn.LeakyReLU(negative_slope=0.2, inplace=True),
               Lambda(lambda x: x), nn.ConvTranspose3d(d[0], d[1], kernel_size=4, stride=1), nn.BatchNorm3d(d[1]),
               nn.LeakyReLU(negative_slope=0.2, inplace=True)]
        for i in (1, 2):
            dec += [nn.ConvTranspose3d(d[i], d[i + 1], kernel_size=4, stride=2, padding=1), nn.BatchNorm3d(d[i + 1]),
                    nn.LeakyReLU(negative_slope=0.2, inplace=True)]
        dec += [nn.ConvTranspose3d(d[3], d[4], kernel_size=4, stride=2, padding=1)]
        self.decoder = nn.Sequential(*dec)

        self._e0 = ops.Conv1Op(e[1])
        self._e1 = ops.ConvOp(e[1], e[2])
        self._e2 = ops.ConvOp(e[2], e[3])
        # Conv3d(96->256,k4,s1) on the 4^3 grid: dense over (position, channel)
        self._e3 = ops.DenseOp(1, e[4], 0, e[3] * 64, 64, e[3], 1, 64, 'ae_e3')
        self._e4 = ops.linear_op(e[4], LATENT_CODE_SIZE, 'ae_e4')
        self._em = ops.linear_op(LATENT_CODE_SIZE, LATENT_CODE_SIZE, 'ae_m')
        self._d0 = ops.linear_op(LATENT_CODE_SIZE, d[0], 'ae_d0')
        # ConvTranspose3d(256->96,k4,s1) on a 1^3 grid: GEMM onto (position, channel)
        self._d1 = ops.DenseOp(64, d[1], 1, 64, 1, d[0], 0, d[1] * 64, 'ae_d1')
        self._d2 = ops.ConvTOp(d[1], d[2])
        self._d3 = ops.ConvTOp(d[2], d[3])
        self._d4 = ops.ConvT1Op(d[3])
        self._to_default_device()

    def encode(self, x, return_mean_and_log_variance=False):
        x = x.reshape((-1, 32, 32, 32)).float()
        _require_cuda(x, 'Autoencoder.encode')
        enc, e = self.encoder, _ENC
        b = x.shape[0]
        h = ops.linear_layer(self._e0, x, enc[0].weight, enc[0].bias)
        h = ops.batchnorm_act(h, enc[1], ops.ACT_LRELU, e[1])
        h = ops.linear_layer(self._e1, h, enc[3].weight, enc[3].bias)
        h = ops.batchnorm_act(h, enc[4], ops.ACT_LRELU, e[2])
        h = ops.linear_layer(self._e2, h, enc[6].weight, enc[6].bias)
        h = ops.batchnorm_act(h, enc[7], ops.ACT_LRELU, e[3])
        h = h.reshape(h.shape[0], b, 64 * e[3])
        h = ops.linear_layer(self._e3, h, enc[9].weight, enc[9].bias)
        h = ops.batchnorm_act(h, enc[10], ops.ACT_LRELU, e[4])
        h = ops.linear_layer(self._e4, h, enc[13].weight, enc[13].bias)
        if not self.is_variational:
            return ops.from_planes(h)
        h = ops.batchnorm_act(h, getattr(enc, 'vae-bn'), ops.ACT_LRELU, LATENT_CODE_SIZE)
        h, h2 = ops.fanout2(h)
        mean = ops.from_planes(ops.linear_layer(self._em, h, self.encode_mean.weight, self.encode_mean.bias)).squeeze()
        if self.training or return_mean_and_log_variance:
            log_variance = ops.from_planes(ops.linear_layer(self._em, h2, self.encode_log_variance.weight,
                                                            self.encode_log_variance.bias)).squeeze()
            # reparameterisation on [B,128] vectors (autoencoder.py:78-82); eps from the CPU generator like the reference
            standard_deviation = torch.exp(log_variance * 0.5)
            eps = standard_normal_distribution.sample(mean.shape).to(x.device)
        z = mean + standard_deviation * eps if self.training else mean
        if return_mean_and_log_variance:
            return z, mean, log_variance
        return z

    def decode(self, x):
        if len(x.shape) == 1:
            x = x.unsqueeze(dim=0)
        _require_cuda(x, 'Autoencoder.decode')
        dec, d = self.decoder, _DEC
        b = x.shape[0]
        h = ops.to_planes(x.float())
        h = ops.linear_layer(self._d0, h, dec[0].weight, dec[0].bias)
        h = ops.batchnorm_act(h, dec[1], ops.ACT_LRELU, d[0])
        h = ops.linear_layer(self._d1, h, dec[4].weight, dec[4].bias)
        h = h.reshape(h.shape[0], b, 4, 4, 4, d[1])
        h = ops.batchnorm_act(h, dec[5], ops.ACT_LRELU, d[1])
        h = ops.linear_layer(self._d2, h, dec[7].weight, dec[7].bias)
        h = ops.batchnorm_act(h, dec[8], ops.ACT_LRELU, d[2])
        h = ops.linear_layer(self._d3, h, dec[10].weight, dec[10].bias)
        h = ops.batchnorm_act(h, dec[11], ops.ACT_LRELU, d[3])
        out = ops.convt1_act(self._d4, h, dec[13].weight, dec[13].bias, ops.ACT_NONE)
        return out.squeeze()

    def forward(self, x):
        if not self.is_variational:
            return self.decode(self.encode(x))
        z, mean, log_variance = self.encode(x, return_mean_and_log_variance=True)
        return self.decode(z), mean, log_variance
