"""Progressive discriminator — drop-in for model/progressive_gan.py (RESOLUTIONS, FEATURE_COUNTS, from_SDF,
Discriminator with .iteration/.set_iteration/.fade_in_progress/.filename_base and the doubled
`optional_layers.{i}.0.*` / `optional_layer_{i}.0.*` state_dict entries)."""
import torch
import torch.nn as nn

from .. import ops
from . import LATENT_CODE_SIZE, Lambda, SavableModule, _require_cuda  # noqa: F401  (LATENT_CODE_SIZE re-exported, train_hybrid_progressive_gan.py:16)

RESOLUTIONS = [8, 16, 32, 64]              # model/progressive_gan.py:4
FEATURE_COUNTS = [128, 64, 32, 1]          # :5
FINAL_LAYER_FEATURES = 256                 # :6


def from_SDF(x, iteration):
    """model/progressive_gan.py:9-16 (kept for API parity; the kernels never materialise the zero channels)."""
    resolution = RESOLUTIONS[iteration]
    x = x.reshape((-1, 1, resolution, resolution, resolution))
    pad = torch.zeros((x.shape[0], FEATURE_COUNTS[iteration] - 1, resolution, resolution, resolution), device=x.device)
    return torch.cat((x, pad), dim=1)


def _block_channels(i):
    return FEATURE_COUNTS[i], (FEATURE_COUNTS[i - 1] if i > 0 else FINAL_LAYER_FEATURES)


class Discriminator(SavableModule):
    def __init__(self):
        self.iteration = 0
        self.filename_base = "hybrid_progressive_gan_discriminator_{:d}.to"
        super().__init__(filename=self.filename_base.format(self.iteration))
        self.fade_in_progress = 1
        self.head = nn.Sequential(
            Lambda(lambda x: x),                                    # reshape happens in forward (NDHWC planes)
            nn.Linear(64 * FINAL_LAYER_FEATURES, 128),
            nn.LeakyReLU(negative_slope=0.2),
            nn.Linear(128, 1))
        self.optional_layers = nn.ModuleList()
        self._ops_first, self._ops_inner = [], []
        for i in range(len(FEATURE_COUNTS)):
            cin, cout = _block_channels(i)
            block = nn.Sequential(nn.Conv3d(cin, cout, kernel_size=4, stride=2, padding=1), nn.LeakyReLU(negative_slope=0.2))
            self.optional_layers.append(block)
            self.add_module('optional_layer_{:d}'.format(i), block)          # alias, progressive_gan.py:41-42
            # block i as FIRST layer sees from_SDF input: only channel 0 of its weight meets non-zero data
            self._ops_first.append(ops.Conv1Op(cout, w_cin=cin))
            self._ops_inner.append(ops.ConvOp(cin, cout) if cin % 8 == 0 else None)
        # Linear(16384 -> 128) on the NCDHW flatten (index c*64 + pos) of the NDHWC activation (row = pos*256 + c)
        self._op_head = ops.DenseOp(1, 128, 0, 64 * FINAL_LAYER_FEATURES, 64, FINAL_LAYER_FEATURES, 1, 64, 'pdhead')

    def forward(self, x):
        it = self.iteration
        r = RESOLUTIONS[it]
        _require_cuda(x, 'progressive_gan.Discriminator.forward')
        x_in = x.reshape((-1, r, r, r)).float()
        b = x_in.shape[0]
        conv = self.optional_layers[it][0]
        h = ops.linear_layer(self._ops_first[it], x_in, conv.weight, conv.bias, ops.ACT_LRELU)
        if (self.fade_in_progress < 1.0) and it > 0:
            h = ops.fade(h, x_in, float(self.fade_in_progress))               # progressive_gan.py:48-50
        i = it - 1
        while i >= 0:
            conv = self.optional_layers[i][0]
            h = ops.linear_layer(self._ops_inner[i], h, conv.weight, conv.bias, ops.ACT_LRELU)
            i -= 1
        h = h.reshape(h.shape[0], b, 64 * FINAL_LAYER_FEATURES)
        h = ops.linear_layer(self._op_head, h, self.head[1].weight, self.head[1].bias, ops.ACT_LRELU)
        y = ops.rowdot(h, self.head[3].weight, self.head[3].bias)
        return y.squeeze()

    def set_iteration(self, value):
        self.iteration = value
        self.filename = self.filename_base.format(self.iteration)
