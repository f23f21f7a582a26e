"""Drop-in module surface of the reference's `model` package (model/__init__.py:1-47): constants, `Lambda`,
`SavableModule` (checkpoint path layout + state_dict save/load + .device).  The classes in this package keep the
reference's class names, constructor arguments, mutable attributes, `state_dict()` keys/shapes/dtypes and
`.forward()` signatures; everything under `.forward()` runs on libsg_b200 (no aten conv/GEMM on the hot path)."""
import os

import torch
import torch.nn as nn
from torch.nn import BatchNorm1d, Linear, ReLU, Sequential  # noqa: F401  (re-exported by the reference's `from model import *`)

MODEL_PATH = "models"                                                   # model/__init__.py:7
CHECKPOINT_PATH = os.path.join(MODEL_PATH, 'checkpoints')               # :8
LATENT_CODES_FILENAME = os.path.join(MODEL_PATH, "sdf_net_latent_codes.to")   # :9
LATENT_CODE_SIZE = 128                                                  # :10


class Lambda(nn.Module):
    """model/__init__.py:12-18"""

    def __init__(self, function):
        super().__init__()
        self.function = function

    def forward(self, x):
        return self.function(x)


class SavableModule(nn.Module):
    """model/__init__.py:20-47: `models/<filename>` and `models/checkpoints/<stem>-epoch-%05d.<ext>`."""

    def __init__(self, filename):
        super().__init__()
        self.filename = filename

    def get_filename(self, epoch=None, filename=None):
        name = self.filename if filename is None else filename
        if epoch is None:
            return os.path.join(MODEL_PATH, name)
        parts = name.split('.')
        parts[-2] += '-epoch-{:05d}'.format(epoch)
        return os.path.join(CHECKPOINT_PATH, '.'.join(parts))

    def load(self, epoch=None):
        # strict=False like the reference (:38); map_location added so CPU-saved checkpoints load on the local GPU
        state = torch.load(self.get_filename(epoch=epoch), map_location=self.device)
        self.load_state_dict(state, strict=False)
        from ..ops import invalidate_weight_cache
        invalidate_weight_cache()

    def save(self, epoch=None):
        if epoch is not None and not os.path.exists(CHECKPOINT_PATH):
            os.makedirs(CHECKPOINT_PATH, exist_ok=True)
        torch.save(self.state_dict(), self.get_filename(epoch=epoch))

    @property
    def device(self):
        return next(self.parameters()).device

    def _to_default_device(self):
        """The reference hard-codes self.cuda() in constructors (model/gan.py:25,59; autoencoder.py:65).  Same here
        when a GPU is visible; on a CPU-only host the parameters stay on the CPU so checkpoints/state_dicts can be
        inspected — forward() then fails loudly (no CPU fallback)."""
        if torch.cuda.is_available():
            self.cuda()


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError('%s: shapegan_b200 runs on CUDA (sm_100a) only; got a %s tensor — there is no CPU fallback'
                           % (who, t.device))
