import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))


def check_digest(store, key, tensor, rtol, what='', atol=0.0, ntol=None):
    """Compare a tensor with a golden digest written by oracle/gen_golden.py:put().
    Metric: relative L2 over the strided subsample, and relative error of the full L2 norm.
    `atol` is an absolute max-abs escape for tensors that are analytically zero (e.g. the gradient
    of a conv bias that feeds a train-mode BatchNorm), where only rounding noise remains."""
    f = tensor.detach().double().flatten().cpu()
    stride = int(store[key + '@stride'])
    sub = torch.from_numpy(store[key + '@sub']).double()
    assert list(tensor.shape) == list(store[key + '@shape']), (what, key, tensor.shape)
    mine = f[::stride]
    if atol > 0 and (mine - sub).abs().max().item() <= atol:
        return 0.0
    denom = max(sub.norm().item(), 1e-30)
    err = (mine - sub).norm().item() / denom
    l2 = float(store[key + '@l2'])
    l2err = abs(f.norm().item() - l2) / max(l2, 1e-30)
    ntol = rtol if ntol is None else ntol
    assert err <= rtol and l2err <= ntol, '%s %s: rel-L2 %.3e, norm err %.3e > %.1e' % (what, key, err, l2err, rtol)
    return err


def rel_l2(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
