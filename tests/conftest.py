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


# ------------------------------------------------------------------------------------------------- parity bookkeeping
import json  # noqa: E402


def digest_errors(store, key, tensor):
    """(rel-L2 over the golden's strided subsample, relative error of the full L2 norm, max |diff| over the subsample)"""
    f = tensor.detach().double().flatten().cpu()
    stride = int(store[key + '@stride'])
    sub = torch.from_numpy(store[key + '@sub']).double()
    assert list(tensor.shape) == list(store[key + '@shape']), (key, tensor.shape)
    mine = f[::stride]
    err = (mine - sub).norm().item() / max(sub.norm().item(), 1e-30)
    l2 = float(store[key + '@l2'])
    return err, abs(f.norm().item() - l2) / max(l2, 1e-30), (mine - sub).abs().max().item()


def _load_json(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


SELF_NOISE = _load_json('fp32_self_noise.json')          # oracle/gen_noise.py: the reference's own fp32-vs-fp64 error per tensor
MEASURED = _load_json('parity_measured.json')            # tools/update_parity_gates.py: this library's measured error per tensor


class ParityRecorder:
    """Every per-tensor parity error the GPU tests measure, with the gate it was held to; written to gpurun_out/ at session end
    (tools/update_parity_gates.py turns a run into tests/golden/parity_measured.json, profiles/ keeps the readable table)."""

    def __init__(self):
        self.rows = []

    def add(self, prec, case, key, err, gate, norm_err=None, note=''):
        self.rows.append({'prec': prec, 'case': case, 'key': key, 'err': err, 'gate': gate, 'norm_err': norm_err, 'note': note})

    def dump(self):
        if not self.rows:
            return
        out = os.path.join(REPO, 'gpurun_out')
        try:
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, 'parity_errors.json'), 'w') as f:
                json.dump(self.rows, f, indent=0)
            with open(os.path.join(out, 'parity_errors.txt'), 'w') as f:
                f.write('# per-tensor parity errors (rel-L2 vs the goldens of the unmodified reference) measured by tests/test_parity_gpu.py\n')
                f.write('# %-6s %-30s %-44s %10s %10s %10s  %s\n' % ('prec', 'case', 'tensor', 'err', 'gate', 'ref-noise', 'note'))
                for r in self.rows:
                    noise = SELF_NOISE.get(r['case'], {}).get(r['key'])
                    f.write('  %-6s %-30s %-44s %10.2e %10.2e %10s  %s\n' % (r['prec'], r['case'], r['key'], r['err'], r['gate'],
                                                                            ('%.1e' % noise) if noise is not None else '-', r['note']))
        except OSError:
            pass


PARITY = ParityRecorder()


def pytest_sessionfinish(session, exitstatus):
    PARITY.dump()
