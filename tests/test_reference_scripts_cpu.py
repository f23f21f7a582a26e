"""CPU: the reference's training scripts, UNMODIFIED, resolve every import / constructor / optimizer / dataset against the drop-in
`model` package (tools/run_reference_script.py).  This container has no GPU and the GPU box has no /root/reference, so the scripts
can be followed exactly up to their first forward pass, where the package raises its "no CPU fallback" error -- which is the assertion.
Skipped where the reference tree is absent."""
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'train_wgan.py')), reason='reference tree not present')


@pytest.mark.parametrize('script,extra', [('train_wgan.py', ['nogui']), ('train_gan.py', ['nogui']), ('train_autoencoder.py', ['nogui'])])
def test_unmodified_script_reaches_our_forward(script, extra, tmp_path):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.join(REPO, 'tools', 'run_reference_script.py'), os.path.join(REF, script), '--samples', '8',
                        '--workdir', str(tmp_path), '--'] + extra, capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode != 0 and 'shapegan_b200 runs on CUDA (sm_100a) only' in tail, tail
    assert 'ImportError' not in tail and 'AttributeError' not in tail and 'ModuleNotFoundError' not in tail, tail
    assert os.path.isdir(os.path.join(str(tmp_path), 'models'))          # util.py created the reference's directory layout in the scratch cwd
