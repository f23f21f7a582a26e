"""bench.py's reference arm (`--impl reference`) runs on CPU: check the JSON contract the round driver parses, and that only
rank 0 prints under a multi-rank launch."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--batch', '4', *args],
                          capture_output=True, text=True, timeout=600, env=env, cwd=REPO)


def test_reference_arm_json_contract():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                            # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'wgan_gd_step_voxels_per_s' and d['unit'] == 'voxels/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['config']['workload'].startswith('configs[1]')           # same workload string as the B200 arm
    assert d['e2e'] == {'value': d['value'], 'unit': 'voxels/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'full B=4 step' in cb['sample']
    assert 'WGAN-GP' in d['config']['workload'] and d['config']['global_batch'] == 4      # default workload = BASELINE configs[1]


def test_reference_arm_prints_on_rank0_only():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}, '--gpus', '2')
    assert r.returncode == 0 and r.stdout.strip() == ''
