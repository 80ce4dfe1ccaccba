"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the reference forward on
the host cores through the oracle port) prints ONE JSON line with the agreed keys, also under torchrun-style environment
variables where only rank 0 may print; the product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, env=e,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    r = run(['--impl', 'reference', '--steps', '1', '--warmup', '1', '--hw', '64', '96'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['value'] > 0 and d['unit'] == 'images/s'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_stay_silent():
    r = run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1', '--hw', '64', '96'],
            env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_product_arm_has_no_cpu_fallback():
    r = run(['--steps', '1', '--warmup', '1'])
    assert r.returncode != 0 and 'GPU' in (r.stderr + r.stdout)
