"""bench.py's contract where it can be checked without a GPU: the reference arm prints one JSON
line with the agreed keys, and the product arm refuses to run without a CUDA device instead of
falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], cwd=ROOT,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    proc = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-sample', '64')
    assert proc.returncode == 0, proc.stderr
    lines = [l for l in proc.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'ivp_solves_per_sec_fwd_adjoint'
    assert d['unit'] == 'solves/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['config']['workload'] == 'lv_adj'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['gpu_launches'] == 0


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    proc = _run('--steps', '1', '--warmup', '1')
    assert proc.returncode != 0
    assert 'no CPU fallback' in (proc.stdout + proc.stderr)
    assert not [l for l in proc.stdout.splitlines() if l.startswith('{')]
