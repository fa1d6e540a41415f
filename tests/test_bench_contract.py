"""CPU: the reference arm of bench.py prints ONE JSON line with the contract keys (the native arm needs a GPU and is exercised by
the driver / tools/gpu_check.sh).  Small shape so the test takes seconds."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0', '--size', '64',
                        '--batch', '2', '--ref-sample', '2'], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = _run({'RY_BENCH_FORCE_PORT': '1'})
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['vs_baseline'] is None and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_uses_the_staged_reference_when_present():
    """baseline/_ref (tools/make_baseline_ref.py) holds the reference's own models/ utils/: the CPU arm then runs THEM (kind "reference")."""
    staged = os.path.exists(os.path.join(ROOT, 'baseline', '_ref', 'models', 'yolo.py'))
    d = _run()
    assert d['cpu_baseline']['kind'] == ('reference' if staged else 'port'), d['cpu_baseline']
