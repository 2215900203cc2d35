"""bench.py contract on the CPU: the reference arm (the one arm that runs without a GPU) prints exactly one JSON
line with the keys the driver reads; ranks other than 0 print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'impl', 'e2e', 'cpu_baseline', 'gpu_launches')


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                           '--warmup', '0', '--cpu-sample', '1'], cwd=ROOT, env=env, capture_output=True, text=True,
                          timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'images/s'
    assert d['value'] > 0 and d['e2e']['value'] == d['value'] == d['cpu_baseline']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert 'workload' in d['config'] and d['gpu_launches'] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''
