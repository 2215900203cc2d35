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


def test_roofline_traffic_comes_from_the_committed_capture_of_both_gemm_kernels():
    """bench.measured_traffic: DRAM bytes per GEMM launch = the committed ncu capture of one bench step, summed over the
    one-CTA and the CTA-pair kernel; a capture whose launch count does not make whole steps is not quoted."""
    sys.path.insert(0, ROOT)
    import bench
    per_launch, src = bench.measured_traffic(('k_gemm_tc', 'k_gemm_tc2'), 37)
    table = json.load(open(os.path.join(ROOT, 'profiles', 'r2_dram_traffic_bench_step.json')))['kernels']
    total = sum(table[k]['dram_read_bytes'] + table[k]['dram_write_bytes'] for k in ('k_gemm_tc', 'k_gemm_tc2'))
    assert table['k_gemm_tc']['launches'] + table['k_gemm_tc2']['launches'] == 37
    assert per_launch == round(total / 37) and 'r2_dram_traffic_bench_step.json' in src
    # within 5 % of the algorithmic bytes per launch of the committed bench line: no wasted traffic in the GEMMs
    line = json.loads(open(os.path.join(ROOT, 'profiles', 'r2_final_bench_n1_full.json')).read().strip().splitlines()[-1])
    assert abs(per_launch / line['roofline']['algorithmic_bytes_per_launch'] - 1.0) < 0.05
    assert bench.measured_traffic(('k_gemm_tc', 'k_gemm_tc2'), 36)[0] is None
