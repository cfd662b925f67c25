"""bench.py's JSON contract: the committed line of the final GPU run carries every key the driver reads, and the reference arm
(`--impl reference`, the CPU ref path through the oracle) prints one line with the same metric / unit / config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype',
            'data', 'config', 'roofline', 'cpu_baseline', 'e2e', 'clocks', 'gpu_launches')


def test_committed_bench_line_has_the_contract_keys():
    line = json.load(open(os.path.join(ROOT, 'profiles', 'r02_bench_v7.json')))
    assert all(k in line for k in REQUIRED), [k for k in REQUIRED if k not in line]
    assert line['metric'] == 'generator_512px_images_per_sec' and line['unit'] == 'images/s' and line['higher_is_better'] is True
    assert line['n_gpus'] == 1 and line['warmup'] >= 3 and line['scaling'] == 'weak' and line['vs_baseline'] is None
    assert 'workload' in line['config'] and 'model' not in line['config']
    roof = line['roofline']
    assert roof['bound'] in ('hbm', 'tensor') and abs(roof['frac'] - roof['achieved'] / roof['peak']) < 1e-9 and roof['traffic']
    assert set(('value', 'unit', 'cores', 'kind', 'sample')) <= set(line['cpu_baseline'])
    e2e = line['e2e']
    assert e2e['h2d_bytes_per_step'] > 0 and e2e['d2h_bytes_per_step'] > 0 and e2e['value'] != line['value']
    assert line['gpu_launches'] > 0 and 'sm_mhz' in line['clocks'] and 'reasons' in line['clocks']
    assert line['eager']['ms_per_step'] > 0 and 'launch' in line['config']          # graph replay is the timed entry; eager figure beside it
    assert set(('modulated_conv2d', 'upfirdn2d', 'bias_act', 'conv2d_gradfix_train')) <= set(line['ops'])       # BASELINE configs[3]
    assert line['train']['ms_per_step'] > 0 and line['train']['batch_per_gpu'] == 8                            # BASELINE configs[4]


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'] == 'generator_512px_images_per_sec' and line['unit'] == 'images/s'
    assert line['value'] > 0 and line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
