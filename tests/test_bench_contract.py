"""bench.py keeps the driver's contract: one JSON line with the agreed keys, for the reference (CPU) arm here and
for the GPU arm on a B200 (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'cpu_baseline'}


def _run(args, timeout):
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, res.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(['--impl', 'reference', '--workload', 'xcorr128', '--steps', '2', '--warmup', '3'], 600)
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference'
    assert d['metric'] == 'xcorr_block_matches_per_sec' and d['unit'] == 'matches/s' and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'matches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


@pytest.mark.gpu
def test_gpu_arm_line():
    d = _run(['--steps', '5', '--warmup', '3', '--no-cpu-baseline'], 600)
    assert BASE_KEYS <= set(d) and {'clocks', 'gpu_launches', 'roofline'} <= set(d)
    assert d['n_gpus'] == 1 and d['steps'] == 5 and d['warmup'] >= 3 and d['dtype'] == 'f32' and d['data'] == 'synthetic'
    assert d['gpu_launches'] >= 4 * 5 and d['value'] > 1e4
    r = d['roofline']
    assert r['bound'] == 'hbm' and r['unit'] == 'GB/s' and 0 < r['frac'] < 1 and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    assert r['kernel'] in r['kernels'] and 0.3 < r['kernel_share_of_step'] < 0.8
    e = d['e2e']
    assert e['h2d_bytes_per_step'] == 2 * 256 * 512 * 512 * 4 and e['d2h_bytes_per_step'] > 0 and 0 < e['value'] < d['value']
    assert d['run_info']['ground_truth_recovered'] == 1.0 and set(d['config']) == {'workload', 'pairs_per_step_per_gpu', 'fft', 'l2'}
    assert d['clocks']['samples'] >= 3 and d['e2e']['h2d_ceiling_gbs'] > 0 and d['e2e']['pageable']['value'] > 0
    assert {'burst', 'sustained'} == set(d['regimes'])
