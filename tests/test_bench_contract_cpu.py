"""CPU check of bench.py's reference arm (the one bench leg that runs without a GPU): it prints ONE JSON line with the
contract's keys, times the oracle port on the host cores, and exits 0 on non-zero ranks without doing any work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', *args],
                          capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run({}, '--steps', '2', '--warmup', '0', '--envs', '256', '--batch', '128')
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'env-steps/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 2 and d['n_gpus'] == 1 and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_is_silent_on_other_ranks():
    out = _run({'RANK': '1', 'WORLD_SIZE': '2'}, '--gpus', '2', '--steps', '2', '--warmup', '0')
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_committed_bench_line_carries_the_contract_keys():
    """The bench line committed under profiles/ (the last `python bench.py` of the round, one B200) has every key the contract
    names, with the meanings the contract gives them: guards the evidence file against a stale format."""
    import glob
    import json
    paths = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r3*_bench_n1.json')))
    assert paths, 'no committed single-GPU bench line'
    line = json.loads(open(paths[-1]).read().strip().splitlines()[-1])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
                'dtype', 'data', 'config', 'roofline', 'cpu_baseline', 'e2e', 'gpu_launches', 'clocks'):
        assert key in line, key
    assert line['n_gpus'] == 1 and line['higher_is_better'] is True and line['scaling'] == 'weak' and line['vs_baseline'] is None
    assert line['unit'] == 'env-steps/s' and 'workload' in line['config'] and 'l2' in line['config']
    r = line['roofline']
    assert r['bound'] in ('hbm', 'tensor', 'fp32') and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9 and r['traffic'] > 0
    c = line['cpu_baseline']
    assert c['kind'] in ('port', 'reference') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
    e = line['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0 and e['value'] < line['value']
    assert line['gpu_launches'] > 0 and line['value'] / c['value'] > 10
    assert not set(line['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}
