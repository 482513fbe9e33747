"""bench.py contract checks that need no GPU: the algorithmic-byte model against SURVEY.md 8(d)'s totals, and the reference
arm (the oracle port of the training step on the host cores) printing ONE JSON line with the contract's keys - alone on rank
0 when launched with WORLD_SIZE > 1."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location('kgc_bench', os.path.join(ROOT, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_algorithmic_bytes_match_survey_totals():
    b = _bench()
    want = {'wn18rr': (121.0e6, 206.5e6), 'fb15k237': (242.0e6, 465.6e6), 'wikidata5m': (22.5e9, 40.8e9)}
    for name, (fwd_w, bwd_w) in want.items():
        n, r, e, _ = b.WORKLOADS[name]
        fwd, bwd = b.algorithmic_bytes(n, r, e)
        assert abs(fwd - fwd_w) <= 0.005 * fwd_w and abs(bwd - bwd_w) <= 0.005 * bwd_w, (name, fwd, bwd)
    n, r, e, _ = b.WORKLOADS['wn18rr']
    assert round(sum(b.algorithmic_bytes(n, r, e)) / (2 * e)) == 1886            # bytes per directed edge


def _run_reference(env_extra):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return [ln for ln in p.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_contract_line():
    lines = _run_reference({'RANK': '0', 'WORLD_SIZE': '1'})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'edges/sec GCN fwd+bwd' and d['unit'] == 'edges/s'
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['higher_is_better'] is True and d['gpu_launches'] == 0
    assert d['value'] > 0 and abs(d['value'] - 2 * 86835 / (d['ms_per_step'] * 1e-3)) <= 1e-6 * d['value']
    assert d['config']['workload'] == 'wn18rr_shape'
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    assert _run_reference({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}) == []
