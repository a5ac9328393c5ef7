"""bench.py's CPU arm (`--impl reference`: the fp64 oracle on the host cores, the one leg of bench.py that may execute
`oracle/`) keeps the JSON contract the driver parses. No GPU needed; the GPU arm's line is checked on the GPU box
(profiles/r2c_bench*.json are its outputs)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '2',
                        '--warmup', '1', '--preroll', '300'] + extra,
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run([])
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out                      # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['higher_is_better'] is True and d['unit'] == 'env-steps/s'
    assert d['metric'] == 'monopod env-steps/sec at 64K envs/GPU' and d['n_gpus'] == 1
    assert d['steps'] == 2 and d['warmup'] == 1 and d['dtype'] == 'f64' and d['gpu_launches'] == 0
    assert d['value'] > 0 and d['ms_per_step'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'NOT Gazebo' in d['config']['note']
    # the sample really is in the contact regime the metric names (hip proxy pressing in a fair share of the envs)
    assert d['config']['contact_frac'][0] > 0.05


def test_reference_arm_is_rank_zero_only_under_torchrun():
    out = _run(['--gpus', '2'], env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''                         # the other ranks exit 0 without work and without a line
