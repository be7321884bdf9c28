"""CPU checks of bench.py: the reference arm (oracle port on the host cores) prints one well-formed JSON line, and the roofline
bookkeeping reproduces the SURVEY section 8d figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bytes_fwd_step_matches_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    # cfg 2 (B = 64, L = 180, M = 288) -> 86.54 MB per forward step in fp32; cfg 1 (B = 16, M = 512) -> 80.38 MB
    assert abs(bench.bytes_fwd_step(64, 180, 288) / 1e6 - 86.54) < 0.01
    assert abs(bench.bytes_fwd_step(16, 180, 512) / 1e6 - 80.38) < 0.01


def test_reference_arm_prints_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS='4')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
                          '--frames', '900', '--batch', '10', '--text-len', '40'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'dtype', 'data',
                'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['d2h_bytes_per_step'] == 0
