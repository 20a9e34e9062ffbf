"""bench.py's contract, as far as it can be checked without a GPU: the reference arm's JSON line (the reference's own CPU
path through its public API, bounded sample), the helper arithmetic, and that our arm refuses to run without a CUDA device
(no CPU fallback)."""
import importlib.util
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def _bench():
    spec = importlib.util.spec_from_file_location('vt_bench', ROOT / 'bench.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_line():
    env = dict(os.environ, VT_BENCH_REF_BUDGET_S='6')
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference'
    if 'unavailable' in line:   # the reference package is staged by __graft_entry__.build() from /root/reference
        assert isinstance(line['unavailable'], str) and line['unavailable']
        return
    bench = _bench()
    assert line['metric'] == line['unit'] == bench.METRIC and line['higher_is_better'] is True
    assert line['config']['workload'] == bench.SWEEP['name'] and line['scaling'] == 'strong' and line['dtype'] == 'f32'
    assert line['value'] > 0 and line['steps'] == 1 and line['warmup'] == 1 and line['vs_baseline'] is None
    cb = line['cpu_baseline']
    assert cb['kind'] == 'reference' and cb['cores'] >= 1 and cb['value'] == line['value'] and cb['sample']
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['unit'] == line['unit']
    assert e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0


def test_our_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--steps', '1', '--warmup', '1'], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_bench_helpers():
    bench = _bench()
    import voltools_b200 as vt
    shape = (64, 64, 64)
    assert bench.inbounds_fraction(shape, np.identity(4)) == 1.0
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    m = vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c)
    assert abs(bench.inbounds_fraction(shape, m) - 2 * (np.sqrt(2) - 1)) < 0.03   # 82.8 % for a centred 45 deg turn
    kw = [bench._sweep_kw(i) for i in (0, 1, 179)]
    assert kw[1]['rotation'] == (0, 1, 0) and kw[2]['rotation'] == (0, 179, 0) and kw[0]['rotation_order'] == 'rzxz'
    assert bench.SWEEP['angles'] == 180 and bench.SWEEP['n'] == 256 and bench.SWEEP['interpolation'] == 'filt_bspline'
