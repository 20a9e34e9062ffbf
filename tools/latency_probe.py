"""Per-call latency of the public API on small volumes (the reference's README table: 5^3 .. 50^3 are pure overhead)."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402

kw = dict(rotation=(0, 45, 0), rotation_order='rzxz')
gen = dict(rotation=(30, 45, 60), rotation_order='rzxz')
for n in (16, 32, 64, 100):
    shape = (n, n, n)
    v_np = np.random.default_rng(0).random(shape, dtype=np.float32)
    v_d = torch.from_numpy(v_np).cuda()
    o_d = torch.zeros_like(v_d)
    for mode in ('linear', 'filt_bspline'):
        sv = vt.StaticVolume(v_d, interpolation=mode, device='gpu:0')
        cases = {
            'transform(numpy) -> numpy': lambda: vt.transform(v_np, interpolation=mode, device='gpu:0', **kw),
            'transform(device, output=)': lambda: vt.transform(v_d, interpolation=mode, output=o_d, device='gpu:0', **kw),
            'transform(device, output=) general': lambda: vt.transform(v_d, interpolation=mode, output=o_d, device='gpu:0', **gen),
            'StaticVolume.transform(output=)': lambda: sv.transform(output=o_d, **kw),
        }
        row = []
        for name, fn in cases.items():
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                fn()
            torch.cuda.synchronize()
            row.append(f'{name}: {(time.perf_counter() - t0) / 50 * 1e6:.0f} us')
        print(f'{n}^3 {mode}: ' + ' | '.join(row), flush=True)
t0 = time.perf_counter()
for _ in range(200):
    vt.utils.transform_matrix(center=(7.5, 7.5, 7.5), **kw)
print(f'transform_matrix: {(time.perf_counter() - t0) / 200 * 1e6:.0f} us')
