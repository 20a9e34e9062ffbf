"""Slice-family kernels over a range of in-plane angles: the per-matrix warp shape chosen by the host (auto) against
the fixed 2 x 16 shape (VT_SLICE_LAYOUT=0).  usage: python tools/sweep_probe.py [n]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
src = torch.rand(shape, device='cuda')
dst = torch.zeros(shape, device='cuda')
st = torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


angles = list(range(0, 180, 15)) + [45, 135]
for interp, iname in ((0, 'linear'), (1, 'cubic_tex'), (2, 'cubic_simple')):
    tot = {'auto': 0.0, 'fixed': 0.0}
    for a in sorted(set(angles)):
        m = vt.utils.transform_matrix(rotation=(0, a, 0), rotation_order='rzxz', center=c)
        row = {}
        for label, env in (('fixed', '0'), ('auto', None)):
            if env is None:
                os.environ.pop('VT_SLICE_LAYOUT', None)
            else:
                os.environ['VT_SLICE_LAYOUT'] = env
            row[label] = timeit(lambda: _native.affine(src.data_ptr(), shape, dst.data_ptr(), shape, m, interp,
                                                       _native.OOB_ZERO | _native.KERNEL_SLICE, stream=st))
            tot[label] += row[label]
        print(f'{n}^3 {iname} angle {a:3d}: 2x16 {row["fixed"]:.4f} ms  auto {row["auto"]:.4f} ms  '
              f'({n ** 3 / row["auto"] / 1e6:.0f} Gvox/s)')
    print(f'{n}^3 {iname} sum over angles: 2x16 {tot["fixed"]:.3f} ms  auto {tot["auto"]:.3f} ms')
