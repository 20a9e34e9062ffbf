"""Ground truth for the slice family's (box width, warp shape) choice: time every combination at a few angles.
usage: python tools/slice_grid_probe.py [n] [angles...]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
angles = [int(a) for a in sys.argv[2:]] or [45, 30, 10]
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
src = torch.rand(shape, device='cuda')
dst = torch.zeros(shape, device='cuda')
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, iters=5, warm=2):
    ts = []
    for it in range(warm + iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if it >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for a in angles:
    m = vt.utils.transform_matrix(rotation=(0, a, 0), rotation_order='rzxz', center=c)
    for interp, iname in ((0, 'linear'), (1, 'cubic_tex'), (2, 'cubic_simple')):
        call = lambda: _native.affine(src.data_ptr(), shape, dst.data_ptr(), shape, m, interp,  # noqa: E731
                                      _native.OOB_ZERO | _native.KERNEL_SLICE, stream=st)
        os.environ.pop('VT_SLICE_W', None)
        os.environ.pop('VT_SLICE_LAYOUT', None)
        os.environ['VT_SLICE_DEBUG'] = '1'
        call()
        torch.cuda.synchronize()
        os.environ.pop('VT_SLICE_DEBUG')
        auto = timeit(call)
        cells = []
        for w in (24, 28, 32, 36, 40):
            for layout in (0, 1, 2):
                os.environ['VT_SLICE_W'] = str(w)
                os.environ['VT_SLICE_LAYOUT'] = str(layout)
                cells.append(f'w{w}/s{layout}={timeit(call, iters=3, warm=1):.3f}')
        print(f'{n}^3 angle {a} {iname}: auto {auto:.3f} ms | ' + ' '.join(cells), flush=True)
