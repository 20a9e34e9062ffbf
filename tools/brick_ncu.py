"""General-matrix launches for ncu / timing: 512^3, the full affine G6 and one random rotation, all three
interpolators on the brick family.   usage: python tools/brick_ncu.py [--time]"""
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

n = 512
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
G6 = vt.utils.transform_matrix(center=c, scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                               rotation_order='rzxz', translation=(5.5, -3.25, 2.0))
rots = np.random.default_rng(1).uniform(-180, 180, (100, 3))
mats = {'G6': G6}
for i in (0, 1, 2, 3):
    mats[f'rand{i}'] = vt.utils.transform_matrix(rotation=tuple(rots[i]), rotation_order='sxyz', center=(256, 256, 256))
src = torch.rand(shape, device='cuda')
dst = torch.zeros(shape, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
st = torch.cuda.current_stream().cuda_stream
timing = '--time' in sys.argv
for name, m in mats.items():
    row = []
    for interp in (0, 1, 2):
        def go():
            _native.affine(src.data_ptr(), shape, dst.data_ptr(), shape, m, interp, _native.KERNEL_BRICK, device=0, stream=st)
        if not timing:
            if name in ('G6', 'rand0'):
                go()
                go()
            continue
        ts = []
        for it in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            go()
            e1.record()
            e1.synchronize()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        row.append(f'{n ** 3 / statistics.median(ts) / 1e6:.0f}')
    if timing:
        print(name, 'linear / cubic_tex / cubic_simple Gvox/s:', ' / '.join(row), flush=True)
torch.cuda.synchronize()
print('ok')
