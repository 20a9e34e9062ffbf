"""Two 32-matrix slice4 launches at 256^3 (the headline's launch shape) for ncu.   usage: python tools/z4_sweep_ncu.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
n = 256
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
src = torch.rand(shape, device='cuda')
z4 = torch.empty(_native.z4_bytes(shape, 0) // 4, dtype=torch.float32, device='cuda')
_native.pack_z4(src.data_ptr(), shape, z4.data_ptr(), 0, device=0, stream=st)
mats = np.stack([vt.utils.transform_matrix(rotation=(0, a, 0), rotation_order='rzxz', center=c) for a in range(0, 180, 6)] * 2)[:32]
out = torch.empty((32,) + shape, device='cuda')
for _ in range(2):
    _native.affine_z4(z4.data_ptr(), 0, shape, out.data_ptr(), shape, mats, 1, 1, device=0, stream=st)
torch.cuda.synchronize()
print('ok')
