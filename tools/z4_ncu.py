"""A handful of slice4 launches for ncu: 512^3 cubic_tex / cubic_simple / linear at 0 and 45 degrees (axis 0), one
axis-2 launch, and a 32-matrix batch at 256^3.   usage: python tools/z4_ncu.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

st = torch.cuda.current_stream().cuda_stream


def z4_of(src, axis):
    buf = torch.empty(_native.z4_bytes(src.shape, axis) // 4, dtype=torch.float32, device=src.device)
    _native.pack_z4(src.data_ptr(), src.shape, buf.data_ptr(), axis, device=0, stream=st)
    return buf


n = 512
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
src = torch.rand(shape, device='cuda')
dst = torch.zeros(shape, device='cuda')
z4 = z4_of(src, 0)
for interp in (1, 2, 0):
    for angle in (0, 45):
        m = vt.utils.transform_matrix(rotation=(0, angle, 0), rotation_order='rzxz', center=c)
        for _ in range(2):
            _native.affine_z4(z4.data_ptr(), 0, shape, dst.data_ptr(), shape, m, interp, 0, device=0, stream=st)
z4 = z4_of(src, 2)
m = vt.utils.transform_matrix(rotation=(45, 0, 0), rotation_order='rzxz', center=c)
for _ in range(2):
    _native.affine_z4(z4.data_ptr(), 2, shape, dst.data_ptr(), shape, m, 1, 0, device=0, stream=st)
del z4, src, dst
n = 256
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
src = torch.rand(shape, device='cuda')
z4 = z4_of(src, 0)
mats = np.stack([vt.utils.transform_matrix(rotation=(0, a, 0), rotation_order='rzxz', center=c) for a in range(32)])
out = torch.empty((32,) + shape, device='cuda')
for _ in range(2):
    _native.affine_z4(z4.data_ptr(), 0, shape, out.data_ptr(), shape, mats, 1, 1, device=0, stream=st)
torch.cuda.synchronize()
print('ok')
