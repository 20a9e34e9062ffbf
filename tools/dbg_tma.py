import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import voltools_b200 as vt
from voltools_b200 import _native as N
shape = (16, 32, 32)
vol = torch.rand(shape, device='cuda')
out = torch.zeros(shape, device='cuda')
m = vt.utils.transform_matrix(rotation=(0, 30, 0), center=np.divide(np.subtract(shape, 1), 2, dtype=np.float32))
torch.cuda.synchronize()
t0 = time.time()
try:
    N.affine(vol.data_ptr(), shape, out.data_ptr(), shape, m, 0, N.OOB_ZERO | N.KERNEL_SLICE)
    torch.cuda.synchronize()
    print('ok', time.time() - t0, float(out.sum()))
except Exception as e:
    print('FAIL after', time.time() - t0, e)
