"""Launch each hot kernel a few times at one size (for ncu -k regex captures).  usage: prof_kernels.py [n] [what...]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native as N  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
what = sys.argv[2:] or ['prefilter', 'slice', 'gather']
shape = (n, n, n)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
rot = vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c)
aff = vt.utils.transform_matrix(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                                rotation_order='rzxz', translation=(5.5, -3.25, 2.0), center=c)
src = torch.rand(shape, device='cuda')
dst = torch.empty_like(src)
st = torch.cuda.current_stream().cuda_stream
for it in range(2):
    if 'prefilter' in what:
        N.prefilter(src.data_ptr(), shape, 0, st, variant=0, dst_ptr=dst.data_ptr())
    for interp in (0, 1, 2):
        if 'slice' in what:
            N.affine(src.data_ptr(), shape, dst.data_ptr(), shape, rot, interp, N.OOB_ZERO | N.KERNEL_SLICE, stream=st)
        if 'gather' in what:
            N.affine(src.data_ptr(), shape, dst.data_ptr(), shape, aff, interp, N.OOB_ZERO | N.KERNEL_GATHER, stream=st)
        if 'brick' in what:
            N.affine(src.data_ptr(), shape, dst.data_ptr(), shape, aff, interp, N.OOB_ZERO | N.KERNEL_BRICK, stream=st)
if 'project' in what:
    sv = vt.StaticVolume(src, interpolation='filt_bspline', device='gpu:0')
    tilt = vt.utils.transform_matrix(rotation=(30, 0, 0), rotation_order='sxyz', center=c)
    for it in range(2):
        sv.project_many([tilt])
        sv.project_many([aff])
torch.cuda.synchronize()
print('ok', N.launch_count())
