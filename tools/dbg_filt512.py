"""Debug: where do ours / reference disagree on zero-ness at 512^3 filt_bspline rot45?"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import oracle
import voltools_b200 as vt
from voltools_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = np.random.default_rng(2).random((n, n, n), dtype=np.float32)
shape = vol.shape
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
m = vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c)
st = torch.cuda.current_stream().cuda_stream
src = torch.from_numpy(vol).cuda()
# prefilter -> Z4 directly vs prefilter plain + pack
plain = torch.empty(shape, device='cuda')
_native.prefilter(src.data_ptr(), shape, 0, st, dst_ptr=plain.data_ptr())
want = torch.empty(_native.z4_bytes(shape, 0) // 4, device='cuda')
_native.pack_z4(plain.data_ptr(), shape, want.data_ptr(), 0, device=0, stream=st)
got = torch.full_like(want, float('nan'))
ws = torch.empty(shape, device='cuda')
_native.prefilter_z4(src.data_ptr(), shape, got.data_ptr(), ws.data_ptr(), ws.numel() * 4, 0, st)
torch.cuda.synchronize()
d = (got - want).abs().view((n + 3) // 4, n, n, 4)
print('prefilter_z4 vs pack: max', float(d.nan_to_num(1e9).max()), 'nan', int(torch.isnan(got).sum()))
bad = (d.nan_to_num(1e9) > 1e-4).nonzero()
print('bad entries', len(bad), bad[:10].tolist())
for mode in ('filt_bspline', 'bspline'):
    ref, _, _ = oracle.transform_ref_gpu(vol, m, mode)
    for trial in range(3):
        out = torch.zeros(shape, device='cuda')
        vt.affine(src, m, interpolation=mode, output=out, device='gpu:0')
        o = out.cpu().numpy()
        diff = (o == 0) != (ref == 0)
        per_plane = diff.reshape(n, -1).sum(1)
        zs = np.nonzero(per_plane)[0]
        print(mode, 'trial', trial, 'mismatching voxels', int(diff.sum()), 'planes', zs[:20].tolist(), 'max abs err', float(np.abs(o - ref).max()))
        if diff.sum():
            idx = np.argwhere(diff)[:8]
            for z, y, x in idx:
                print('   at', (z, y, x), 'ours', o[z, y, x], 'ref', ref[z, y, x])
