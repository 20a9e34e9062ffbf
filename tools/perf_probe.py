"""Quick device-side timing of our kernels vs the reference's own kernels on this GPU (not the bench)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402


def timeit(fn, iters=10, warm=3):
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


def main():
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [256, 512]
    for n in sizes:
        shape = (n, n, n)
        nvox = n ** 3
        rng = np.random.default_rng(0)
        vol = rng.random(shape, dtype=np.float32)
        c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
        mats = {
            'rot45x': vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c),
            'affine': vt.utils.transform_matrix(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02),
                                                rotation=(30, 45, 60), rotation_order='rzxz',
                                                translation=(5.5, -3.25, 2.0), center=c),
        }
        src = torch.from_numpy(vol).cuda()
        # the resampling kernels see what the API hands them: rows padded to 16 bytes (TMA staging)
        row = _native.padded_row(n)
        srcp = torch.zeros((n, n, row), device='cuda')
        srcp[:, :, :n].copy_(src)
        strides = (row, n * row)
        dst = torch.zeros(shape, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        tmp = torch.empty_like(src)
        for variant in (1, 0):
            _native.profile_enable(True)
            ms = timeit(lambda: _native.prefilter(src.data_ptr(), shape, 0, st, variant=variant, dst_ptr=tmp.data_ptr()),
                        iters=5, warm=2)
            prof = _native.profile_read()
            _native.profile_enable(False)
            print(f'{n}^3 prefilter variant {variant}: {ms:.3f} ms  {nvox / ms / 1e6:.1f} Gvox/s  '
                  f'({8 * nvox / ms / 1e6 / 6549.1 * 100:.1f}% of 8B/vox roofline)  '
                  + ' '.join(f'{k}={v[0] / v[1]:.3f}ms' for k, v in prof.items()))
        for mname, m in mats.items():
            for interp, iname in ((0, 'linear'), (1, 'cubic_tex'), (2, 'cubic_simple')):
                for fam, fname in ((_native.KERNEL_GATHER, 'gather'), (_native.KERNEL_BRICK, 'brick'),
                                   (_native.KERNEL_SLICE, 'slice')):
                    try:
                        ms = timeit(lambda: _native.affine(srcp.data_ptr(), shape, dst.data_ptr(), shape, m, interp,
                                                           _native.OOB_ZERO | fam, stream=st, src_strides=strides))
                    except RuntimeError as e:
                        continue
                    print(f'{n}^3 {mname} {iname} {fname}: {ms:.3f} ms  {nvox / ms / 1e6:.1f} Gvox/s  '
                          f'({8 * nvox / ms / 1e6 / 6549.1 * 100:.1f}% roofline)')
        # texture family (general matrices, linear / cubic_tex): upload cost and kernel
        ms_up = timeit(lambda: _native.Texture(srcp.data_ptr(), shape, strides, 0, st).close(), iters=3, warm=1)
        tex = _native.Texture(srcp.data_ptr(), shape, strides, 0, st)
        print(f'{n}^3 texture create+upload+destroy: {ms_up:.3f} ms')
        ms_up = timeit(lambda: tex.upload(srcp.data_ptr(), strides, st), iters=5, warm=1)
        print(f'{n}^3 texture upload only: {ms_up:.3f} ms  ({8 * nvox / ms_up / 1e6:.0f} GB/s)')
        for mname, m in mats.items():
            for interp, iname in ((0, 'linear'), (1, 'cubic_tex')):
                ms = timeit(lambda: tex.affine(dst.data_ptr(), shape, m, interp, _native.OOB_ZERO, stream=st))
                ref = torch.zeros(shape, device='cuda')
                _native.affine(srcp.data_ptr(), shape, ref.data_ptr(), shape, m, interp, _native.OOB_ZERO | _native.KERNEL_GATHER,
                               stream=st, src_strides=strides)
                err = float((ref - dst).abs().max())
                print(f'{n}^3 {mname} {iname} tex: {ms:.3f} ms  {nvox / ms / 1e6:.1f} Gvox/s  '
                      f'({8 * nvox / ms / 1e6 / 6549.1 * 100:.1f}% roofline)  max|tex - gather| = {err:.2e}')
        tex.close()
        if oracle.ref_gpu_available() and n <= 512 and '--ref' in sys.argv:
            for mname, m in mats.items():
                for mode in ('linear', 'bspline', 'bspline_simple', 'filt_bspline'):
                    _, msk, msp = oracle.transform_ref_gpu(vol, m, mode, iters=5)
                    print(f'{n}^3 {mname} REFERENCE kernels {mode}: kernel {msk:.3f} ms '
                          f'({nvox / msk / 1e6:.1f} Gvox/s)' + (f' prefilter {msp:.3f} ms' if msp else ''))


if __name__ == '__main__':
    main()
