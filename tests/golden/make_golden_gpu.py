"""Generate golden vectors by running the REFERENCE's own CUDA kernels on a real GPU (B200).

Run on the GPU box:  python tests/golden/make_golden_gpu.py   (writes gpurun_out/golden/*.npz; copy them
into tests/golden/ and commit).  Uses only oracle/_ref/libvt_ref_gpu.so + the cubins compiled from the
reference's captured kernel sources (oracle/build_ref.py); never reads /root/reference at run time.

Contents
  ref_gpu_cases.npz : for a seeded 20x24x28 volume and 5 matrices, the outputs of all five interpolation
                      modes of the reference GPU path, plus the reference prefilter of that volume, plus an
                      `output=` case (pre-filled destination, out-of-bounds voxels retained).
  ref_tex_probe.npz : raw tex3D<float> fetches (border/linear/unnormalised, the reference's descriptor) of a
                      ramp volume at 1/4096-spaced coordinates and of a random volume at random coordinates:
                      pins the fixed-point rule of the texture unit that the oracle emulates.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from voltools_b200.utils import transform_matrix  # noqa: E402

MODES = ['linear', 'bspline', 'bspline_simple', 'filt_bspline', 'filt_bspline_simple']
SHAPE = (20, 24, 28)


def matrices(shape):
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    return {
        'identity': np.identity(4, dtype=np.float32),
        'shift': transform_matrix(translation=(0.37, -1.21, 2.5)),
        'rot45': transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c),
        'rot_general': transform_matrix(rotation=(33.3, -71.0, 12.5), rotation_order='sxyz', center=c),
        'full_affine': transform_matrix(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                                        rotation_order='rzxz', translation=(1.5, -0.75, 0.5), center=c),
    }


def main():
    out = ROOT / 'gpurun_out' / 'golden'
    out.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(1234)
    vol = rng.random(SHAPE, dtype=np.float32)
    data = {'volume': vol, 'prefiltered': oracle.prefilter_ref_gpu(vol)}
    for name, m in matrices(SHAPE).items():
        data[f'm_{name}'] = m
        for mode in MODES:
            res, _, _ = oracle.transform_ref_gpu(vol, m, mode)
            data[f'out_{name}_{mode}'] = res
    prefill = rng.random(SHAPE, dtype=np.float32) + 10.0
    data['prefill'] = prefill
    for mode in ('linear', 'filt_bspline'):
        res, _, _ = oracle.transform_ref_gpu(vol, data['m_rot45'], mode, output=prefill)
        data[f'outprefill_rot45_{mode}'] = res
    np.savez_compressed(out / 'ref_gpu_cases.npz', **data)

    # texture probe 1: ramp along x (axis 2), sampled at 2.5 + k/4096 (+ exact ties k/512), y = z = texel centre
    ramp = np.broadcast_to(np.arange(8, dtype=np.float32), (2, 2, 8)).copy()
    k = np.arange(0, 4097, dtype=np.float64)
    xs = (2.5 + k / 4096.0).astype(np.float32)
    xyz = np.stack([xs, np.full_like(xs, 0.5), np.full_like(xs, 0.5)], axis=1)
    ramp_out = oracle.tex3d_ref_gpu(ramp, xyz)
    # same ramp, far from the origin (coordinate magnitude matters if the conversion is done in float)
    big = np.broadcast_to(np.arange(600, dtype=np.float32), (2, 2, 600)).copy()
    xs2 = (517.5 + k / 4096.0).astype(np.float32)
    xyz2 = np.stack([xs2, np.full_like(xs2, 0.5), np.full_like(xs2, 0.5)], axis=1)
    big_out = oracle.tex3d_ref_gpu(big, xyz2)
    # probe 2: random volume, random coordinates including the border rim
    rv = rng.random((9, 10, 11), dtype=np.float32)
    n = 200000
    coords = np.stack([rng.uniform(-1.5, 12.5, n), rng.uniform(-1.5, 11.5, n), rng.uniform(-1.5, 10.5, n)],
                      axis=1).astype(np.float32)
    rand_out = oracle.tex3d_ref_gpu(rv, coords)
    np.savez_compressed(out / 'ref_tex_probe.npz', ramp=ramp, ramp_xyz=xyz, ramp_out=ramp_out, big_xyz=xyz2,
                        big_out=big_out, rand_vol=rv, rand_xyz=coords[:20000], rand_out=rand_out[:20000])

    # analysis printed for the log
    alpha = ramp_out - 2.0
    for rule, nm in ((oracle.TEX_RN, 'round-nearest'), (oracle.TEX_TRUNC, 'truncate'), (oracle.TEX_EXACT, 'exact')):
        e1 = np.abs(oracle.tex3d_many(ramp, xyz, rule) - ramp_out).max()
        e2 = np.abs(oracle.tex3d_many(big, xyz2, rule) - big_out).max()
        e3 = np.abs(oracle.tex3d_many(rv, coords, rule) - rand_out).max()
        print(f'tex rule {nm:14s}: ramp max|d|={e1:.3e}  far-ramp max|d|={e2:.3e}  random max|d|={e3:.3e}')
    print('distinct alpha levels on the ramp:', len(np.unique(alpha)), ' first steps at k =',
          np.nonzero(np.diff(alpha))[0][:6])


if __name__ == '__main__':
    main()
