"""Randomised GPU parity: random shapes (odd widths hit the cp.async staging and the gather family, multiples of 4 the TMA
paths), random modes and random matrices of every class the dispatcher distinguishes -- in-plane rotations about axis 0
at arbitrary angles (slice family: every warp shape / box width the host may pick), integer and fractional shifts,
anisotropic scales (minification: bricks that do not fit -> gather), general rotations and full affines -- through the
public API against the CPU oracle, with the per-mode tolerances of tests/test_gpu_parity.py (asserted here: 1e-6 of the
sampled volume's range for every mode, as there).  Seeds are fixed: failures reproduce.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

MODES = ['linear', 'bspline', 'bspline_simple', 'filt_bspline', 'filt_bspline_simple']


@pytest.fixture(scope='module')
def vt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import voltools_b200 as vt
    return vt


def _random_matrix(vt, rng, shape, kind):
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    tm = vt.utils.transform_matrix
    if kind == 'axis0_rotation':
        t = (float(rng.integers(-3, 4)), float(rng.uniform(-4, 4)), float(rng.uniform(-4, 4)))
        return tm(rotation=(float(rng.uniform(-180, 180)), 0, 0), rotation_order='sxyz', translation=t, center=c)
    if kind == 'axis0_rotation_scaled':
        return tm(rotation=(0, float(rng.uniform(-180, 180)), 0), rotation_order='rzxz',
                  scale=(1.0, float(rng.uniform(0.7, 1.5)), float(rng.uniform(0.7, 1.5))), center=c)
    if kind == 'shift':
        return tm(translation=tuple(float(v) for v in rng.uniform(-5, 5, 3)))
    if kind == 'integer_shift':
        return tm(translation=tuple(float(v) for v in rng.integers(-5, 6, 3)))
    if kind == 'scale':
        return tm(scale=tuple(float(v) for v in rng.uniform(0.4, 3.0, 3)), center=c)
    if kind == 'rotation':
        return tm(rotation=tuple(float(v) for v in rng.uniform(-180, 180, 3)),
                  rotation_order=str(rng.choice(['rzxz', 'sxyz', 'rxyz', 'szyx'])), center=c)
    return tm(scale=tuple(float(v) for v in rng.uniform(0.8, 1.25, 3)), shear=tuple(float(v) for v in rng.uniform(-0.1, 0.1, 3)),
              rotation=tuple(float(v) for v in rng.uniform(-180, 180, 3)), rotation_order='rzxz',
              translation=tuple(float(v) for v in rng.uniform(-4, 4, 3)), center=c)


KINDS = ['axis0_rotation', 'axis0_rotation_scaled', 'shift', 'integer_shift', 'scale', 'rotation', 'full_affine']


@pytest.mark.parametrize('seed', range(12))
def test_random_transforms_vs_oracle(vt, seed):
    rng = np.random.default_rng(1000 + seed)
    shape = tuple(int(v) for v in rng.integers(9, 72, 3))
    if seed % 3 == 0:  # rows that are multiples of 16 bytes: TMA staging
        shape = shape[:2] + (int(rng.integers(3, 18)) * 4,)
    vol = rng.random(shape, dtype=np.float32)
    coef_range = float(np.ptp(oracle.prefilter(vol)))
    for kind in KINDS:
        mode = MODES[int(rng.integers(0, len(MODES)))]
        m = _random_matrix(vt, rng, shape, kind)
        r = coef_range if mode.startswith('filt') else float(np.ptp(vol))
        want = oracle.affine(vol, m, mode)
        got = vt.affine(vol, m, interpolation=mode, device='gpu:0')
        e = float(np.abs(got.astype(np.float64) - want).max()) / r
        assert e <= 1e-6, f'seed {seed} {shape} {kind} {mode}: {e:.3e}'


@pytest.mark.parametrize('seed', range(6))
def test_random_static_volume_batches_and_projections(vt, seed):
    """One resident volume, a batch mixing slice-family and general matrices (affine_many chunks by family per launch),
    output= retention, and the fused projections of the same matrices."""
    import torch
    rng = np.random.default_rng(2000 + seed)
    shape = tuple(int(v) for v in rng.integers(12, 56, 3))
    vol = rng.random(shape, dtype=np.float32)
    mode = MODES[seed % len(MODES)]
    r = float(np.ptp(oracle.prefilter(vol))) if mode.startswith('filt') else float(np.ptp(vol))
    mats = [_random_matrix(vt, rng, shape, KINDS[int(rng.integers(0, len(KINDS)))]) for _ in range(5)]
    sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
    out = torch.full((len(mats),) + shape, -7.0, device='cuda:0')
    assert sv.affine_many(mats, output=out) is None
    proj = sv.project_many(mats).cpu().numpy()
    got = out.cpu().numpy()
    for k, m in enumerate(mats):
        want = oracle.affine(vol, m, mode, output=np.full(shape, -7.0, np.float32))
        e = float(np.abs(got[k].astype(np.float64) - want).max()) / r
        assert e <= 1e-6, f'seed {seed} {shape} matrix {k} {mode}: {e:.3e}'
        assert np.array_equal(got[k] == -7.0, want == -7.0), 'skipped voxels must keep the previous contents'
        psum = oracle.affine(vol, m, mode).astype(np.float64).sum(axis=0)
        ep = float(np.abs(proj[k] - psum).max()) / (r * shape[0])
        assert ep <= 2e-6, f'seed {seed} {shape} projection {k} {mode}: {ep:.3e}'
