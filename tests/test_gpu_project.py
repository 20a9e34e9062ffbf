"""GPU parity of the fused rotate-and-project path.

The reference has no projection function: examples/projections.py:20-26 transforms the volume and sums it over axis 0.
The checker is therefore `oracle.affine(...)` (the CPU restatement of the reference's GPU kernels) summed over axis 0 in
float64, and -- on the GPU box, when oracle/_ref is built -- the reference's own kernels summed the same way.

Tolerance: the projection is a sum of d0 voxels, each within the mode's tolerance (tests/test_gpu_parity.py) of the
reference; its error is stated against the projection's own scale d0 * range(sampled volume).  The fused path only
reorders float32 additions, so the observed error is ~1e-7 of that scale; asserted: 2e-6.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

MODES = ['linear', 'bspline', 'bspline_simple', 'filt_bspline', 'filt_bspline_simple']
TOL = 2e-6


@pytest.fixture(scope='module')
def vt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import voltools_b200 as vt
    return vt


def _center(shape):
    return np.divide(np.subtract(shape, 1), 2, dtype=np.float32)


def _scale(vol, mode):
    v = oracle.prefilter(vol) if mode.startswith('filt') else vol
    return float(np.ptp(v)) * vol.shape[0]


def _matrices(vt, shape):
    c = _center(shape)
    tm = vt.utils.transform_matrix
    return {
        # the reference example's tilt series: rotation about axis 0 (slice family -> plane sums + 2-D resample)
        'tilt_-60': tm(rotation=(-60, 0, 0), rotation_order='sxyz', center=c),
        'tilt_33': tm(rotation=(33, 0, 0), rotation_order='sxyz', center=c),
        'rot45_rzxz': tm(rotation=(0, 45, 0), rotation_order='rzxz', center=c),
        # same with an integer shift along axis 0 (slice family, planes shifted out of the volume are dropped)
        'tilt_shift': tm(rotation=(20, 0, 0), rotation_order='sxyz', center=c, translation=(3, 0.5, -1.25)),
        'identity': np.identity(4, dtype=np.float32),
        # general matrices: brick / gather kernels with register accumulation + atomic adds
        'rot_general': tm(rotation=(33.3, -71.0, 12.5), rotation_order='sxyz', center=c),
        'full_affine': tm(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                          rotation_order='rzxz', translation=(1.5, -0.75, 0.5), center=c),
    }


@pytest.mark.parametrize('shape', [(20, 24, 28), (33, 17, 45), (48, 64, 40)])
@pytest.mark.parametrize('mode', MODES)
def test_project_vs_oracle(vt, shape, mode):
    rng = np.random.default_rng(sum(shape))
    vol = rng.random(shape, dtype=np.float32)
    sc = _scale(vol, mode)
    sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
    mats = _matrices(vt, shape)
    got_all = sv.project_many(list(mats.values())).cpu().numpy()
    assert got_all.shape == (len(mats),) + shape[1:] and got_all.dtype == np.float32
    for k, (name, m) in enumerate(mats.items()):
        want = oracle.affine(vol, m, mode).astype(np.float64).sum(axis=0)
        e = float(np.abs(got_all[k] - want).max()) / sc
        assert e <= TOL, f'{mode} {name} {shape}: {e:.3e}'
        one = sv.affine_project(m)
        assert float(np.abs(one - want).max()) / sc <= TOL, f'{mode} {name} {shape} (single)'


@pytest.mark.parametrize('mode', ['linear', 'filt_bspline', 'bspline_simple'])
def test_project_matches_transform_then_sum(vt, mode):
    """The fused path against this library's own transform followed by torch's sum, at a size where the slice family
    runs its TMA ring (256^2 planes), for a tilt and for a general matrix, plus the texture family once resident."""
    import torch
    shape = (96, 256, 256)
    vol = np.random.default_rng(3).random(shape, dtype=np.float32)
    sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
    c = _center(shape)
    tm = vt.utils.transform_matrix
    mats = [tm(rotation=(a, 0, 0), rotation_order='sxyz', center=c) for a in (-60, -3, 45)]
    mats.append(tm(rotation=(30, 45, 60), rotation_order='rzxz', translation=(1.5, -0.75, 0.5), center=c))
    full = sv.affine_many(mats)
    want = full.double().sum(dim=1)
    sc = float(full.max() - full.min()) * shape[0]
    got = sv.project_many(mats)
    assert float((got.double() - want).abs().max()) / sc <= TOL
    # output= variant and the texture family (general matrix, linear / cubic_tex only)
    out = torch.full(shape[1:], 7.0, device='cuda:0')
    assert sv.affine_project(mats[3], output=out) is None
    assert float((out.double() - want[3]).abs().max()) / sc <= TOL
    if mode != 'bspline_simple':
        sv._tex = vt._native.Texture(sv.coefficient_buffer.data_ptr(), sv.shape, sv._strides, 0,
                                     torch.cuda.current_stream().cuda_stream)
        got_tex = sv.project_many(mats)
        assert float((got_tex.double() - want).abs().max()) / sc <= TOL


def test_project_z_slabs_add_up(vt):
    """Partial projections of output z-slabs (multi-GPU sharding of one projection) sum to the whole."""
    shape = (50, 40, 44)
    vol = np.random.default_rng(5).random(shape, dtype=np.float32)
    c = _center(shape)
    tm = vt.utils.transform_matrix
    for mode in ('linear', 'filt_bspline_simple', 'filt_bspline'):
        sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
        for m in (tm(rotation=(25, 0, 0), rotation_order='sxyz', center=c, translation=(-2, 0, 0)),
                  tm(rotation=(30, 45, 60), rotation_order='rzxz', center=c)):
            whole = sv.project_many([m]).double()
            parts = sum(sv.project_many([m], z_range=(z0, z1)).double() for z0, z1 in ((0, 13), (13, 14), (14, 50)))
            sc = _scale(vol, mode)
            assert float((whole - parts).abs().max()) / sc <= TOL, mode


def test_project_vs_reference_kernels(vt):
    """The reference's own kernels (hardware texture) on this GPU, summed over axis 0 as examples/projections.py does."""
    if not oracle.ref_gpu_available():
        pytest.skip('oracle/_ref/libvt_ref_gpu.so not built')
    shape = (40, 52, 60)
    vol = np.random.default_rng(7).random(shape, dtype=np.float32)
    c = _center(shape)
    for mode in MODES:
        sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
        for angle in (-60, 12):
            m = vt.utils.transform_matrix(rotation=(angle, 0, 0), rotation_order='sxyz', center=c)
            want = oracle.transform_ref_gpu(vol, m, mode)[0].astype(np.float64).sum(axis=0)
            got = sv.affine_project(m)
            # per-voxel tolerance of the mode (2e-3 for the 8-bit-weight modes) -- observed ~1e-7
            assert float(np.abs(got - want).max()) / _scale(vol, mode) <= 1e-5, (mode, angle)


def test_module_level_project(vt):
    shape = (24, 30, 36)
    vol = np.random.default_rng(9).random(shape, dtype=np.float32)
    got = vt.project(vol, rotation=(15, 0, 0), rotation_order='sxyz', interpolation='filt_bspline', device='gpu:0')
    m = vt.utils.transform_matrix(rotation=(15, 0, 0), rotation_order='sxyz', center=_center(shape))
    want = oracle.affine(vol, m, 'filt_bspline').astype(np.float64).sum(axis=0)
    assert got.shape == shape[1:]
    assert float(np.abs(got - want).max()) / _scale(vol, 'filt_bspline') <= TOL


def test_host_path_from_several_threads(vt):
    """transform() with numpy in / numpy out from concurrent host threads (one host context per thread, uploads chained
    per device): every result equals the single-threaded one bit for bit."""
    from concurrent.futures import ThreadPoolExecutor
    shape = (136, 60, 68)
    rng = np.random.default_rng(11)
    vols = [rng.random(shape, dtype=np.float32) for _ in range(6)]
    kw = dict(rotation=(0, 33, 0), rotation_order='rzxz', interpolation='filt_bspline', device='gpu:0')
    want = [vt.transform(v, **kw) for v in vols]
    with ThreadPoolExecutor(3) as pool:
        for _ in range(3):
            got = list(pool.map(lambda v: vt.transform(v, **kw), vols))
            for g, w in zip(got, want):
                assert np.array_equal(g, w)
