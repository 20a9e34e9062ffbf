"""CPU suite: host-side logic (matrix builders, shape helpers, API surface) and the C-ABI library's exports."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle

ROOT = Path(__file__).resolve().parents[1]


def _utils():
    import voltools_b200.utils as u
    return u


# golden host matrices generated from the reference
G1 = np.array([[0.7712806, 0.6337184, 0.059391174, 0], [-0.613092, 0.71461016, 0.3368241, 0],
               [0.17101008, -0.29619813, 0.9396926, 0], [0, 0, 0, 1]], np.float32)
G2 = np.array([[0.81379765, 0.54383814, -0.20487413, 0], [-0.4698463, 0.8231729, 0.31879577, 0],
               [0.34202015, -0.16317591, 0.9254166, 0], [0, 0, 0, 1]], np.float32)
G6 = np.array([[1.3950913e-01, 8.3980620e-01, 3.8669834e-01, -9.9016479e+01],
               [-8.5836309e-01, -1.4925867e-01, 6.6490811e-01, 3.4631332e+02],
               [6.7360973e-01, -2.9064128e-01, 7.1574771e-01, -2.7221985e+01], [0, 0, 0, 1]], np.float32)


def test_golden_matrices():
    u = _utils()
    np.testing.assert_allclose(u.rotation_matrix((10, 20, 30), 'deg', 'rzxz'), G1, atol=1e-6)
    np.testing.assert_allclose(u.rotation_matrix((10, 20, 30), 'deg', 'sxyz'), G2, atol=1e-6)
    t = u.translation_matrix((1, 2, 3))
    assert np.array_equal(t[:3, 3], [-1, -2, -3]) and np.array_equal(t[:3, :3], np.identity(3))
    s = u.shear_matrix((0.1, 0.2, 0.3))
    assert s[0, 1] == np.float32(0.1) and s[0, 2] == np.float32(0.2) and s[1, 2] == np.float32(0.3)
    assert np.array_equal(np.diag(u.scale_matrix((2, 3, 4))), [2, 3, 4, 1])
    g6 = u.transform_matrix(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                            rotation_order='rzxz', translation=(5.5, -3.25, 2.0), center=(255.5,) * 3)
    np.testing.assert_allclose(g6, G6, rtol=1e-6, atol=1e-5)
    g7 = u.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=(49.5,) * 3)
    np.testing.assert_allclose(g7, [[1, 0, 0, 0], [0, 0.70710677, 0.70710677, -20.50357],
                                    [0, -0.70710677, 0.70710677, 49.5], [0, 0, 0, 1]], atol=1e-5)
    assert g7.dtype == np.float32
    pb, pa, nd = u.compute_post_transform_dimensions((100,) * 3, g7)
    assert tuple(pb) == (0, 21, 21) and tuple(pa) == (0, 20, 21) and tuple(nd) == (100, 141, 142)


def test_matrices_vs_imported_reference():
    """Every builder against the unmodified reference package, where it can be imported."""
    path = oracle.ref_python_path()
    if path is None:
        pytest.skip('reference python package not staged')
    saved = list(sys.path)
    sys.path.insert(0, path)
    try:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            import voltools.utils as ru
    finally:
        sys.path[:] = saved
    u = _utils()
    assert sorted(u.AVAILABLE_ROTATIONS) == sorted(ru.AVAILABLE_ROTATIONS)
    assert u.AVAILABLE_UNITS == ru.AVAILABLE_UNITS
    rng = np.random.default_rng(0)
    for order in ru.AVAILABLE_ROTATIONS:
        for units in ('deg', 'rad'):
            ang = rng.uniform(-180, 180, 3) if units == 'deg' else rng.uniform(-np.pi, np.pi, 3)
            np.testing.assert_allclose(u.rotation_matrix(ang, units, order), ru.rotation_matrix(ang, units, order),
                                       atol=1e-6)
            kw = dict(scale=rng.uniform(0.5, 2, 3), shear=rng.uniform(-0.2, 0.2, 3), rotation=ang,
                      rotation_units=units, rotation_order=order, translation=rng.uniform(-9, 9, 3),
                      center=rng.uniform(0, 200, 3))
            a, b = u.transform_matrix(**kw), ru.transform_matrix(**kw)
            assert a.dtype == b.dtype == np.float32
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-4)
    for shape in ((100, 100, 100), (30, 50, 70)):
        for _ in range(10):
            m = ru.transform_matrix(rotation=rng.uniform(-180, 180, 3), scale=rng.uniform(0.7, 1.5, 3),
                                    center=np.divide(shape, 2))
            for x, y in zip(u.compute_post_transform_dimensions(shape, m),
                            ru.compute_post_transform_dimensions(shape, m)):
                assert np.array_equal(x, y)


def test_api_surface():
    import voltools_b200 as vt
    for name in ('transform', 'affine', 'rotate', 'scale', 'shear', 'translate', 'StaticVolume',
                 'AVAILABLE_INTERPOLATIONS', 'AVAILABLE_DEVICES', 'utils'):
        assert hasattr(vt, name), name
    assert vt.AVAILABLE_INTERPOLATIONS == ['linear', 'bspline', 'bspline_simple', 'filt_bspline',
                                           'filt_bspline_simple']
    for name in ('transform_matrix', 'rotation_matrix', 'translation_matrix', 'shear_matrix', 'scale_matrix',
                 'AVAILABLE_ROTATIONS', 'AVAILABLE_UNITS', 'get_available_devices', 'switch_to_device',
                 'compute_post_transform_dimensions'):
        assert hasattr(vt.utils, name), name
    assert 'cpu' not in vt.AVAILABLE_DEVICES  # no CPU fallback in the product
    with pytest.raises(ValueError):
        vt.affine(np.zeros((4, 4, 4), np.float32), np.identity(4), device='cpu')
    with pytest.raises(ValueError):
        vt.utils.rotation_matrix((1, 2, 3), rotation_order='zxz')


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports everything include/voltools_b200.h declares (no compute calls)."""
    header = (ROOT / 'include' / 'voltools_b200.h').read_text()
    declared = set(re.findall(r'\b(vt_[a-z0-9_]+)\s*\(', header))
    declared -= {'vt_host_ctx'}
    assert {'vt_prefilter_f32', 'vt_affine_f32', 'vt_host_affine_f32', 'vt_error_string'} <= declared
    so = ROOT / 'voltools_b200' / 'libvoltools_b200.so'
    assert so.exists(), 'run python voltools_b200/csrc/build.py (or __graft_entry__.build())'
    lib = ctypes.CDLL(str(so))
    for name in sorted(declared):
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    lib.vt_abi_version.restype = ctypes.c_int
    from voltools_b200 import _native
    version = int(re.search(r'#define\s+VT_ABI_VERSION\s+(\d+)', header).group(1))
    assert lib.vt_abi_version() == version == _native.ABI_VERSION
    lib.vt_error_string.restype = ctypes.c_char_p
    assert lib.vt_error_string(0) == b'ok'
    assert lib.vt_error_string(2) != b'ok'


def test_product_does_not_touch_the_oracle():
    """No file of the product package may import/link/execute anything under oracle/."""
    for p in (ROOT / 'voltools_b200').rglob('*'):
        if p.suffix in ('.py', '.cu', '.cuh', '.h', '.cpp'):
            text = p.read_text()
            assert 'import oracle' not in text and 'from oracle' not in text and 'vt_oracle' not in text, p
            assert 'import scipy' not in text and 'from scipy' not in text, f'{p}: CPU fallback dependency'


def test_project_abi_argument_checks_without_a_gpu():
    """vt_project_* validate their arguments before touching CUDA; the workspace size is pure host arithmetic."""
    lib = ctypes.CDLL(str(ROOT / 'voltools_b200' / 'libvoltools_b200.so'))
    lib.vt_project_workspace_bytes.restype = ctypes.c_size_t
    lib.vt_project_workspace_bytes.argtypes = [ctypes.c_int] * 3
    assert lib.vt_project_workspace_bytes(0, 4, 4) == 0
    small, big = lib.vt_project_workspace_bytes(16, 32, 32), lib.vt_project_workspace_bytes(512, 512, 512)
    # 4 zero-bordered planes + 3 partial-sum planes per z-chunk
    assert small >= 4 * (32 + 4) * (32 + 4) * 4 + 3 * 32 * 32 * 4
    assert big > small and big < 64 << 20
    f32p, vp, i, ll = ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.vt_project_strided_f32.argtypes = [vp, i, i, i, ll, ll, vp, i, i, i, ll, f32p, i, i, ctypes.c_uint, i, i, vp,
                                           ctypes.c_size_t, i, vp]
    m = np.identity(4, dtype=np.float32)
    mp = m.ctypes.data_as(f32p)
    args = lambda **kw: [kw.get('src', 8), 4, 4, 4, 4, 16, kw.get('dst', 8), 4, 4, 4, kw.get('stride', 16), kw.get('m', mp),  # noqa: E731
                         kw.get('k', 1), kw.get('interp', 0), 0, 0, 4, None, 0, -1, None]
    assert lib.vt_project_strided_f32(*args(m=None)) == 1        # VT_ERR_INVALID_ARG: no matrices
    assert lib.vt_project_strided_f32(*args(k=-1)) == 1
    assert lib.vt_project_strided_f32(*args(interp=7)) == 1
    assert lib.vt_project_strided_f32(*args(stride=15)) == 1     # images would overlap
    assert lib.vt_project_strided_f32(*args(k=0)) == 0           # nothing to do
    assert lib.vt_project_strided_f32(*args(src=None)) == 1


def test_project_python_api_surface():
    import voltools_b200 as vt
    for name in ('project', 'affine_project', 'project_many'):
        assert callable(getattr(vt.StaticVolume, name))
    assert callable(vt.project)
    with pytest.raises(ValueError):
        vt.project(np.zeros((4, 4, 4), np.float32), interpolation='nearest')
    with pytest.raises(ValueError):
        vt.project(np.zeros((4, 4, 4), np.float32), device='cpu')


def test_build_script_loads_without_the_package():
    """__graft_entry__.build() must work on a clean checkout: the build script is loaded by path, because importing the
    package needs the shared library the build creates."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('vt_csrc_build_test', ROOT / 'voltools_b200' / 'csrc' / 'build.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert callable(mod.build) and mod.SO.name == 'libvoltools_b200.so'
    entry = (ROOT / '__graft_entry__.py').read_text()
    assert 'spec_from_file_location' in entry and 'from voltools_b200.csrc import build' not in entry


def test_slice_planner_decisions_without_a_gpu():
    """vt_slice_plan is host-only: pins the tuned launch heuristics of the slice family."""
    import voltools_b200 as vt
    from voltools_b200 import _native as N

    def rot(n, angle):
        c = np.divide(np.subtract((n, n, n), 1), 2, dtype=np.float32)
        return vt.utils.transform_matrix(rotation=(0, angle, 0), rotation_order='rzxz', center=c)

    # a general rotation is not of the slice family
    assert N.slice_plan((64, 64, 64), vt.utils.transform_matrix(rotation=(10, 20, 30), center=(31.5,) * 3), N.LINEAR) is None
    # long marches over 4 MB planes lose the L2 reuse between tiles: the linear kernel's marches are capped at 512 MB
    big = N.slice_plan((1024, 1024, 1024), rot(1024, 45), N.LINEAR)
    assert big['tma'] and big['chunks'] >= 8 and big['z_chunk'] * 4 * 1024 * 1024 <= 512 << 20
    assert N.slice_plan((1024, 1024, 1024), rot(1024, 45), N.CUBIC_TEX)['chunks'] <= 4   # on-chip bound: fewer start-ups
    assert N.slice_plan((512, 512, 512), rot(512, 45), N.LINEAR)['z_chunk'] <= 128
    # 0 and 90 degrees: 2 x 16 warps would put both rows on the same banks; a 4 x 8 / 8 x 4 shape is conflict free
    for angle in (0, 90):
        for interp in (N.CUBIC_TEX, N.CUBIC_SIMPLE):
            assert N.slice_plan((512, 512, 512), rot(512, angle), interp)['shapes'][0] in (1, 2), (angle, interp)
    # 45 degrees: no shape helps, the box is the 36-wide one (32 conflicts 3-way), 27 rows
    for n in (250, 256, 512):
        p = N.slice_plan((n, n, n), rot(n, 45), N.CUBIC_TEX)
        assert (p['box_w'], p['box_h'], p['shapes'][0]) == (36, 27, 0), (n, p)
    # a batch: one box width for the launch, a shape per matrix; planning is deterministic (memoised tables)
    mats = [rot(256, a) for a in range(0, 180, 6)]
    p1, p2 = N.slice_plan((256,) * 3, mats, N.CUBIC_SIMPLE), N.slice_plan((256,) * 3, mats, N.CUBIC_SIMPLE)
    assert p1 == p2 and len(p1['shapes']) == 30 and len(set(p1['pitches'])) == 1 and len(set(p1['shapes'])) > 1
    # strong in-plane magnification: narrow footprints, boxes down to 12 texels wide (the tables cover 12..40)
    c64 = np.divide(np.subtract((64,) * 3, 1), 2, dtype=np.float32)
    for sc, wmax in ((0.1, 12), (0.3, 16)):
        zoom = vt.utils.transform_matrix(rotation=(0, 20, 0), rotation_order='rzxz', scale=(1.0, sc, sc), center=c64)
        pz = N.slice_plan((64,) * 3, zoom, N.CUBIC_SIMPLE)
        assert pz['box_w'] % 4 == 0 and 12 <= pz['box_w'] <= 40 and pz['box_w'] >= wmax and pz['shapes'][0] in (0, 1, 2)
    # rows that are not multiples of 16 bytes: per-element cp.async staging, a pitch per matrix in 32..40
    odd = N.slice_plan((40, 50, 61), rot(50, 30), N.LINEAR, src_strides=(61, 50 * 61))
    assert not odd['tma'] and 32 <= odd['pitches'][0] <= 40


def test_slice4_planner_decisions_without_a_gpu():
    """vt_z4_axis_of / vt_z4_plan are host-only: which axis a batch of matrices leaves alone, and the per-matrix
    quarter-warp shape / pitch class the slice4 family picks from its bank simulation."""
    import voltools_b200 as vt
    from voltools_b200 import _native as N
    shape = (256, 256, 256)
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    tm = vt.utils.transform_matrix
    # the march axis: rotations about axis 0 ('rzxz' middle angle), 1 ('ryzy' first angle), 2 ('rzxz' first angle)
    assert N.z4_axis(shape, shape, tm(rotation=(0, 30, 0), rotation_order='rzxz', center=c), N.CUBIC_TEX) == 0
    assert N.z4_axis(shape, shape, tm(rotation=(30, 0, 0), rotation_order='ryzy', center=c), N.CUBIC_TEX) == 1
    assert N.z4_axis(shape, shape, tm(rotation=(25, 0, 0), rotation_order='rzxz', center=c), N.CUBIC_TEX) == 2  # the example's
    assert N.z4_axis(shape, shape, np.identity(4, dtype=np.float32), N.LINEAR) == 0                             # any: lowest
    assert N.z4_axis(shape, shape, tm(rotation=(10, 20, 30), center=c), N.LINEAR) == -1
    assert N.z4_axis(shape, shape, tm(rotation=(0, 30, 0), center=c, translation=(0.5, 0, 0)), N.LINEAR) == -1  # fractional shift
    assert N.z4_axis(shape, shape, tm(rotation=(0, 30, 0), center=c, translation=(3, 0.5, -2)), N.LINEAR) == 0  # integer shift
    # a batch must agree on one axis; a batch of more than VT_MAX_BATCH matrices is checked chunk by chunk
    sweep = [tm(rotation=(0, a, 0), rotation_order='rzxz', center=c) for a in range(180)]
    assert N.z4_axis(shape, shape, sweep, N.CUBIC_TEX) == 0
    assert N.z4_axis(shape, shape, sweep + [tm(rotation=(25, 0, 0), rotation_order='rzxz', center=c)], N.CUBIC_TEX) == -1
    assert N.launch_plan(None, shape, shape, sweep[45], N.CUBIC_TEX) == {'family': 'slice4', 'axis': 0}
    assert N.launch_plan(None, shape, shape, sweep[45], N.CUBIC_TEX, resident=False, filtered=False)['family'] == 'slice'
    assert N.launch_plan(None, shape, shape, sweep[45], N.CUBIC_TEX, resident=False, filtered=True)['family'] == 'slice4'
    # strong minification does not fit the 16 x 16 tile's footprint box: not this family
    assert N.z4_axis(shape, shape, tm(scale=(1.0, 2.5, 2.5), center=c), N.LINEAR) == -1
    # the plan: 0 and 90 degrees are conflict free with 1 x 8 patches; 45 degrees is the worst case (~1.4 wavefronts per
    # quarter-warp load instead of the ~2 of a full warp on the plain layout); box = footprint, at most 7 texels wider
    for axis, order, rot in ((0, 'rzxz', lambda a: (0, a, 0)), (1, 'ryzy', lambda a: (a, 0, 0)), (2, 'rzxz', lambda a: (a, 0, 0))):
        for angle in (0, 90):
            p = N.z4_plan(shape, tm(rotation=rot(angle), rotation_order=order, center=c), N.CUBIC_TEX, axis)
            assert p['wavefronts'][0] <= 1.01 and p['box_w'] <= 23, (axis, angle, p)
        p45 = N.z4_plan(shape, tm(rotation=rot(45), rotation_order=order, center=c), N.CUBIC_SIMPLE, axis)
        # (cubic kernels: 8-row tiles, footprint 0.707 * (7 + 15) + 6 = 21 texels; 16-row tiles: 27)
        assert 1.2 <= p45['wavefronts'][0] <= 1.6 and 21 <= p45['box_w'] <= 28 and p45['box_w'] <= p45['pitches'][0] <= p45['box_w'] + 7
    ps = N.z4_plan(shape, sweep[:32], N.CUBIC_TEX, 0)
    assert len(ps['shapes']) == 32 and all(0 <= s <= 3 for s in ps['shapes']) and ps['chunks'] == 1 and ps['m_chunk'] == 256
    assert N.z4_plan(shape, sweep[:32], N.CUBIC_TEX, 0) == ps                     # deterministic (memoised tables)
    wf = [w for b in range(0, 180, 32) for w in N.z4_plan(shape, sweep[b:b + 32], N.CUBIC_TEX, 0)['wavefronts']]
    assert np.mean(wf) <= 1.3 and max(wf) <= 1.8                                  # over the whole sweep
    # a general matrix is refused
    assert N.z4_plan(shape, tm(rotation=(10, 20, 30), center=c), N.LINEAR, 0) is None
    # one large volume alone: the march is cut into chunks so that the grid fills the GPU
    big = N.z4_plan((512, 512, 512), tm(rotation=(0, 45, 0), center=np.divide(np.subtract((512,) * 3, 1), 2, dtype=np.float32)),
                    N.CUBIC_TEX, 0)
    assert big['chunks'] >= 2 and big['m_chunk'] % 4 == 0
    # sizes of the layout: the march axis rounded up to a multiple of four
    assert N.z4_bytes((250, 30, 40), 0) == 252 * 30 * 40 * 4 and N.z4_bytes((250, 30, 41), 2) == 250 * 30 * 44 * 4


def test_slice4_routing_policy(monkeypatch):
    """_native.z4_wanted / launch_plan: which one-shot launches pay the Z4 pack pass.  Resident
    volumes and axes 1 / 2 always; axis 0 only when the prefilter writes the layout (filt_*) or when the rows are not a
    multiple of 16 bytes (250^3: the pack replaces the pad copy every TMA-staged kernel would need)."""
    import voltools_b200 as vt
    from voltools_b200 import _native as N
    monkeypatch.delenv('VT_Z4', raising=False)
    assert N.z4_wanted(N.CUBIC_TEX, True)
    assert N.z4_wanted(N.LINEAR, False, axis=2)
    assert not N.z4_wanted(N.LINEAR, False, axis=0, filtered=False, width=256)
    assert N.z4_wanted(N.LINEAR, False, axis=0, filtered=True, width=256)
    assert N.z4_wanted(N.LINEAR, False, axis=0, filtered=False, width=250)
    assert N.padded_row(250) == 252 and N.padded_row(256) == 256
    for n, family in ((250, 'slice4'), (256, 'slice')):
        shape = (n, n, n)
        m = vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=np.divide(np.subtract(shape, 1), 2, dtype=np.float32))
        assert N.launch_plan(None, shape, shape, m, N.LINEAR, resident=False)['family'] == family
    monkeypatch.setenv('VT_Z4', '0')
    assert not N.z4_wanted(N.CUBIC_TEX, True)
