"""GPU parity: libvoltools_b200 (through the public API / C ABI) vs the CPU oracle and vs the reference's own
CUDA kernels run through a real texture object (oracle/_ref, when present).

Tolerances (max |d| / value range of the SAMPLED volume, i.e. of the coefficient volume for filt_*):
  north_star:  bspline_simple, filt_bspline_simple, prefilter 1e-5 (pure float32 in the reference);
               linear, bspline, filt_bspline 2e-3 (the reference uses 8-bit hardware weights).
  asserted:    TOL = 5e-6 for EVERY mode, against the oracle and against the live reference kernels: the kernels
               reproduce the texture unit's integer weight rule, so the hardware-weight modes agree to float32 rounding
               (observed <= 3e-7) -- a regression to exact float weights (~1e-3 off) must fail, not pass.
"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

MODES = ['linear', 'bspline', 'bspline_simple', 'filt_bspline', 'filt_bspline_simple']
TOL = {'linear': 5e-6, 'bspline': 5e-6, 'filt_bspline': 5e-6, 'bspline_simple': 5e-6, 'filt_bspline_simple': 5e-6}


@pytest.fixture(scope='module')
def vt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import voltools_b200 as vt
    return vt


def _center(shape):
    return np.divide(np.subtract(shape, 1), 2, dtype=np.float32)


def _matrices(vt, shape):
    c = _center(shape)
    tm = vt.utils.transform_matrix
    return {
        'identity': np.identity(4, dtype=np.float32),
        'shift': tm(translation=(0.37, -1.21, 2.5)),
        'rot45': tm(rotation=(0, 45, 0), rotation_order='rzxz', center=c),
        'rot_general': tm(rotation=(33.3, -71.0, 12.5), rotation_order='sxyz', center=c),
        'full_affine': tm(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                          rotation_order='rzxz', translation=(1.5, -0.75, 0.5), center=c),
        'downscale': tm(scale=(2.3, 1.7, 3.1), center=c),
        'rotate_origin': vt.utils.rotation_matrix((10, 20, 30)),
    }


def _range(vol, mode):
    v = oracle.prefilter(vol) if mode.startswith('filt') else vol
    return float(np.ptp(v))


def _err(a, b, rng):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / rng


@pytest.mark.parametrize('shape', [(20, 24, 28), (33, 17, 45), (64, 64, 64)])
@pytest.mark.parametrize('mode', MODES)
def test_affine_vs_oracle(vt, shape, mode):
    rng = np.random.default_rng(sum(shape))
    vol = rng.random(shape, dtype=np.float32)
    r = _range(vol, mode)
    for name, m in _matrices(vt, shape).items():
        got = vt.affine(vol, m, interpolation=mode, device='gpu:0')
        want = oracle.affine(vol, m, mode)
        assert got.shape == vol.shape and got.dtype == np.float32
        e = _err(got, want, r)
        assert e <= TOL[mode], f'{mode} {name} {shape}: {e:.3e}'
        # the set of skipped voxels must be identical, not just close: read it off a sentinel the output is prefilled
        # with (with signed coefficients a result can be exactly 0.0, and not in the same voxel for two implementations)
        import torch
        out = torch.full(shape, -7.0, device='cuda')
        vt.affine(torch.from_numpy(vol).cuda(), m, interpolation=mode, output=out, device='gpu:0')
        want_s = oracle.affine(vol, m, mode, output=np.full(shape, -7.0, np.float32))
        assert np.array_equal(out.cpu().numpy() == -7.0, want_s == -7.0), f'{mode} {name} {shape}: skipped voxel sets differ'


@pytest.mark.parametrize('mode', MODES)
def test_affine_vs_reference_kernels(vt, mode):
    """Same inputs through the reference's own kernels + hardware texture on this GPU."""
    if not oracle.ref_gpu_available():
        pytest.skip('oracle/_ref/libvt_ref_gpu.so not built')
    shape = (40, 52, 60)  # all even: the reference's prefilter launch needs power-of-two divisors
    rng = np.random.default_rng(7)
    vol = rng.random(shape, dtype=np.float32)
    r = _range(vol, mode)
    for name, m in _matrices(vt, shape).items():
        ref, _, _ = oracle.transform_ref_gpu(vol, m, mode)
        got = vt.affine(vol, m, interpolation=mode, device='gpu:0')
        want = oracle.affine(vol, m, mode)
        e_mine, e_oracle = _err(got, ref, r), _err(want, ref, r)
        assert e_oracle <= TOL[mode], f'oracle vs reference {mode} {name}: {e_oracle:.3e}'
        assert e_mine <= TOL[mode], f'ours vs reference {mode} {name}: {e_mine:.3e}'


@pytest.mark.parametrize('shape', [(16, 16, 16), (20, 24, 28), (7, 5, 3), (1, 1, 1), (33, 17, 45), (50, 60, 130),
                                   (3, 200, 11), (70, 9, 300), (40, 130, 33)])
def test_prefilter(vt, shape):
    import torch
    rng = np.random.default_rng(3)
    vol = rng.random(shape, dtype=np.float32)
    want = oracle.prefilter(vol)
    r = float(np.ptp(want)) or 1.0
    st = torch.cuda.current_stream().cuda_stream
    # variant 1 (sequential, reference operation order), in place and out of place
    t = torch.from_numpy(vol).cuda()
    vt._native.prefilter(t.data_ptr(), shape, 0, st, variant=1)
    seq = t.cpu().numpy()
    assert _err(seq, want, r) <= 1e-6, (shape, _err(seq, want, r))
    src = torch.from_numpy(vol).cuda()
    dst = torch.full(shape, np.nan, device='cuda')
    vt._native.prefilter(src.data_ptr(), shape, 0, st, variant=1, dst_ptr=dst.data_ptr())
    assert np.array_equal(dst.cpu().numpy(), seq) and np.array_equal(src.cpu().numpy(), vol)
    # variant 0 out of place: windowed kernels (truncated warm-up, |pole|^12 = 1.4e-7)
    dst = torch.full(shape, np.nan, device='cuda')
    vt._native.prefilter(src.data_ptr(), shape, 0, st, variant=0, dst_ptr=dst.data_ptr())
    got = dst.cpu().numpy()
    assert np.array_equal(src.cpu().numpy(), vol), 'source modified'
    assert _err(got, want, r) <= 1e-6, ('windowed', shape, _err(got, want, r))
    # variant 0 into a buffer with rows padded to 16 bytes (what the filt_* paths use): pad columns are zero
    row = vt._native.padded_row(shape[2]) + 4
    pad = torch.full((shape[0], shape[1], row), np.nan, device='cuda')
    vt._native.prefilter(src.data_ptr(), shape, 0, st, variant=0, dst_ptr=pad.data_ptr(),
                         dst_strides=(row, shape[1] * row))
    pad = pad.cpu().numpy()
    assert np.array_equal(pad[:, :, :shape[2]], got) and np.all(pad[:, :, shape[2]:] == 0)
    # variant 0 in place falls back to the sequential kernels
    t = torch.from_numpy(vol).cuda()
    vt._native.prefilter(t.data_ptr(), shape, 0, st, variant=0)
    assert np.array_equal(t.cpu().numpy(), seq)
    if oracle.ref_gpu_available() and all(s % 2 == 0 for s in shape):
        ref = oracle.prefilter_ref_gpu(vol)
        assert _err(want, ref, r) <= 1e-6
        assert np.array_equal(seq, ref), 'sequential variant must be bit-identical to the reference'


def test_prefilter_windows_large(vt):
    """Sizes at which the windowed kernels really cut lines into strips / steps (several y-strips, many z steps)."""
    import torch
    st = torch.cuda.current_stream().cuda_stream
    for shape in ((100, 300, 250), (64, 120, 512)):
        src = torch.rand(shape, device='cuda', generator=torch.Generator('cuda').manual_seed(5))
        a = src.clone()
        vt._native.prefilter(a.data_ptr(), shape, 0, st, variant=1)
        b = torch.empty_like(src)
        vt._native.prefilter(src.data_ptr(), shape, 0, st, variant=0, dst_ptr=b.data_ptr())
        r = float(a.max() - a.min())
        assert float((a - b).abs().max()) / r <= 1e-6, shape
        # interior reconstruction: (c[i-1] + 4 c[i] + c[i+1]) / 6 == s[i] along every axis
        rec = b
        for ax in range(3):
            rec = (rec.roll(1, ax) + 4 * rec + rec.roll(-1, ax)) / 6
        inner = (slice(16, -16),) * 3
        assert float((rec[inner] - src[inner]).abs().max()) <= 2e-5


@pytest.mark.parametrize('shape,chunks', [((100, 40, 52), 3), ((64, 33, 70), 8), ((150, 20, 250), 5), ((13, 16, 18), 3)])
def test_prefilter_streaming(vt, shape, chunks):
    """vt_prefilter_planes_f32 fed chunk by chunk (the multi-GPU layer's pipelined prepare, with its own plan) gives
    the coefficients of the one-shot prefilter, including chunk boundaries inside the first 12 planes."""
    import torch
    from voltools_b200 import multigpu
    N = vt._native
    st = torch.cuda.current_stream().cuda_stream
    src = torch.rand(shape, device='cuda', generator=torch.Generator('cuda').manual_seed(11))
    row = N.padded_row(shape[2])
    strides = (row, shape[1] * row)
    ws = torch.full((shape[0], shape[1], row), float('nan'), device='cuda')   # planes not yet produced are poison
    got = torch.full((shape[0], shape[1], row), float('nan'), device='cuda')
    for xy0, xy1, z0, z1 in multigpu.stream_plan(shape[0], True, chunks):
        N.prefilter_planes(src.data_ptr(), ws.data_ptr(), got.data_ptr(), shape, strides, (xy0, xy1), (z0, z1), 0, st)
    ref = src.clone()
    N.prefilter(ref.data_ptr(), shape, 0, st, variant=1)
    r = float(ref.max() - ref.min())
    assert not bool(torch.isnan(got).any())
    assert float((got[:, :, :shape[2]] - ref).abs().max()) / r <= 1e-6
    assert float(got[:, :, shape[2]:].abs().max() if row > shape[2] else 0.0) == 0.0


def _slice_matrices(vt, shape):
    c = _center(shape)
    tm = vt.utils.transform_matrix
    mats = {f'rot{a}': tm(rotation=(0, a, 0), rotation_order='rzxz', center=c) for a in (0, 1, 45, 90, 133.7, 180, -60)}
    mats['rot30_shift'] = tm(rotation=(0, 30, 0), rotation_order='rzxz', center=c, translation=(3, 1.25, -2.5))
    mats['rot30_shift_neg'] = tm(rotation=(0, -30, 0), rotation_order='rzxz', center=c, translation=(-5, 0.5, 0.75))
    mats['inplane_scale'] = tm(scale=(1.0, 1.15, 0.9), center=c)
    mats['origin_rotate'] = vt.utils.rotation_matrix((0, 20, 0), rotation_order='rzxz')
    return mats


@pytest.mark.parametrize('shape', [(20, 24, 28), (33, 47, 45), (9, 70, 130)])
@pytest.mark.parametrize('interp', [0, 1, 2])
def test_slice_family(vt, shape, interp):
    """Matrices that leave axis 0 alone run on the plane-marching kernels: same results as the general gather
    kernels (float32 summation order aside) and as the oracle."""
    import torch
    N = vt._native
    rng = np.random.default_rng(17)
    vol_np = rng.random(shape, dtype=np.float32)
    vol = torch.from_numpy(vol_np).cuda()
    mode = ['linear', 'bspline', 'bspline_simple'][interp]
    for name, m in _slice_matrices(vt, shape).items():
        assert N.affine_plan(vol.data_ptr(), shape, shape, m, interp) == 'slice', name
        for flag in (N.OOB_ZERO, N.OOB_SKIP):
            a = torch.full(shape, -7.0, device='cuda')
            b = torch.full(shape, -7.0, device='cuda')
            N.affine(vol.data_ptr(), shape, a.data_ptr(), shape, m, interp, flag | N.KERNEL_SLICE)
            N.affine(vol.data_ptr(), shape, b.data_ptr(), shape, m, interp, flag | N.KERNEL_GATHER)
            a, b = a.cpu().numpy(), b.cpu().numpy()
            assert np.array_equal(a == -7.0, b == -7.0), (name, 'skipped sets differ')
            assert _err(a, b, 1.0) <= 1e-6, (name, _err(a, b, 1.0))
        want = oracle.affine(vol_np, m, mode)
        assert _err(a if flag == N.OOB_ZERO else np.where(a == -7.0, 0, a), want, 1.0) <= 1e-6, name
    # staging variants (TMA box loads vs per-element cp.async) are the same arithmetic: bit-identical; the TMA
    # variant needs 16-byte aligned rows, so run it on a padded copy of the volume
    row = N.padded_row(shape[2])
    padded = torch.zeros((shape[0], shape[1], row), device='cuda')
    padded[:, :, :shape[2]] = vol
    for name in ('rot45', 'rot30_shift', 'inplane_scale'):
        m = _slice_matrices(vt, shape)[name]
        a = torch.zeros(shape, device='cuda')
        b = torch.zeros(shape, device='cuda')
        N.affine(padded.data_ptr(), shape, a.data_ptr(), shape, m, interp, N.OOB_ZERO | N.KERNEL_SLICE,
                 src_strides=(row, shape[1] * row))
        N.affine(padded.data_ptr(), shape, b.data_ptr(), shape, m, interp,
                 N.OOB_ZERO | N.KERNEL_SLICE | N.STAGE_CP_ASYNC, src_strides=(row, shape[1] * row))
        assert torch.equal(a, b), name
        c = torch.zeros(shape, device='cuda')
        N.affine(vol.data_ptr(), shape, c.data_ptr(), shape, m, interp, N.OOB_ZERO | N.KERNEL_GATHER)
        assert float((a - c).abs().max()) <= 1e-6, name
    # a matrix with a fractional offset along axis 0 must NOT take the slice path
    m = vt.utils.transform_matrix(rotation=(0, 30, 0), rotation_order='rzxz', center=_center(shape),
                                  translation=(0.5, 0, 0))
    assert N.affine_plan(vol.data_ptr(), shape, shape, m, interp) != 'slice'
    # batched: several angles in one launch, z-slab restricted
    mats = [vt.utils.transform_matrix(rotation=(0, a, 0), center=_center(shape)) for a in range(0, 180, 9)]
    out = torch.zeros((len(mats),) + shape, device='cuda')
    N.affine(vol.data_ptr(), shape, out.data_ptr(), shape, mats, interp, N.OOB_ZERO | N.KERNEL_SLICE, z_range=(2, 7))
    ref = torch.zeros((len(mats),) + shape, device='cuda')
    N.affine(vol.data_ptr(), shape, ref.data_ptr(), shape, mats, interp, N.OOB_ZERO | N.KERNEL_GATHER, z_range=(2, 7))
    assert float((out - ref).abs().max()) <= 1e-6
    assert float(out[:, :2].abs().max()) == 0 and float(out[:, 7:].abs().max()) == 0


@pytest.mark.parametrize('shape', [(20, 24, 28), (33, 47, 44), (9, 70, 132), (64, 64, 64)])
@pytest.mark.parametrize('interp', [0, 1, 2])
def test_brick_family(vt, shape, interp):
    """General matrices on 16-byte aligned rows run on the TMA-staged brick kernels: equal to the gather kernels up to
    float32 summation order (<= 1e-6), identical skipped-voxel sets; also against the oracle."""
    import torch
    N = vt._native
    rng = np.random.default_rng(23)
    vol_np = rng.random(shape, dtype=np.float32)
    vol = torch.from_numpy(vol_np).cuda()
    mode = ['linear', 'bspline', 'bspline_simple'][interp]
    mats = _matrices(vt, shape)
    c = _center(shape)
    mats['big_shift'] = vt.utils.transform_matrix(rotation=(40, 20, 10), translation=(30, -25, 12), center=c)
    mats['upscale'] = vt.utils.transform_matrix(scale=(0.5, 0.4, 0.7), rotation=(10, 50, 80), center=c)
    for name, m in mats.items():
        fam = N.affine_plan(vol.data_ptr(), shape, shape, m, interp)
        if name in ('identity', 'shift', 'rot45'):
            continue  # leave axis 0 alone (or have a fractional axis-0 shift): covered elsewhere
        if name == 'downscale':
            assert fam in ('brick', 'gather')  # strong minification may not fit a brick: the gather family runs
        else:
            assert fam == 'brick', (name, fam)
        if fam != 'brick':
            continue
        for flag in (N.OOB_ZERO, N.OOB_SKIP):
            a = torch.full(shape, -7.0, device='cuda')
            b = torch.full(shape, -7.0, device='cuda')
            N.affine(vol.data_ptr(), shape, a.data_ptr(), shape, m, interp, flag | N.KERNEL_BRICK)
            N.affine(vol.data_ptr(), shape, b.data_ptr(), shape, m, interp, flag | N.KERNEL_GATHER)
            # same weights as the gather kernels (the float-pipeline form of the texture rule is exact); the float32
            # summation order differs (z-near / z-far halves in FFMA2 pairs; separable sums for cubic_simple)
            assert float((a - b).abs().max()) <= 1e-6, (name, float((a - b).abs().max()))
            assert torch.equal(a == -7.0, b == -7.0), name
        want = oracle.affine(vol_np, m, mode)
        got = np.where(a.cpu().numpy() == -7.0, 0, a.cpu().numpy())
        assert _err(got, want, 1.0) <= 1e-6, name
    # batch + z-slab
    ms = [vt.utils.transform_matrix(rotation=(a, 2 * a, 30), center=c) for a in range(5, 60, 11)]
    out = torch.zeros((len(ms),) + shape, device='cuda')
    ref = torch.zeros((len(ms),) + shape, device='cuda')
    N.affine(vol.data_ptr(), shape, out.data_ptr(), shape, ms, interp, N.OOB_ZERO | N.KERNEL_BRICK, z_range=(3, 8))
    N.affine(vol.data_ptr(), shape, ref.data_ptr(), shape, ms, interp, N.OOB_ZERO | N.KERNEL_GATHER, z_range=(3, 8))
    assert float((out - ref).abs().max()) <= 1e-6
    # unaligned rows cannot be staged with TMA
    if shape[2] % 4 == 0:
        odd = (shape[0], shape[1], shape[2] - 1)
        v2 = torch.zeros(odd, device='cuda')
        assert N.affine_plan(v2.data_ptr(), odd, odd, mats['rot_general'], interp) == 'gather'


@pytest.mark.parametrize('mode', MODES)
def test_host_pipeline(vt, mode):
    """numpy in -> numpy out on volumes deep enough for the chunked upload/prefilter/resample/download pipeline
    (>= 64 planes: several chunks): same result as the device-resident path and as the oracle, for matrices that
    can stream (axis-0 rotations, with and without an integer shift along axis 0) and for one that cannot."""
    import torch
    shape = (100, 40, 52)
    rng = np.random.default_rng(23)
    vol = rng.random(shape, dtype=np.float32)
    r = _range(vol, mode)
    c = _center(shape)
    tm = vt.utils.transform_matrix
    mats = {'rot45': tm(rotation=(0, 45, 0), rotation_order='rzxz', center=c),
            'rot30_shift_z+7': tm(rotation=(0, 30, 0), rotation_order='rzxz', center=c, translation=(7, 1.5, -2)),
            'rot30_shift_z-9': tm(rotation=(0, 30, 0), rotation_order='rzxz', center=c, translation=(-9, 0, 0.25)),
            'full_affine': _matrices(vt, shape)['full_affine']}
    dvol = torch.from_numpy(vol).cuda()
    for name, m in mats.items():
        got = vt.affine(vol, m, interpolation=mode, device='gpu:0')          # host path (pipelined)
        dev = torch.zeros(shape, device='cuda')
        vt.affine(dvol, m, interpolation=mode, output=dev, device='gpu:0')   # device path
        want = oracle.affine(vol, m, mode)
        assert _err(got, want, r) <= TOL[mode], (mode, name, _err(got, want, r))
        assert _err(got, dev.cpu().numpy(), r) <= 1e-6, (mode, name, _err(got, dev.cpu().numpy(), r))
    # numpy output= buffer: voxels outside the source are written as zeros on this path too
    out = np.full(shape, 5.0, dtype=np.float32)
    vt.affine(vol, mats['rot45'], interpolation=mode, output=out, device='gpu:0')
    assert _err(out, oracle.affine(vol, mats['rot45'], mode), r) <= TOL[mode]


def test_output_semantics(vt):
    """output= given: written in place, out-of-bounds voxels keep their contents, returns None
    (transforms.py:207-210, :224-226); input arrays are never modified."""
    import torch
    shape = (24, 28, 32)
    rng = np.random.default_rng(11)
    vol = rng.random(shape, dtype=np.float32)
    m = _matrices(vt, shape)['rot45']
    prefill = (rng.random(shape, dtype=np.float32) + 10).astype(np.float32)
    for mode in ('linear', 'filt_bspline', 'filt_bspline_simple'):
        out = torch.from_numpy(prefill.copy()).cuda()
        src = torch.from_numpy(vol).cuda()
        ret = vt.affine(src, m, interpolation=mode, output=out, device='gpu')
        assert ret is None
        assert np.array_equal(src.cpu().numpy(), vol), 'input was modified'
        want = oracle.affine(vol, m, mode, output=prefill)
        got = out.cpu().numpy()
        assert _err(got, want, _range(vol, mode)) <= TOL[mode]
        kept = want == prefill
        assert kept.any() and np.array_equal(got[kept], prefill[kept])


def test_static_volume_and_batch(vt):
    import torch
    shape = (30, 36, 40)
    rng = np.random.default_rng(5)
    vol = rng.random(shape, dtype=np.float32)
    for mode in ('linear', 'filt_bspline', 'bspline_simple'):
        sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
        r = _range(vol, mode)
        mats = [vt.utils.transform_matrix(rotation=(0, a, 0), center=_center(shape)) for a in range(0, 180, 4)]
        single = sv.transform(rotation=(0, 8, 0))
        want = oracle.affine(vol, mats[2], mode)
        assert _err(single, want, r) <= TOL[mode]
        batch = sv.affine_many(mats).cpu().numpy()
        assert batch.shape == (len(mats),) + shape
        for k in (0, 2, 17, len(mats) - 1):
            assert _err(batch[k], oracle.affine(vol, mats[k], mode), r) <= TOL[mode]
        assert np.array_equal(batch[2], single), 'batched and single launches must agree bit for bit'
        out = torch.zeros(shape, device='cuda:0')
        assert sv.rotate((0, 0, 10), output=out) is None
        assert _err(out.cpu().numpy(), oracle.affine(vol, vt.utils.rotation_matrix((0, 0, 10)), mode), r) <= TOL[mode]


@pytest.mark.parametrize('shape', [(20, 24, 28), (33, 47, 45), (64, 64, 64)])
@pytest.mark.parametrize('interp', [0, 1])
def test_texture_family(vt, shape, interp):
    """General matrices through a hardware texture object (vt_tex_* of the C ABI): the unit's own arithmetic, so
    the oracle's texture model, the software-emulating kernels and this path must all agree to float32 rounding;
    OOB policies, padded sources, batches and z-ranges behave as on the other families."""
    import torch
    N = vt._native
    rng = np.random.default_rng(31)
    vol_np = rng.random(shape, dtype=np.float32)
    row = N.padded_row(shape[2])
    padded = torch.zeros((shape[0], shape[1], row), device='cuda')
    padded[:, :, :shape[2]] = torch.from_numpy(vol_np).cuda()
    strides = (row, shape[1] * row)
    tex = N.Texture(padded.data_ptr(), shape, strides, 0, torch.cuda.current_stream().cuda_stream)
    mode = ['linear', 'bspline'][interp]
    mats = _matrices(vt, shape)
    for name, m in mats.items():
        for flag in (N.OOB_ZERO, N.OOB_SKIP):
            a = torch.full(shape, -7.0, device='cuda')
            b = torch.full(shape, -7.0, device='cuda')
            tex.affine(a.data_ptr(), shape, m, interp, flag)
            N.affine(padded.data_ptr(), shape, b.data_ptr(), shape, m, interp, flag | N.KERNEL_GATHER, src_strides=strides)
            a, b = a.cpu().numpy(), b.cpu().numpy()
            assert np.array_equal(a == -7.0, b == -7.0), (name, 'skipped sets differ')
            assert _err(a, b, 1.0) <= 1e-6, (name, _err(a, b, 1.0))
        want = oracle.affine(vol_np, m, mode)
        assert _err(np.where(a == -7.0, 0, a), want, 1.0) <= 1e-6, name
    # batch + z-range
    ms = [mats['rot_general'], mats['full_affine'], mats['downscale']]
    out = torch.zeros((3,) + shape, device='cuda')
    ref = torch.zeros((3,) + shape, device='cuda')
    tex.affine(out.data_ptr(), shape, ms, interp, N.OOB_ZERO, z_range=(3, 11))
    N.affine(padded.data_ptr(), shape, ref.data_ptr(), shape, ms, interp, N.OOB_ZERO | N.KERNEL_GATHER, z_range=(3, 11),
             src_strides=strides)
    assert float((out - ref).abs().max()) <= 1e-6
    assert float(out[:, :3].abs().max()) == 0 and float(out[:, 11:].abs().max()) == 0
    # re-upload of another volume of the same shape
    vol2 = torch.rand(shape, device='cuda')
    tex.upload(vol2.data_ptr())
    tex.affine(out[0].data_ptr(), shape, ms[0], interp, N.OOB_ZERO)
    N.affine(vol2.data_ptr(), shape, ref[0].data_ptr(), shape, ms[0], interp, N.OOB_ZERO | N.KERNEL_GATHER)
    assert float((out[0] - ref[0]).abs().max()) <= 1e-6
    tex.close()
    # cubic_simple has no texture path
    t2 = N.Texture(vol2.data_ptr(), shape)
    with pytest.raises(RuntimeError):
        t2.affine(out[0].data_ptr(), shape, ms[0], 2, N.OOB_ZERO)


def test_static_volume_switches_to_texture(vt):
    """A StaticVolume that keeps getting general matrices moves to the texture family after TEXTURE_AFTER of them;
    results before and after the switch agree to float32 rounding, slice-family matrices never switch."""
    import torch
    shape = (36, 40, 44)
    vol = np.random.default_rng(9).random(shape, dtype=np.float32)
    c = _center(shape)
    for mode in ('linear', 'filt_bspline'):
        sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
        r = _range(vol, mode)
        mats = [vt.utils.transform_matrix(rotation=(10 + a, 20 + 2 * a, 30 - a), rotation_order='rzxz', center=c)
                for a in range(sv.TEXTURE_AFTER + 4)]
        first = sv.affine(mats[3])
        assert getattr(sv, '_tex', None) is None
        batch = sv.affine_many(mats).cpu().numpy()     # crosses the threshold: texture family
        assert getattr(sv, '_tex', None) is not None
        assert _err(batch[3], first, r) <= 1e-6
        for k in (0, 7, len(mats) - 1):
            assert _err(batch[k], oracle.affine(vol, mats[k], mode), r) <= TOL[mode]
        again = sv.affine(mats[3])                      # single launches use it too from now on
        assert np.array_equal(again, batch[3])
        rot = sv.transform(rotation=(0, 30, 0))         # slice family: unaffected
        assert _err(rot, oracle.affine(vol, vt.utils.transform_matrix(rotation=(0, 30, 0), center=c), mode), r) <= TOL[mode]
    sv = vt.StaticVolume(vol, interpolation='bspline_simple', device='gpu:0')
    sv.affine_many([vt.utils.transform_matrix(rotation=(10, 20, 30 + a), center=c) for a in range(20)])
    assert getattr(sv, '_tex', None) is None


def test_z_slabs_compose(vt):
    """z-slab sharding: disjoint slabs written by separate calls reproduce the single-call result exactly."""
    import torch
    shape = (37, 20, 24)
    rng = np.random.default_rng(9)
    vol = torch.from_numpy(rng.random(shape, dtype=np.float32)).cuda()
    m = _matrices(vt, shape)['full_affine']
    for interp in (0, 1, 2):
        full = torch.zeros(shape, device='cuda')
        vt._native.affine(vol.data_ptr(), shape, full.data_ptr(), shape, m, interp, vt._native.OOB_ZERO)
        parts = torch.full(shape, -1.0, device='cuda')
        for z0, z1 in ((0, 9), (9, 10), (10, 30), (30, 37)):
            vt._native.affine(vol.data_ptr(), shape, parts.data_ptr(), shape, m, interp, vt._native.OOB_ZERO,
                              z_range=(z0, z1))
        assert torch.equal(full, parts)


@pytest.mark.parametrize('mode', ['linear', 'filt_bspline', 'bspline_simple'])
@pytest.mark.parametrize('mname', ['rot45', 'rot_general', 'full_affine'])
def test_reshape(vt, mode, mname):
    """reshape=True: pad + conjugated matrix (transforms.py:171-178), host and device inputs."""
    import torch
    shape = (20, 24, 28)
    rng = np.random.default_rng(2)
    vol = rng.random(shape, dtype=np.float32)
    m = _matrices(vt, shape)[mname]
    got = vt.affine(vol, m, interpolation=mode, reshape=True, device='gpu:0')
    pb, pa, nd = vt.utils.compute_post_transform_dimensions(shape, m)
    assert got.shape == tuple(int(x) for x in nd)
    padded = np.pad(vol, list(zip(pb, pa)))
    m2 = (vt.utils.translation_matrix(-1 * pb) @ m @ vt.utils.translation_matrix(pb)).astype(np.float32)
    want = oracle.affine(padded, m2, mode)
    assert _err(got, want, _range(padded, mode)) <= TOL[mode]
    if not mode.startswith('filt'):  # (positive samples, non-negative weights: zero <=> skipped)
        assert np.array_equal(got == 0, want == 0)
    got_dev = vt.affine(torch.from_numpy(vol).cuda(), m, interpolation=mode, reshape=True, device='gpu:0')
    assert got_dev.shape == got.shape and _err(got_dev, want, _range(padded, mode)) <= TOL[mode]


@pytest.mark.parametrize('mode', ['linear', 'filt_bspline', 'bspline_simple'])
def test_convenience_wrappers(vt, mode):
    """translate / shear / scale / rotate, one-shot (transforms.py:51-106) and on a StaticVolume (volume.py:125-165):
    each is `affine` with the matching utils matrix -- about the array origin, like the reference."""
    import torch
    shape = (22, 26, 30)
    vol = np.random.default_rng(8).random(shape, dtype=np.float32)
    r = _range(vol, mode)
    U = vt.utils
    cases = [
        ('translate', ((1.5, -2.25, 0.75),), {}, U.translation_matrix((1.5, -2.25, 0.75))),
        ('shear', ((0.1, -0.05, 0.2),), {}, U.shear_matrix((0.1, -0.05, 0.2))),
        ('shear', (0.07,), {}, U.shear_matrix((0.07, 0.07, 0.07))),
        ('scale', ((1.2, 0.8, 1.1),), {}, U.scale_matrix((1.2, 0.8, 1.1))),
        ('scale', (1.3,), {}, U.scale_matrix((1.3, 1.3, 1.3))),
        ('rotate', ((4, 7, -5),), dict(rotation_order='sxyz'), U.rotation_matrix((4, 7, -5), rotation_order='sxyz')),
        ('rotate', ((0.05, 0.1, -0.08),), dict(rotation_units='rad'), U.rotation_matrix((0.05, 0.1, -0.08), rotation_units='rad')),
    ]
    sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:0')
    for name, args, kw, m in cases:
        want = oracle.affine(vol, m, mode)
        got = getattr(vt, name)(vol, *args, interpolation=mode, device='gpu:0', **kw)
        assert got.dtype == np.float32 and _err(got, want, r) <= TOL[mode], (name, args)
        got_sv = getattr(sv, name)(*args, **kw)
        assert _err(got_sv, want, r) <= TOL[mode], ('StaticVolume.' + name, args)
        out = torch.full(shape, 9.0, device='cuda')
        assert getattr(sv, name)(*args, output=out, **kw) is None
        want_o = oracle.affine(vol, m, mode, output=np.full(shape, 9.0, np.float32))
        assert _err(out.cpu().numpy(), want_o, r) <= TOL[mode], ('StaticVolume.' + name + ' output=', args)
        assert np.array_equal(out.cpu().numpy() == 9.0, want_o == 9.0)
    # the example's own calls (examples/transformation.py:16-24)
    got = vt.transform(vol, rotation=(25, 0, 0), rotation_order='rzxz', interpolation=mode, device='gpu:0')
    m = U.transform_matrix(rotation=(25, 0, 0), rotation_order='rzxz', center=_center(shape))
    assert _err(got, oracle.affine(vol, m, mode), r) <= TOL[mode]


@pytest.mark.parametrize('shape', [(20, 24, 28), (33, 47, 45), (70, 9, 130), (64, 64, 64)])
@pytest.mark.parametrize('interp', [0, 1, 2])
def test_slice4_family(vt, shape, interp):
    """Matrices that leave ONE axis alone run on the slice4 kernels (Z4 layout of that axis, vt_resample_z4.cu): same
    results as the gather kernels (float32 summation order aside) and as the oracle, for every march axis, both OOB
    policies, z-ranges and batches; bit-identical to the plain-layout slice kernels for axis 0."""
    import torch
    N = vt._native
    rng = np.random.default_rng(29)
    vol_np = rng.random(shape, dtype=np.float32)
    vol = torch.from_numpy(vol_np).cuda()
    mode = ['linear', 'bspline', 'bspline_simple'][interp]
    c = _center(shape)
    tm = vt.utils.transform_matrix
    orders = {0: ('rzxz', lambda a: (0, a, 0)), 1: ('ryzy', lambda a: (a, 0, 0)), 2: ('rzxz', lambda a: (a, 0, 0))}
    st = torch.cuda.current_stream().cuda_stream
    for axis in range(3):
        z4 = torch.empty(N.z4_bytes(shape, axis) // 4, device='cuda')
        N.pack_z4(vol.data_ptr(), shape, z4.data_ptr(), axis, device=0, stream=st)
        order, rot = orders[axis]
        shift = [0.5, -1.25, 2.5]
        shift[axis] = 3.0  # integer along the march axis
        mats = {f'rot{a}': tm(rotation=rot(a), rotation_order=order, center=c) for a in (0, 1, 45, 90, 133.7, -60)}
        mats['rot30_shift'] = tm(rotation=rot(30), rotation_order=order, center=c, translation=tuple(shift))
        shift[axis] = -4.0
        mats['rot-30_shift'] = tm(rotation=rot(-30), rotation_order=order, center=c, translation=tuple(shift))
        for name, m in mats.items():
            ax = N.z4_axis(shape, shape, m, interp)
            assert ax == axis or (name == 'rot0' and ax == 0), (axis, name, ax)
            for flag in (N.OOB_ZERO, N.OOB_SKIP):
                a = torch.full(shape, -7.0, device='cuda')
                b = torch.full(shape, -7.0, device='cuda')
                N.affine_z4(z4.data_ptr(), axis, shape, a.data_ptr(), shape, m, interp, flag, device=0, stream=st)
                N.affine(vol.data_ptr(), shape, b.data_ptr(), shape, m, interp, flag | N.KERNEL_GATHER)
                assert torch.equal(a == -7.0, b == -7.0), (axis, name, 'skipped sets differ')
                assert float((a - b).abs().max()) <= 1e-6, (axis, name, float((a - b).abs().max()))
            want = oracle.affine(vol_np, m, mode)
            got = np.where(a.cpu().numpy() == -7.0, 0, a.cpu().numpy())
            assert _err(got, want, 1.0) <= 1e-6, (axis, name)
        # a fractional shift along the march axis must NOT be accepted for that axis
        shift[axis] = 0.5
        bad = tm(rotation=rot(30), rotation_order=order, center=c, translation=tuple(shift))
        assert N.z4_axis(shape, shape, bad, interp) != axis
        with pytest.raises(RuntimeError):
            N.affine_z4(z4.data_ptr(), axis, shape, a.data_ptr(), shape, bad, interp, 0, device=0, stream=st)
        # batch + z-range
        ms = [tm(rotation=rot(a), rotation_order=order, center=c) for a in range(0, 180, 9)]
        out = torch.zeros((len(ms),) + shape, device='cuda')
        ref = torch.zeros((len(ms),) + shape, device='cuda')
        N.affine_z4(z4.data_ptr(), axis, shape, out.data_ptr(), shape, ms, interp, N.OOB_ZERO, z_range=(2, 7), device=0, stream=st)
        N.affine(vol.data_ptr(), shape, ref.data_ptr(), shape, ms, interp, N.OOB_ZERO | N.KERNEL_GATHER, z_range=(2, 7))
        assert float((out - ref).abs().max()) <= 1e-6, axis
        assert float(out[:, :2].abs().max()) == 0 and float(out[:, 7:].abs().max()) == 0
    # axis 0: the plain-layout slice kernels are the same arithmetic in the same order
    if shape[2] % 4 == 0:
        z4 = torch.empty(N.z4_bytes(shape, 0) // 4, device='cuda')
        N.pack_z4(vol.data_ptr(), shape, z4.data_ptr(), 0, device=0, stream=st)
        for a_deg in (0, 17, 45):
            m = tm(rotation=(0, a_deg, 0), rotation_order='rzxz', center=c)
            a = torch.zeros(shape, device='cuda')
            b = torch.zeros(shape, device='cuda')
            N.affine_z4(z4.data_ptr(), 0, shape, a.data_ptr(), shape, m, interp, N.OOB_ZERO, device=0, stream=st)
            N.affine(vol.data_ptr(), shape, b.data_ptr(), shape, m, interp, N.OOB_ZERO | N.KERNEL_SLICE)
            assert torch.equal(a, b), a_deg


@pytest.mark.parametrize('shape', [(40, 44, 48), (130, 50, 37), (250, 30, 250), (5, 7, 9)])
def test_prefilter_z4(vt, shape):
    """The prefilter writing the Z4 layout of axis 0 directly == prefilter into the plain layout, then pack."""
    import torch
    N = vt._native
    st = torch.cuda.current_stream().cuda_stream
    src = torch.rand(shape, device='cuda', generator=torch.Generator('cuda').manual_seed(3))
    plain = torch.empty(shape, device='cuda')
    N.prefilter(src.data_ptr(), shape, 0, st, dst_ptr=plain.data_ptr())
    want = torch.empty(N.z4_bytes(shape, 0) // 4, device='cuda')
    N.pack_z4(plain.data_ptr(), shape, want.data_ptr(), 0, device=0, stream=st)
    got = torch.full_like(want, float('nan'))
    ws = torch.empty(shape, device='cuda')
    N.prefilter_z4(src.data_ptr(), shape, got.data_ptr(), ws.data_ptr(), ws.numel() * 4, 0, st)
    assert not bool(torch.isnan(got).any())
    assert float((got - want).abs().max()) / float(plain.max() - plain.min()) <= 1e-6
    groups = (shape[0] + 3) // 4
    tail = got.view(groups, shape[1], shape[2], 4)[-1, :, :, shape[0] - 4 * (groups - 1):]
    assert float(tail.abs().max()) == 0.0 if tail.numel() else True


def test_errors(vt):
    vol = np.zeros((4, 4, 4), np.float32)
    with pytest.raises(ValueError):
        vt.affine(vol, np.identity(4), device='cpu')
    with pytest.raises(ValueError):
        vt.affine(vol, np.identity(4), interpolation='nearest', device='gpu')
    with pytest.raises(ValueError):
        vt.StaticVolume(np.zeros((4, 4), np.float32))
    with pytest.raises(ValueError):
        vt.utils.rotation_matrix((1, 2, 3), rotation_units='grad')
    with pytest.raises(ValueError):
        vt.utils.rotation_matrix((1, 2, 3), rotation_order='xyz')


@pytest.mark.parametrize('n', [512])
def test_full_size_properties(vt, n):
    """BASELINE sizes, where the oracle would take minutes: size-independent properties instead.
    identity (exact for linear, reconstruction for prefilter + cubic), linearity, z-slab composition, batching."""
    import torch
    N = vt._native
    shape = (n, n, n)
    g = torch.Generator('cuda').manual_seed(11)
    u = torch.rand(shape, device='cuda', generator=g)
    eye = np.identity(4, dtype=np.float32)
    out = torch.empty_like(u)
    vt.affine(u, eye, interpolation='linear', output=out, device='gpu')
    assert torch.equal(out, u), 'linear identity must be exact (alpha = 0 at texel centres)'
    vt.affine(u, eye, interpolation='filt_bspline_simple', output=out, device='gpu')
    inner = (slice(16, -16),) * 3
    assert float((out[inner] - u[inner]).abs().max()) <= 2e-5, 'prefilter followed by cubic sampling at the knots'
    vt.affine(u, eye, interpolation='filt_bspline', output=out, device='gpu')
    assert float((out[inner] - u[inner]).abs().max()) <= 2e-3 * 12, 'same through the 8-fetch path (8-bit weights)'
    # linearity of the whole pipeline (prefilter + resample) under a general matrix
    c = _center(shape)
    m = vt.utils.transform_matrix(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                                  translation=(5.5, -3.25, 2.0), center=c)
    v = torch.rand(shape, device='cuda', generator=g)
    tu, tv, tw = torch.zeros_like(u), torch.zeros_like(u), torch.zeros_like(u)
    for mode, tol in (('filt_bspline_simple', 2e-5), ('linear', 1e-5)):
        vt.affine(u, m, interpolation=mode, output=tu, device='gpu')
        vt.affine(v, m, interpolation=mode, output=tv, device='gpu')
        vt.affine(2.0 * u - 0.5 * v, m, interpolation=mode, output=tw, device='gpu')
        assert float((tw - (2.0 * tu - 0.5 * tv)).abs().max()) <= tol * 12, mode
    # z-slabs written by separate calls reproduce the single call bit for bit (both kernel families)
    rot = vt.utils.transform_matrix(rotation=(0, 33, 0), center=c)
    for mat in (m, rot):
        full = torch.zeros(shape, device='cuda')
        N.affine(u.data_ptr(), shape, full.data_ptr(), shape, mat, 2, N.OOB_ZERO)
        parts = torch.full(shape, -1.0, device='cuda')
        for z0, z1 in ((0, 100), (100, 101), (101, 400), (400, n)):
            N.affine(u.data_ptr(), shape, parts.data_ptr(), shape, mat, 2, N.OOB_ZERO, z_range=(z0, z1))
        assert torch.equal(full, parts)
    # rotating by +a then -a about the centre returns the interior of a smooth volume
    zz, yy, xx = torch.meshgrid(*(torch.linspace(-1, 1, n, device='cuda'),) * 3, indexing='ij')
    smooth = torch.exp(-4 * (zz ** 2 + yy ** 2 + xx ** 2)) * torch.cos(6 * xx) * torch.sin(5 * yy + 1)
    fwd = vt.utils.transform_matrix(rotation=(0, 20, 0), center=c)
    bwd = vt.utils.transform_matrix(rotation=(0, -20, 0), center=c)
    a, b = torch.zeros_like(smooth), torch.zeros_like(smooth)
    vt.affine(smooth, fwd, interpolation='filt_bspline_simple', output=a, device='gpu')
    vt.affine(a, bwd, interpolation='filt_bspline_simple', output=b, device='gpu')
    core = (slice(n // 4, -n // 4),) * 3
    assert float((b[core] - smooth[core]).abs().max()) <= 1e-4


def test_largest_config_1024(vt):
    """BASELINE configs[4] size (1024^3, 4 GiB per volume): 32-bit index limits, strides and grid sizes of every
    kernel family, through size-independent properties and family-vs-family agreement on a few output planes."""
    import torch
    N = vt._native
    n = 1024
    shape = (n, n, n)
    if torch.cuda.mem_get_info()[0] < 40 * 2 ** 30:
        pytest.skip('needs ~25 GiB of free device memory')
    u = torch.rand(shape, device='cuda', generator=torch.Generator('cuda').manual_seed(3))
    out = torch.empty_like(u)
    eye = np.identity(4, dtype=np.float32)
    vt.affine(u, eye, interpolation='linear', output=out, device='gpu')
    assert torch.equal(out, u)
    vt.affine(u, eye, interpolation='filt_bspline_simple', output=out, device='gpu')   # windowed prefilter, z-chunked
    probe = (slice(500, 524), slice(16, -16), slice(16, -16))
    assert float((out[probe] - u[probe]).abs().max()) <= 2e-5
    last = (slice(n - 40, n - 16), slice(16, -16), slice(16, -16))
    assert float((out[last] - u[last]).abs().max()) <= 2e-5
    c = _center(shape)
    general = vt.utils.transform_matrix(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60),
                                        translation=(5.5, -3.25, 2.0), center=c)
    rot = vt.utils.transform_matrix(rotation=(0, 45, 0), center=c)
    zr = (1000, 1008)
    ref = torch.zeros((8, n, n), device='cuda')
    got = torch.zeros((8, n, n), device='cuda')
    base = lambda t: t.data_ptr() - zr[0] * n * n * 4   # the ABI addresses plane z at dst + z*plane
    tex = N.Texture(u.data_ptr(), shape)
    for interp in (0, 1, 2):
        for mat, fams in ((rot, (N.KERNEL_SLICE,)), (general, (N.KERNEL_BRICK, 'tex'))):
            ref.zero_()
            N.affine(u.data_ptr(), shape, base(ref), shape, mat, interp, N.OOB_ZERO | N.KERNEL_GATHER, z_range=zr)
            for fam in fams:
                got.fill_(-1.0)
                if fam == 'tex':
                    if interp == 2:
                        continue
                    tex.affine(base(got), shape, mat, interp, N.OOB_ZERO, z_range=zr)
                else:
                    N.affine(u.data_ptr(), shape, base(got), shape, mat, interp, N.OOB_ZERO | fam, z_range=zr)
                assert float((got - ref).abs().max()) <= 1e-6, (interp, fam)
    tex.close()


def test_second_device_in_one_process(vt):
    """device='gpu:1' after 'gpu:0' in the same process: per-device kernel attributes (dynamic shared memory > 48 KB for
    the brick and slice kernels), contexts and the caller's current device being left alone."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs in one process')
    shape = (40, 48, 56)
    vol = np.random.default_rng(21).random(shape, dtype=np.float32)
    mats = _matrices(vt, shape)
    for mode in ('linear', 'filt_bspline', 'bspline_simple'):
        r = _range(vol, mode)
        for name in ('rot45', 'full_affine', 'downscale'):
            want = oracle.affine(vol, mats[name], mode)
            for dev in ('gpu:0', 'gpu:1'):
                got = vt.affine(vol, mats[name], interpolation=mode, device=dev)
                assert _err(got, want, r) <= TOL[mode], (mode, name, dev)
                assert torch.cuda.current_device() == 0
        sv = vt.StaticVolume(vol, interpolation=mode, device='gpu:1')
        out = sv.affine_many([mats['rot45'], mats['full_affine']])
        assert out.device.index == 1
        assert _err(out[1].cpu().numpy(), oracle.affine(vol, mats['full_affine'], mode), r) <= TOL[mode]


@pytest.mark.parametrize('pinned', [True, False])
def test_affine_many_into_host_memory(vt, pinned, monkeypatch):
    """affine_many(output=<numpy array>): K results stream into host memory through the three-chunk ring (K `.get()`s of
    volume.py:99 overlapped with the kernels): identical to the device results, ring slots reused, ragged last chunk."""
    import torch
    shape = (24, 28, 32)
    vol = np.random.default_rng(11).random(shape, dtype=np.float32)
    sv = vt.StaticVolume(vol, interpolation='filt_bspline', device='gpu:0')
    c = _center(shape)
    mats = [vt.utils.transform_matrix(rotation=(0, 9 * a, 0), rotation_order='rzxz', center=c) for a in range(11)]
    want = sv.affine_many(mats).cpu().numpy()
    monkeypatch.setattr(vt.StaticVolume, 'HOST_CHUNK_BYTES', 2 * 4 * int(np.prod(shape)))   # 2 volumes per chunk: 6 chunks
    out = vt.pinned_empty((len(mats),) + shape) if pinned else np.empty((len(mats),) + shape, np.float32)
    assert torch.from_numpy(out).is_pinned() == pinned
    out[:] = 7.0
    assert sv.affine_many(mats, output=out) is None
    assert np.array_equal(out, want)
    for a, m in enumerate(mats[:3]):
        assert np.array_equal(out[a], sv.affine(m))
    with pytest.raises(ValueError):
        sv.affine_many(mats, output=np.empty((3,) + shape, np.float32))
    with pytest.raises(ValueError):
        sv.affine_many(mats, output=np.empty((len(mats),) + shape, np.float64))
    # a registered (not allocator-owned) pinned array: the > 1 GiB route of pinned_empty, forced small
    from voltools_b200 import _native
    monkeypatch.setattr(_native, 'PINNED_CACHE_LIMIT', 0)
    big = vt.pinned_empty((len(mats),) + shape)
    assert torch.from_numpy(big).is_pinned()
    sv.affine_many(mats, output=big)
    assert np.array_equal(big, want)
    del big


@pytest.mark.parametrize('shape', [(5, 7, 9), (30, 20, 250), (3, 4, 8)])
def test_pad_rows(vt, shape):
    """vt_pad_rows_f32: dense rows -> rows padded to 16 bytes, pad columns zero (the copy an unfiltered volume of odd
    width needs before TMA can stage it; the reference's array copy, transforms.py:197-199)."""
    import torch
    from voltools_b200 import _native
    src = torch.rand(shape, device='cuda')
    row = _native.padded_row(shape[2]) + 4
    dst = torch.full((shape[0], shape[1], row), 5.0, device='cuda')
    _native.pad_rows(src.data_ptr(), shape, dst.data_ptr(), row, device=0, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :, :shape[2]], src) and bool((dst[:, :, shape[2]:] == 0).all())
    # one-shot calls on an odd width: a matrix that leaves axis 0 alone goes through the Z4 pack, a general one through
    # the padded copy -- both equal to the resident volume's results
    vol = src.cpu().numpy()
    sv = vt.StaticVolume(vol, interpolation='bspline_simple', device='gpu:0')
    for kw in (dict(rotation=(0, 30, 0), rotation_order='rzxz'), dict(rotation=(10, 20, 30), rotation_order='sxyz')):
        got = vt.transform(src, interpolation='bspline_simple', device='gpu:0', **kw)
        want = sv.transform(**kw)
        assert float(np.abs(got - want).max()) <= 1e-6
