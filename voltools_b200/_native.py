"""ctypes binding of libvoltools_b200.so (include/voltools_b200.h).

There is no fallback: if the shared library is missing or a call fails this raises.  The library is built
in-tree by `python voltools_b200/csrc/build.py` (or `__graft_entry__.build()`).
"""
import ctypes
from pathlib import Path

import numpy as np

_SO = Path(__file__).resolve().parent / 'libvoltools_b200.so'

LINEAR, CUBIC_TEX, CUBIC_SIMPLE = 0, 1, 2
OOB_SKIP, OOB_ZERO = 0x0, 0x1
WEIGHTS_TEX_HW, WEIGHTS_EXACT = 0x0, 0x4
KERNEL_AUTO, KERNEL_GATHER, KERNEL_BRICK, KERNEL_SLICE = 0x00, 0x10, 0x20, 0x30
STAGE_CP_ASYNC = 0x100
MAX_BATCH = 32
ABI_VERSION = 6   # include/voltools_b200.h VT_ABI_VERSION

# interpolation name -> (device function, needs prefilter)      voltools/transforms.py:11-17
INTERPOLATIONS = {
    'linear': (LINEAR, False),
    'bspline': (CUBIC_TEX, False),
    'bspline_simple': (CUBIC_SIMPLE, False),
    'filt_bspline': (CUBIC_TEX, True),
    'filt_bspline_simple': (CUBIC_SIMPLE, True),
}

_lib = None
_f32p = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_i = ctypes.c_int


def lib():
    global _lib
    if _lib is None:
        if not _SO.exists():
            raise RuntimeError(f'{_SO} is missing: build it with `python voltools_b200/csrc/build.py` '
                               '(there is no CPU or PyTorch fallback)')
        L = ctypes.CDLL(str(_SO))
        L.vt_abi_version.restype = _i
        L.vt_error_string.restype = ctypes.c_char_p
        L.vt_error_string.argtypes = [_i]
        L.vt_device_count.argtypes = [ctypes.POINTER(_i)]
        L.vt_prefilter_f32.argtypes = [_vp, _vp, _i, _i, _i, _i, _i, _vp]
        L.vt_affine_f32.argtypes = [_vp, _i, _i, _i, _vp, _i, _i, _i, ctypes.c_longlong, _f32p, _i, _i,
                                    ctypes.c_uint, _i, _i, _i, _vp]
        L.vt_prefilter_strided_f32.argtypes = [_vp, _vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _vp]
        L.vt_prefilter_workspace_bytes.restype = ctypes.c_size_t
        L.vt_prefilter_workspace_bytes.argtypes = [_i, _i, _i, ctypes.c_longlong, ctypes.c_longlong]
        L.vt_prefilter_ws_f32.argtypes = [_vp, _vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, ctypes.c_size_t,
                                          _i, _i, _vp]
        L.vt_affine_strided_f32.argtypes = [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _i, _i, _i,
                                            ctypes.c_longlong, _f32p, _i, _i, ctypes.c_uint, _i, _i, _i, _vp]
        L.vt_affine_plan.argtypes = [_i, _i, _i, _i, _i, _i, _vp, _f32p, _i, _i, ctypes.c_uint, ctypes.POINTER(_i)]
        L.vt_prefilter_planes_f32.argtypes = [_vp, _vp, _vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i,
                                              _i, _i, _vp]
        L.vt_tex_create.argtypes = [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _i, _vp, ctypes.POINTER(_vp)]
        L.vt_tex_upload.argtypes = [_vp, _vp, ctypes.c_longlong, ctypes.c_longlong, _vp]
        L.vt_tex_destroy.argtypes = [_vp]
        L.vt_affine_tex_f32.argtypes = [_vp, _vp, _i, _i, _i, ctypes.c_longlong, _f32p, _i, _i, ctypes.c_uint, _i, _i, _vp]
        L.vt_project_workspace_bytes.restype = ctypes.c_size_t
        L.vt_project_workspace_bytes.argtypes = [_i, _i, _i]
        L.vt_project_strided_f32.argtypes = [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _i, _i, _i,
                                             ctypes.c_longlong, _f32p, _i, _i, ctypes.c_uint, _i, _i, _vp, ctypes.c_size_t,
                                             _i, _vp]
        L.vt_project_tex_f32.argtypes = [_vp, _vp, _i, _i, _i, ctypes.c_longlong, _f32p, _i, _i, ctypes.c_uint, _i, _i, _vp]
        L.vt_slice_plan.argtypes = [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i, _f32p, _i, _i,
                                    ctypes.c_uint, _i] + [ctypes.POINTER(_i)] * 7
        L.vt_host_ctx_create.argtypes = [_i, ctypes.POINTER(_vp)]
        L.vt_host_ctx_destroy.argtypes = [_vp]
        L.vt_host_ctx_trim.argtypes = [_vp]
        L.vt_host_affine_f32.argtypes = [_vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _f32p, _i, _i, ctypes.c_uint]
        L.vt_z4_bytes.restype = ctypes.c_size_t
        L.vt_z4_bytes.argtypes = [_i, _i, _i, _i]
        L.vt_pack_z4_f32.argtypes = [_vp, _i, _i, _i, ctypes.c_longlong, ctypes.c_longlong, _vp, _i, _i, _vp]
        L.vt_prefilter_z4_f32.argtypes = [_vp, _vp, _i, _i, _i, _vp, ctypes.c_size_t, _i, _vp]
        L.vt_z4_axis_of.argtypes = [_i, _i, _i, _i, _i, _i, _f32p, _i, _i, ctypes.POINTER(_i)]
        L.vt_affine_z4_f32.argtypes = [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, ctypes.c_longlong, _f32p, _i, _i,
                                       ctypes.c_uint, _i, _i, _i, _vp]
        L.vt_z4_plan.argtypes = [_i, _i, _i, _i, _i, _i, _i, _f32p, _i, _i, _i] + [ctypes.POINTER(_i)] * 6 + [_f32p]
        L.vt_launch_count.restype = ctypes.c_longlong
        L.vt_profile_enable.argtypes = [_i]
        L.vt_profile_kernel_name.restype = ctypes.c_char_p
        L.vt_profile_kernel_name.argtypes = [_i]
        L.vt_profile_read.argtypes = [_i, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
        L.vt_pad_rows_f32.argtypes = [_vp, _vp, _i, _i, _i, ctypes.c_longlong, _i, _vp]
        if L.vt_abi_version() != ABI_VERSION:
            raise RuntimeError(f'libvoltools_b200.so ABI version {L.vt_abi_version()} != {ABI_VERSION}: rebuild it')
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError(f'libvoltools_b200: {lib().vt_error_string(status).decode()} (status {status})')


def device_count():
    n = _i(0)
    lib().vt_device_count(ctypes.byref(n))
    return n.value


def launch_count():
    return lib().vt_launch_count()


def profile_enable(on=True):
    check(lib().vt_profile_enable(int(bool(on))))


def profile_read():
    """{kernel name: (total ms, launches)} for every kernel launched since profile_enable(True)."""
    out = {}
    for k in range(lib().vt_profile_kernel_count()):
        ms, n = ctypes.c_double(0), ctypes.c_longlong(0)
        check(lib().vt_profile_read(k, ctypes.byref(ms), ctypes.byref(n)))
        if n.value:
            out[lib().vt_profile_kernel_name(k).decode()] = (ms.value, n.value)
    return out


def _mats(matrices):
    m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
    return m, m.ctypes.data_as(_f32p)


def padded_row(width):
    """Row stride (elements) of a coefficient buffer: rows padded to 16 bytes so the kernels can stage with TMA."""
    return (int(width) + 3) // 4 * 4


def prefilter(src_ptr, shape, device=-1, stream=0, variant=0, dst_ptr=None, dst_strides=None, workspace=True):
    """Samples at src_ptr -> coefficients at dst_ptr (default: in place).  dst_strides = (row, plane) in elements.

    workspace=True (out-of-place variant 0 only): a scratch volume from torch's caching allocator is handed to the
    library so that its Z sweep can run out of place in z-chunks (vt_prefilter_ws_f32); the library itself
    allocates nothing.  `stream` must then be torch's current stream on that device (it is, for every caller here).
    """
    dst = src_ptr if dst_ptr is None else dst_ptr
    if dst_strides is None:
        dst_strides = (int(shape[2]), int(shape[1]) * int(shape[2]))
    row, plane = int(dst_strides[0]), int(dst_strides[1])
    ws, ws_ptr, ws_bytes = None, None, 0
    if workspace is not True and workspace is not False and workspace is not None:  # a caller-owned scratch tensor
        ws, ws_ptr, ws_bytes = workspace, workspace.data_ptr(), workspace.numel() * workspace.element_size()
        if ws_bytes < lib().vt_prefilter_workspace_bytes(shape[0], shape[1], shape[2], row, plane):
            ws_ptr, ws_bytes = None, 0
    elif workspace and variant == 0 and dst != src_ptr and int(shape[0]) >= 128:
        import torch
        ws_bytes = lib().vt_prefilter_workspace_bytes(shape[0], shape[1], shape[2], row, plane)
        dev = torch.cuda.current_device() if device < 0 else device
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=f'cuda:{dev}')
        ws_ptr = ws.data_ptr()
    check(lib().vt_prefilter_ws_f32(src_ptr, dst, shape[0], shape[1], shape[2], row, plane, ws_ptr, ws_bytes, variant,
                                    device, stream))
    del ws


def pad_rows(src_ptr, shape, dst_ptr, row, device=-1, stream=0):
    """Dense volume -> rows padded to `row` elements (vt_pad_rows_f32)."""
    check(lib().vt_pad_rows_f32(src_ptr, dst_ptr, int(shape[0]), int(shape[1]), int(shape[2]), int(row), device, stream))


def prefilter_planes(src_ptr, ws_ptr, dst_ptr, shape, dst_strides, xy_range, z_range, device=-1, stream=0):
    """Streaming prefilter step (vt_prefilter_planes_f32): XY passes of sample planes xy_range into the workspace,
    Z pass of coefficient planes z_range from the workspace into dst."""
    check(lib().vt_prefilter_planes_f32(src_ptr, ws_ptr, dst_ptr, int(shape[0]), int(shape[1]), int(shape[2]),
                                        int(dst_strides[0]), int(dst_strides[1]), int(xy_range[0]), int(xy_range[1]),
                                        int(z_range[0]), int(z_range[1]), device, stream))


def affine(src_ptr, src_shape, dst_ptr, dst_shape, matrices, interp, flags=0, batch_stride=None, z_range=None,
           device=-1, stream=0, src_strides=None):
    """src_strides = (row, plane) element strides of the source (default: dense)."""
    m, mp = _mats(matrices)
    if batch_stride is None:
        batch_stride = int(dst_shape[0]) * int(dst_shape[1]) * int(dst_shape[2])
    z0, z1 = (0, dst_shape[0]) if z_range is None else z_range
    if src_strides is None:
        src_strides = (int(src_shape[2]), int(src_shape[1]) * int(src_shape[2]))
    check(lib().vt_affine_strided_f32(src_ptr, *map(int, src_shape), int(src_strides[0]), int(src_strides[1]), dst_ptr,
                                      *map(int, dst_shape), batch_stride, mp, len(m), interp, flags, z0, z1, device,
                                      stream))


# ---- slice4 family: Z4 layout (vt_resample_z4.cu) -----------------------------------------------------------
import os as _os


def z4_wanted(interp, resident, axis=0, filtered=False, width=None):
    """Policy: does a launch whose matrices leave `axis` alone go to the slice4 kernels?  A resident volume always
    (the Z4 copy is packed once and kept).  A one-shot call pays a pack pass (8 B/voxel) unless the prefilter writes
    the layout directly (filt_*, axis 0): measured at 512^3 the pass is paid back whenever
    the alternative is the general-matrix kernels (axes 1 and 2: 2-6x), and for axis 0 only with the prefilter
    (pack + slice4 221 / 236 Gvox/s against 250 / 280 for the plain-layout slice kernels, bspline / bspline_simple).
    A one-shot volume whose rows are not a multiple of 16 bytes (width 250: the reference's benchmark size) needs a
    copy pass before ANY TMA-staged kernel can read it; the pack pass is that copy, so it goes to slice4 as well.
    VT_Z4=0/1 forces it off/on (A/B measurements)."""
    force = _os.environ.get('VT_Z4')
    if force is not None:
        return force != '0'
    if resident or axis != 0:
        return True
    if width is not None and padded_row(width) != int(width):
        return True
    return bool(filtered)


def z4_bytes(shape, axis):
    return lib().vt_z4_bytes(*map(int, shape), int(axis))


def pack_z4(src_ptr, shape, dst4_ptr, axis, device=-1, stream=0, src_strides=None):
    """Plain volume (src_strides = (row, plane) in elements, default dense) -> Z4 layout of `axis`."""
    if src_strides is None:
        src_strides = (int(shape[2]), int(shape[1]) * int(shape[2]))
    check(lib().vt_pack_z4_f32(src_ptr, *map(int, shape), int(src_strides[0]), int(src_strides[1]), dst4_ptr, int(axis),
                               device, stream))


def prefilter_z4(src_ptr, shape, dst4_ptr, ws_ptr, ws_bytes, device=-1, stream=0):
    """Samples -> coefficients in the Z4 layout of axis 0 (windowed prefilter; workspace of prod(shape) floats)."""
    check(lib().vt_prefilter_z4_f32(src_ptr, dst4_ptr, *map(int, shape), ws_ptr, ws_bytes, device, stream))


def z4_axis(src_shape, dst_shape, matrices, interp):
    """The axis (0..2) every matrix leaves alone in the way the slice4 kernels need, or -1."""
    m, mp = _mats(matrices)
    ax = _i(-1)
    check(lib().vt_z4_axis_of(*map(int, src_shape), *map(int, dst_shape), mp, len(m), interp, ctypes.byref(ax)))
    return ax.value


def affine_z4(src4_ptr, axis, src_shape, dst_ptr, dst_shape, matrices, interp, flags=0, batch_stride=None, z_range=None,
              device=-1, stream=0):
    m, mp = _mats(matrices)
    if batch_stride is None:
        batch_stride = int(dst_shape[0]) * int(dst_shape[1]) * int(dst_shape[2])
    z0, z1 = (0, dst_shape[0]) if z_range is None else z_range
    check(lib().vt_affine_z4_f32(src4_ptr, int(axis), *map(int, src_shape), dst_ptr, *map(int, dst_shape), batch_stride,
                                 mp, len(m), interp, flags, z0, z1, device, stream))


def z4_plan(shape, matrices, interp, axis, sms=148):
    """Host-only: the slice4 family's launch decisions (vt_z4_plan) as a dict, or None if unsupported."""
    m, mp = _mats(matrices)
    out = [_i(0) for _ in range(4)]
    shapes, pitches, wf = (_i * len(m))(), (_i * len(m))(), (ctypes.c_float * len(m))()
    rc = lib().vt_z4_plan(int(axis), *map(int, shape), *map(int, shape), mp, len(m), interp, sms,
                          *[ctypes.byref(o) for o in out], shapes, pitches, wf)
    if rc == 2:
        return None
    check(rc)
    return {'chunks': out[0].value, 'm_chunk': out[1].value, 'box_w': out[2].value, 'box_h': out[3].value,
            'shapes': list(shapes), 'pitches': list(pitches), 'wavefronts': [float(v) for v in wf]}


def download(t, stream=0):
    """Device tensor -> numpy through a pinned staging buffer (a pageable destination makes cudaMemcpy stage through
    the driver's small bounce buffer at a fraction of the link rate): the `.get()` of transforms.py:223."""
    import torch
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return host.numpy()


def project_workspace_bytes(src_shape):
    return lib().vt_project_workspace_bytes(*map(int, src_shape))


def project(src_ptr, src_shape, proj_ptr, dst_shape, matrices, interp, flags=0, batch_stride=None, z_range=None,
            device=-1, stream=0, src_strides=None, workspace_ptr=None, workspace_bytes=0):
    """Sum over axis 0 of the transformed volume (vt_project_strided_f32): K images of dst_shape[1:] at proj_ptr."""
    m, mp = _mats(matrices)
    if batch_stride is None:
        batch_stride = int(dst_shape[1]) * int(dst_shape[2])
    z0, z1 = (0, dst_shape[0]) if z_range is None else z_range
    if src_strides is None:
        src_strides = (int(src_shape[2]), int(src_shape[1]) * int(src_shape[2]))
    check(lib().vt_project_strided_f32(src_ptr, *map(int, src_shape), int(src_strides[0]), int(src_strides[1]), proj_ptr,
                                       *map(int, dst_shape), batch_stride, mp, len(m), interp, flags, z0, z1,
                                       workspace_ptr, workspace_bytes, device, stream))


def slice_plan(shape, matrices, interp, flags=0, sms=148, src_strides=None, src_ptr=None):
    """Host-only: the slice family's launch decisions (vt_slice_plan) as a dict, or None if the matrices are not of the
    slice family."""
    m, mp = _mats(matrices)
    if src_strides is None:
        src_strides = (padded_row(shape[2]), int(shape[1]) * padded_row(shape[2]))
    out = [_i(0) for _ in range(5)]
    shapes, pitches = (_i * len(m))(), (_i * len(m))()
    rc = lib().vt_slice_plan(src_ptr, *map(int, shape), int(src_strides[0]), int(src_strides[1]), *map(int, shape), mp,
                             len(m), interp, flags, sms, *[ctypes.byref(o) for o in out], shapes, pitches)
    if rc == 2:
        return None
    check(rc)
    return {'chunks': out[0].value, 'z_chunk': out[1].value, 'tma': bool(out[2].value), 'box_w': out[3].value,
            'box_h': out[4].value, 'shapes': list(shapes), 'pitches': list(pitches)}


def affine_plan(src_ptr, src_shape, dst_shape, matrices, interp, flags=0):
    m, mp = _mats(matrices)
    fam = _i(0)
    check(lib().vt_affine_plan(*map(int, src_shape), *map(int, dst_shape), src_ptr, mp, len(m), interp, flags,
                               ctypes.byref(fam)))
    return {1: 'gather', 2: 'brick', 3: 'slice'}[fam.value]


def launch_plan(src_ptr, src_shape, dst_shape, matrices, interp, resident=True, filtered=False):
    """Which kernel family the PYTHON layer runs for these matrices (it routes matrices that leave one axis alone to
    the slice4 kernels before asking the plain-layout planner): {'family': 'slice4', 'axis': m} or
    {'family': 'slice' | 'brick' | 'gather'}."""
    axis = z4_axis(src_shape, dst_shape, matrices, interp)
    if axis >= 0 and z4_wanted(interp, resident, axis, filtered, width=src_shape[2]):
        return {'family': 'slice4', 'axis': axis}
    return {'family': affine_plan(src_ptr, src_shape, dst_shape, matrices, interp)}


class Texture:
    """Owner of a vt_tex: the sampled volume as a 3-D CUDA array + texture object (texture kernel family)."""

    def __init__(self, src_ptr, shape, src_strides=None, device=-1, stream=0):
        self.shape = tuple(int(v) for v in shape)
        if src_strides is None:
            src_strides = (self.shape[2], self.shape[1] * self.shape[2])
        self._h = _vp()
        check(lib().vt_tex_create(src_ptr, *self.shape, int(src_strides[0]), int(src_strides[1]), device, stream,
                                  ctypes.byref(self._h)))

    def upload(self, src_ptr, src_strides=None, stream=0):
        if src_strides is None:
            src_strides = (self.shape[2], self.shape[1] * self.shape[2])
        check(lib().vt_tex_upload(self._h, src_ptr, int(src_strides[0]), int(src_strides[1]), stream))

    def affine(self, dst_ptr, dst_shape, matrices, interp, flags=0, batch_stride=None, z_range=None, stream=0):
        m, mp = _mats(matrices)
        if batch_stride is None:
            batch_stride = int(dst_shape[0]) * int(dst_shape[1]) * int(dst_shape[2])
        z0, z1 = (0, dst_shape[0]) if z_range is None else z_range
        check(lib().vt_affine_tex_f32(self._h, dst_ptr, *map(int, dst_shape), batch_stride, mp, len(m), interp, flags,
                                      z0, z1, stream))

    def project(self, proj_ptr, dst_shape, matrices, interp, flags=0, batch_stride=None, z_range=None, stream=0):
        m, mp = _mats(matrices)
        if batch_stride is None:
            batch_stride = int(dst_shape[1]) * int(dst_shape[2])
        z0, z1 = (0, dst_shape[0]) if z_range is None else z_range
        check(lib().vt_project_tex_f32(self._h, proj_ptr, *map(int, dst_shape), batch_stride, mp, len(m), interp, flags,
                                       z0, z1, stream))

    def close(self):
        if self._h:
            lib().vt_tex_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostContext:
    """Owner of a vt_host_ctx (pinned staging + device buffers for the numpy-in / numpy-out path)."""

    def __init__(self, device=-1):
        self._h = _vp()
        check(lib().vt_host_ctx_create(device, ctypes.byref(self._h)))

    def affine(self, src, dst, matrix, interp, prefilter, flags=0):
        m, mp = _mats(matrix)
        assert src.dtype == np.float32 and dst.dtype == np.float32 and src.flags.c_contiguous and dst.flags.c_contiguous
        check(lib().vt_host_affine_f32(self._h, src.ctypes.data, *src.shape, dst.ctypes.data, *dst.shape, mp, interp,
                                       int(prefilter), flags))

    def trim(self):
        """Release the context's device buffers (they grow to the largest volume seen); it stays usable."""
        if self._h:
            check(lib().vt_host_ctx_trim(self._h))

    def close(self):
        if self._h:
            lib().vt_host_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PINNED_CACHE_LIMIT = 1 << 30


def pinned_empty(shape):
    """A float32 numpy array in page-locked memory: device copies into / out of it are asynchronous and run at link
    speed.  Up to 1 GiB it comes from torch's caching host allocator (the block returns to the cache when the array is
    garbage-collected); larger arrays -- the host side of `StaticVolume.affine_many(output=...)` for a whole sweep -- are
    ordinary numpy memory registered with the driver at its exact size (the caching allocator rounds up to a power of
    two) and unregistered when the array dies."""
    import torch
    shape = tuple(int(v) for v in shape)
    n = 4
    for v in shape:
        n *= v
    if n <= PINNED_CACHE_LIMIT:
        return torch.empty(shape, dtype=torch.float32, pin_memory=True).numpy()
    import weakref
    torch.cuda.init()
    rt = torch.cuda.cudart()
    a = np.empty(shape, dtype=np.float32)
    err = rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0)
    if int(err) != 0:
        raise RuntimeError(f'cudaHostRegister of {a.nbytes} bytes failed: error {int(err)}')
    weakref.finalize(a, rt.cudaHostUnregister, a.ctypes.data)
    return a


def as_pinned(a):
    """`a` if it is already page-locked, else a pinned copy (ATen's multi-threaded host copy)."""
    import torch
    t = torch.from_numpy(a)
    if t.is_pinned():
        return a
    p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    p.copy_(t)
    return p.numpy()
