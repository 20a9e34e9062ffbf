"""ctypes front-ends for oracle/vt_oracle.c (the CPU restatement) and oracle/_ref/*.so (the real reference).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFDIR = HERE / '_ref'
_SO = HERE / 'libvt_oracle.so'

# interpolation name -> which of the reference's three device functions it selects (transforms.py:11-17)
INTERP_FN = {'linear': 0, 'bspline': 1, 'bspline_simple': 2, 'filt_bspline': 1, 'filt_bspline_simple': 2}
TEX_RN, TEX_TRUNC, TEX_EXACT, TEX_HW = 0, 1, 2, 3

_lib = None
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """gcc-compile the C restatement (no fast-math, no compiler-chosen FMA contraction)."""
    src = HERE / 'vt_oracle.c'
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(['gcc', '-O2', '-fopenmp', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC',
                        '-o', str(_SO), str(src), '-lm'], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_SO))
        lib.vto_prefilter.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.vto_prefilter.restype = None
        lib.vto_prefilter_line.argtypes = [_f32p, ctypes.c_int, ctypes.c_int]
        lib.vto_prefilter_line.restype = None
        lib.vto_affine.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.vto_affine.restype = ctypes.c_long
        lib.vto_tex3d_many.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_long, _f32p,
                                       ctypes.c_int]
        lib.vto_tex3d_many.restype = None
        for n in ('vto_const_pole', 'vto_const_lambda', 'vto_const_anti'):
            getattr(lib, n).restype = ctypes.c_float
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_f32p)


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def constants():
    lib = _load()
    return lib.vto_const_pole(), lib.vto_const_lambda(), lib.vto_const_anti()


def prefilter(volume):
    """Returns the cubic B-spline coefficient volume (X, Y, Z passes of bspline.h) of a 3-D array."""
    v = _c32(volume).copy()
    _load().vto_prefilter(_p(v), *v.shape)
    return v


def prefilter_line(line):
    v = _c32(line).copy()
    _load().vto_prefilter_line(_p(v), v.size, 1)
    return v


def affine(volume, matrix, interpolation='linear', output=None, tex_rule=TEX_HW, out_shape=None, z_range=None):
    """Restatement of the reference GPU branch of affine(): returns the output volume.

    `volume` is the array the texture would hold; for filt_* names it is prefiltered here first
    (transforms.py:195-196).  `output`: initial contents of the destination (None -> zeros, as output=None).
    """
    v = _c32(volume)
    if interpolation.startswith('filt_bspline'):
        v = prefilter(v)
    m = _c32(matrix).reshape(4, 4)
    oshape = tuple(out_shape) if out_shape is not None else v.shape
    out = np.zeros(oshape, np.float32) if output is None else _c32(output).copy()
    assert out.shape == oshape
    z0, z1 = (0, oshape[0]) if z_range is None else z_range
    _load().vto_affine(_p(v), *v.shape, _p(out), *oshape, _p(m), INTERP_FN[interpolation], tex_rule, z0, z1)
    return out


def tex3d_many(volume, xyz, tex_rule=TEX_HW):
    v = _c32(volume)
    c = _c32(xyz).reshape(-1, 3)
    out = np.empty(len(c), np.float32)
    _load().vto_tex3d_many(_p(v), *v.shape, _p(c), len(c), _p(out), tex_rule)
    return out


# --------------------------------------------------------------------------------------------------
# the real reference (prebuilt by oracle/build_ref.py)
# --------------------------------------------------------------------------------------------------
_ref_gpu = None
_ref_host = None


def ref_host_available():
    return (REFDIR / 'libvt_ref_host.so').exists()


def _host():
    global _ref_host
    if _ref_host is None:
        lib = ctypes.CDLL(str(REFDIR / 'libvt_ref_host.so'))
        lib.ref_host_prefilter_line.argtypes = [_f32p, ctypes.c_uint, ctypes.c_int]
        lib.ref_host_prefilter_line.restype = None
        lib.ref_host_bspline.argtypes = [ctypes.c_float]
        lib.ref_host_bspline.restype = ctypes.c_float
        _ref_host = lib
    return _ref_host


def ref_host_prefilter_line(line):
    """The reference's own ConvertToInterpolationCoefficients (host build) on one contiguous line."""
    v = _c32(line).copy()
    _host().ref_host_prefilter_line(_p(v), v.size, 4)
    return v


def ref_host_bspline(t):
    return _host().ref_host_bspline(float(t))


def ref_gpu_available():
    if not (REFDIR / 'libvt_ref_gpu.so').exists():
        return False
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _gpu():
    global _ref_gpu
    if _ref_gpu is None:
        lib = ctypes.CDLL(str(REFDIR / 'libvt_ref_gpu.so'))
        lib.ref_init.argtypes = [ctypes.c_char_p]
        lib.ref_prefilter.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.ref_affine.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_int, ctypes.c_int,
                                   _f32p, ctypes.c_int, _f32p, _f32p]
        lib.ref_tex3d_sample.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, ctypes.c_long, _f32p]
        rc = lib.ref_init(str(REFDIR).encode())
        if rc:
            raise RuntimeError(f'ref_init failed ({rc})')
        _ref_gpu = lib
    return _ref_gpu


def transform_ref_gpu(volume, matrix, interpolation='linear', output=None, iters=0):
    """Runs the reference's own CUDA kernels (through a real texture object) on cuda:0.

    Returns (result, ms_kernel, ms_prefilter); the timings are None unless iters > 0.
    """
    v = _c32(volume)
    m = _c32(matrix).reshape(4, 4)
    out = np.zeros(v.shape, np.float32) if output is None else _c32(output).copy()
    msk, msp = ctypes.c_float(0), ctypes.c_float(0)
    rc = _gpu().ref_affine(_p(v), *v.shape, _p(m), INTERP_FN[interpolation],
                           int(interpolation.startswith('filt_bspline')), _p(out), iters, ctypes.byref(msk),
                           ctypes.byref(msp))
    if rc:
        raise RuntimeError(f'ref_affine failed ({rc})')
    return out, (msk.value if iters else None), (msp.value if iters else None)


def prefilter_ref_gpu(volume):
    v = _c32(volume).copy()
    rc = _gpu().ref_prefilter(_p(v), *v.shape)
    if rc:
        raise RuntimeError(f'ref_prefilter failed ({rc})')
    return v


def tex3d_ref_gpu(volume, xyz):
    v = _c32(volume)
    c = _c32(xyz).reshape(-1, 3)
    out = np.empty(len(c), np.float32)
    rc = _gpu().ref_tex3d_sample(_p(v), *v.shape, _p(c), len(c), _p(out))
    if rc:
        raise RuntimeError(f'ref_tex3d_sample failed ({rc})')
    return out


def ref_python_path():
    """Directory to put on sys.path to import the unmodified reference package (CPU path), or None."""
    p = REFDIR / 'py'
    if (p / 'voltools' / '__init__.py').exists():
        return str(p)
    if os.path.exists('/root/reference/voltools/__init__.py'):
        return '/root/reference'
    return None
