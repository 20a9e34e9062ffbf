"""One-shot transforms: the reference's public functions (voltools/transforms.py:25-229) on libvoltools_b200.

Same names, arguments and error behaviour as the reference, with these deliberate differences:
  * the only devices are 'gpu' / 'gpu:X' -- there is no 'cpu' device and no fallback of any kind;
  * a device-resident input is never modified (the reference prefilters and zero-fills a CuPy input in place);
  * the matrix is cast to float32 (the reference reinterprets a float64 matrix);
  * the caller's current CUDA device is left unchanged.
Inputs may be numpy arrays, torch CUDA tensors or any object exposing __cuda_array_interface__ (CuPy);
`output=` may be a torch CUDA tensor or a __cuda_array_interface__ object and is written in place:
out-of-bounds voxels keep their previous contents and the function returns None (transforms.py:224-226).
With output=None the result is returned as a numpy array (transforms.py:221-223).
Extension: with a numpy `volume`, `output=` may also be a float32 C-contiguous numpy array (ideally pinned); it is
then completely overwritten (out-of-bounds voxels = 0, like output=None) straight from the device and the function
returns None -- this saves the host-side copy of the result.
"""
import threading
from typing import Tuple, Union

import numpy as np

from . import _native
from . import utils
from .utils import scale_matrix, shear_matrix, rotation_matrix, translation_matrix, transform_matrix

_INTERPOLATIONS = _native.INTERPOLATIONS
AVAILABLE_INTERPOLATIONS = list(_INTERPOLATIONS.keys())
AVAILABLE_DEVICES = utils.get_available_devices()

# numpy-in / numpy-out contexts (pinned staging + device buffers), one per host thread and device: transform() may be
# called from several threads at once, and two calls in flight overlap one's upload with the other's download
_host_tls = threading.local()


_host_registry = []  # every HostContext ever created (weak): release_host_buffers() reaches other threads' contexts too


def _host_context(dev):
    ctxs = getattr(_host_tls, 'ctx', None)
    if ctxs is None:
        ctxs = _host_tls.ctx = {}
    ctx = ctxs.get(dev)
    if ctx is None:
        import weakref
        ctx = ctxs[dev] = _native.HostContext(dev)
        _host_registry.append(weakref.ref(ctx))
    return ctx


def release_host_buffers():
    """Free the device buffers held by the numpy-in / numpy-out contexts of every thread (one context per host thread
    and device; each keeps source, coefficient, workspace and output buffers sized for the largest volume it has seen --
    16 GiB after one 1024^3 filt_* call).  The contexts stay usable and re-allocate on demand.  Not to be called while
    another thread is inside transform() / affine()."""
    alive = []
    for ref in _host_registry:
        ctx = ref()
        if ctx is not None:
            ctx.trim()
            alive.append(ref)
    _host_registry[:] = alive


# ----------------------------------------------------------------------------------------------------
# array adapters (plumbing only: pointers go to the C ABI)
# ----------------------------------------------------------------------------------------------------
class _DeviceView:
    """A float32 C-contiguous device array seen as (pointer, shape, device ordinal) + the owner keeping it alive."""

    def __init__(self, ptr, shape, device, owner):
        self.ptr, self.shape, self.device, self.owner = ptr, tuple(int(s) for s in shape), device, owner


def _torch():
    import torch
    return torch


def _is_host(a):
    return isinstance(a, np.ndarray)


def _device_view(a, what='volume') -> _DeviceView:
    torch = _torch()
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise ValueError(f'{what}: torch tensor must live on a CUDA device')
        if a.dtype != torch.float32 or not a.is_contiguous():
            raise ValueError(f'{what}: expected a contiguous float32 tensor')
        return _DeviceView(a.data_ptr(), a.shape, a.device.index, a)
    cai = getattr(a, '__cuda_array_interface__', None)
    if cai is None:
        raise ValueError(f'{what}: expected numpy array, torch CUDA tensor or __cuda_array_interface__ object')
    if cai['typestr'] not in ('<f4', '=f4', 'f4'):
        raise ValueError(f'{what}: expected float32 data')
    if cai.get('strides') is not None:
        expect = tuple(int(np.prod(cai['shape'][i + 1:])) * 4 for i in range(len(cai['shape'])))
        if tuple(cai['strides']) != expect:
            raise ValueError(f'{what}: expected a C-contiguous array')
    dev = getattr(getattr(a, 'device', None), 'id', None)
    if dev is None:
        dev = torch.cuda.current_device()
    return _DeviceView(int(cai['data'][0]), cai['shape'], int(dev), a)


def _resolve_device(device: str, *views) -> int:
    """'gpu' -> device of the first device-resident argument, else the current device; 'gpu:X' -> X."""
    idx = utils.device_index(device)
    if idx >= 0:
        return idx
    for v in views:
        if v is not None:
            return v.device
    return _torch().cuda.current_device()


def _stream(dev: int) -> int:
    return _torch().cuda.current_stream(dev).cuda_stream


def _check_args(interpolation, device):
    if device not in AVAILABLE_DEVICES:
        raise ValueError(f'Unknown device ({device}), must be one of {AVAILABLE_DEVICES}')
    if interpolation not in AVAILABLE_INTERPOLATIONS:
        raise ValueError(f'Interpolation must be one of {AVAILABLE_INTERPOLATIONS}')


# ----------------------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------------------
def transform(volume,
              scale: Union[float, Tuple[float, float, float], np.ndarray] = None,
              shear: Union[float, Tuple[float, float, float], np.ndarray] = None,
              rotation: Union[Tuple[float, float, float], np.ndarray] = None,
              rotation_units: str = 'deg', rotation_order: str = 'rzxz',
              translation: Union[Tuple[float, float, float], np.ndarray] = None,
              center: Union[Tuple[float, float, float], np.ndarray] = None,
              interpolation: str = 'linear',
              reshape: bool = False,
              profile: bool = False,
              output=None,
              device: str = 'gpu'):
    """transforms.py:25-48."""
    if center is None:
        center = np.divide(np.subtract(tuple(volume.shape), 1), 2, dtype=np.float32)
    if isinstance(scale, float):
        scale = (scale, scale, scale)
    if isinstance(shear, float):
        shear = (shear, shear, shear)
    m = transform_matrix(scale, shear, rotation, rotation_units, rotation_order, translation, center)
    return affine(volume, m, interpolation, reshape, profile, output, device)


def translate(volume, translation: Tuple[float, float, float], interpolation: str = 'linear', reshape: bool = False,
              profile: bool = False, output=None, device: str = 'gpu'):
    """transforms.py:51-60."""
    return affine(volume, translation_matrix(translation), interpolation, reshape, profile, output, device)


def shear(volume, coefficients: Union[float, Tuple[float, float, float]], interpolation: str = 'linear',
          reshape: bool = False, profile: bool = False, output=None, device: str = 'gpu'):
    """transforms.py:63-76."""
    if isinstance(coefficients, float):
        coefficients = (coefficients, coefficients, coefficients)
    return affine(volume, shear_matrix(coefficients), interpolation, reshape, profile, output, device)


def scale(volume, coefficients: Union[float, Tuple[float, float, float]], interpolation: str = 'linear',
          reshape: bool = False, profile: bool = False, output=None, device: str = 'gpu'):
    """transforms.py:79-92."""
    if isinstance(coefficients, float):
        coefficients = (coefficients, coefficients, coefficients)
    return affine(volume, scale_matrix(coefficients), interpolation, reshape, profile, output, device)


def rotate(volume, rotation: Tuple[float, float, float], rotation_units: str = 'deg', rotation_order: str = 'rzxz',
           interpolation: str = 'linear', reshape: bool = False, profile: bool = False, output=None,
           device: str = 'gpu'):
    """transforms.py:95-106 (about the array origin, like the reference)."""
    m = rotation_matrix(rotation=rotation, rotation_units=rotation_units, rotation_order=rotation_order)
    return affine(volume, m, interpolation, reshape, profile, output, device)


def project(volume,
            scale: Union[float, Tuple[float, float, float], np.ndarray] = None,
            shear: Union[float, Tuple[float, float, float], np.ndarray] = None,
            rotation: Union[Tuple[float, float, float], np.ndarray] = None,
            rotation_units: str = 'deg', rotation_order: str = 'rzxz',
            translation: Union[Tuple[float, float, float], np.ndarray] = None,
            center: Union[Tuple[float, float, float], np.ndarray] = None,
            interpolation: str = 'linear',
            device: str = 'gpu'):
    """`transform(volume, ...).sum(axis=0)` fused (examples/projections.py:44-47): the transformed volume is never
    written.  Returns the (d1, d2) projection as a numpy array.  For many projections of one volume use
    StaticVolume.project / project_many (the volume is uploaded and prefiltered once)."""
    from .volume import StaticVolume
    _check_args(interpolation, device)
    return StaticVolume(volume, interpolation=interpolation, device=device).project(
        scale, shear, rotation, rotation_units, rotation_order, translation, center)


def affine(volume, transform_m: np.ndarray, interpolation: str = 'linear', reshape: bool = False,
           profile: bool = False, output=None, device: str = 'gpu'):
    """GPU branch of transforms.py:109-229."""
    _check_args(interpolation, device)
    if len(volume.shape) != 3:
        raise ValueError('Expected a 3D array')
    torch = _torch()
    interp, needs_prefilter = _INTERPOLATIONS[interpolation]
    m = np.ascontiguousarray(transform_m, dtype=np.float32).reshape(4, 4)

    host_in = _is_host(volume)
    host_out = output is not None and _is_host(output)
    if host_out:
        if not host_in:
            raise ValueError('a numpy `output` needs a numpy `volume`')
        if output.dtype != np.float32 or not output.flags.c_contiguous:
            raise ValueError('output: expected a C-contiguous float32 numpy array')
    vin = None if host_in else _device_view(volume, 'volume')
    vout = None if (output is None or host_out) else _device_view(output, 'output')
    dev = _resolve_device(device, vin, vout)

    if reshape:
        # transforms.py:171-178: zero-pad the input so the whole transformed volume fits, and conjugate the
        # matrix with the pad offset.  Done exactly as the reference does (pad first, then prefilter/sample the
        # padded volume) so coordinates and results are identical.
        pad_before, pad_after, _ = utils.compute_post_transform_dimensions(tuple(volume.shape), m)
        m = (translation_matrix(-1 * pad_before) @ m @ translation_matrix(pad_before)).astype(np.float32)
        pads = list(zip((int(p) for p in pad_before), (int(p) for p in pad_after)))
        if host_in:
            volume = np.pad(np.asarray(volume, dtype=np.float32), pads, mode='constant')
        else:
            t = torch.as_tensor(vin.owner, device=f'cuda:{vin.device}') if not isinstance(vin.owner, torch.Tensor) \
                else vin.owner
            flat = [p for pair in reversed(pads) for p in pair]
            volume = torch.nn.functional.pad(t, flat).contiguous()
            vin = _device_view(volume)
    shape = tuple(int(s) for s in volume.shape)
    if vout is not None and vout.shape != shape:
        raise ValueError(f'output shape {vout.shape} does not match the volume shape {shape}')
    if host_out and tuple(output.shape) != shape:
        raise ValueError(f'output shape {tuple(output.shape)} does not match the volume shape {shape}')

    with torch.cuda.device(dev):
        if profile:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()

        if host_in and vout is None:
            # numpy in -> numpy out (transforms.py:180-223): pipelined host path inside the library
            # The library copies straight from / to these arrays, and only page-locked memory copies asynchronously at
            # link speed: a pageable input is staged through a pinned buffer (multi-threaded host copy), the result
            # lives in pinned memory from torch's caching host allocator (returned as an ordinary numpy array).
            src = _native.as_pinned(np.ascontiguousarray(volume, dtype=np.float32))
            result = output if host_out else _native.pinned_empty(shape)
            ctx = _host_context(dev)
            ctx.affine(src, result, m, interp, needs_prefilter)
            if host_out:
                result = None
        else:
            stream = _stream(dev)
            if host_in:
                src_t = torch.from_numpy(np.ascontiguousarray(volume, dtype=np.float32)).to(f'cuda:{dev}')
            else:
                src_t = vin.owner if isinstance(vin.owner, torch.Tensor) \
                    else torch.as_tensor(vin.owner, device=f'cuda:{dev}')
            if src_t.device.index != dev:
                # (a raw pointer of another GPU handed to a kernel on `dev` would fault, not raise)
                src_t = src_t.to(f'cuda:{dev}')
            if vout is not None and vout.device != dev:
                raise ValueError(f'output lives on gpu:{vout.device}, the transform runs on gpu:{dev}')
            src_ptr = src_t.data_ptr()
            if vout is None:
                out_t = torch.empty(shape, dtype=torch.float32, device=f'cuda:{dev}')
                dst_ptr, flags = out_t.data_ptr(), _native.OOB_ZERO
            else:
                out_t, dst_ptr, flags = None, vout.ptr, _native.OOB_SKIP
            # A matrix that leaves one axis alone (rotations about an axis through the centre: the slice4 kernels,
            # vt_resample_z4.cu) samples the Z4 layout of that axis: the prefilter's Z sweep writes it directly for
            # axis 0, otherwise one pack pass (the reference pays the same pass for its CUDA-array copy,
            # transforms.py:197-199).
            axis = _native.z4_axis(shape, shape, m, interp)
            if axis >= 0 and _native.z4_wanted(interp, False, axis, needs_prefilter, width=shape[2]):
                z4_t = torch.empty(_native.z4_bytes(shape, axis) // 4, dtype=torch.float32, device=f'cuda:{dev}')
                if needs_prefilter:
                    ws_t = torch.empty(shape, dtype=torch.float32, device=f'cuda:{dev}')
                    if axis == 0:
                        _native.prefilter_z4(src_ptr, shape, z4_t.data_ptr(), ws_t.data_ptr(), ws_t.numel() * 4, dev, stream)
                    else:
                        coef_t = torch.empty(shape, dtype=torch.float32, device=f'cuda:{dev}')
                        _native.prefilter(src_ptr, shape, dev, stream, dst_ptr=coef_t.data_ptr(), workspace=ws_t)
                        _native.pack_z4(coef_t.data_ptr(), shape, z4_t.data_ptr(), axis, device=dev, stream=stream)
                        del coef_t
                    del ws_t
                else:
                    _native.pack_z4(src_ptr, shape, z4_t.data_ptr(), axis, device=dev, stream=stream)
                _native.affine_z4(z4_t.data_ptr(), axis, shape, dst_ptr, shape, m, interp, flags, device=dev, stream=stream)
                del z4_t
            else:
                # the sampled volume lives in a private buffer whose rows are padded to 16 bytes: the kernels can then
                # stage it with TMA whatever the width is.  The prefilter writes that layout directly (out of place:
                # the caller's array is never modified); an unfiltered volume with an odd width is copied once.
                row = _native.padded_row(shape[2])
                src_strides = (row, shape[1] * row)
                if needs_prefilter:
                    coef_t = torch.empty((shape[0], shape[1], row), dtype=torch.float32, device=f'cuda:{dev}')
                    _native.prefilter(src_ptr, shape, dev, stream, dst_ptr=coef_t.data_ptr(), dst_strides=src_strides)
                    src_t, src_ptr = coef_t, coef_t.data_ptr()
                elif row != shape[2]:
                    pad_t = torch.empty((shape[0], shape[1], row), dtype=torch.float32, device=f'cuda:{dev}')
                    _native.pad_rows(src_ptr, shape, pad_t.data_ptr(), row, device=dev, stream=stream)
                    src_t, src_ptr = pad_t, pad_t.data_ptr()
                _native.affine(src_ptr, shape, dst_ptr, shape, m, interp, flags, device=dev, stream=stream,
                               src_strides=src_strides)
            result = None if out_t is None else _native.download(out_t, stream)
            del src_t

        if profile:
            t1.record()
            t1.synchronize()
            print(f'transform finished in {t0.elapsed_time(t1):.3f}ms')
    return result
