"""StaticVolume: a volume kept resident on the GPU and resampled many times.

Mirror of voltools/volume.py:13-165.  What differs underneath:
  * the resident object is the (optionally prefiltered) coefficient volume in linear HBM -- no CUDA array /
    texture object (volume.py:36-50): the kernels gather from the linear buffer directly;
  * the matrix travels in kernel parameters (no per-call 64-byte H2D copy, volume.py:70);
  * `affine_many` / `transform_many` push a whole batch of matrices through one launch per VT_MAX_BATCH
    matrices (the reference launches once per matrix);
  * with output=None the zero-fill (volume.py:73) is fused into the kernel.
"""
from typing import Sequence, Tuple, Union

import numpy as np

from . import _native
from .transforms import _INTERPOLATIONS, _device_view, _is_host, _resolve_device, _stream, _torch
from .utils import get_available_devices, scale_matrix, shear_matrix, rotation_matrix, translation_matrix, \
    transform_matrix


class StaticVolume:
    """
    For StaticVolume transforms the boolean reshape cannot be given as an argument.
    """

    def __init__(self, data, interpolation: str = 'linear', device: str = 'gpu'):
        if len(data.shape) != 3:
            raise ValueError('Expected a 3D array')
        devices = get_available_devices()
        if device not in devices:
            raise ValueError(f'Unknown device ({device}), must be one of {devices}')
        if interpolation not in _INTERPOLATIONS:
            raise ValueError(f'Interpolation must be one of {list(_INTERPOLATIONS.keys())}')
        torch = _torch()
        self.device = device
        self.interpolation = interpolation
        self._interp, needs_prefilter = _INTERPOLATIONS[interpolation]
        vin = None if _is_host(data) else _device_view(data, 'data')
        self._dev = _resolve_device(device, vin)
        self.shape = tuple(int(s) for s in data.shape)
        self.d_type = np.float32
        # always a private copy (volume.py:30 `cp.array(data)` copies too)
        with torch.cuda.device(self._dev):
            if vin is None:
                host = np.ascontiguousarray(data, dtype=np.float32)
                self._coeffs = torch.from_numpy(host).to(f'cuda:{self._dev}')
            else:
                src = vin.owner if isinstance(vin.owner, torch.Tensor) \
                    else torch.as_tensor(vin.owner, device=f'cuda:{self._dev}')
                self._coeffs = src.to(f'cuda:{self._dev}', copy=True)
            # the resident buffer has rows padded to 16 bytes (TMA staging for any width); the prefilter writes that
            # layout directly, an unfiltered volume with an odd width is copied into it once
            raw = self._coeffs
            row = _native.padded_row(self.shape[2])
            self._strides = (row, self.shape[1] * row)
            if needs_prefilter:
                self._coeffs = torch.empty((self.shape[0], self.shape[1], row), dtype=torch.float32,
                                           device=f'cuda:{self._dev}')
                _native.prefilter(raw.data_ptr(), self.shape, self._dev, _stream(self._dev),
                                  dst_ptr=self._coeffs.data_ptr(), dst_strides=self._strides)
            elif row != self.shape[2]:
                self._coeffs = torch.empty((self.shape[0], self.shape[1], row), dtype=torch.float32,
                                           device=f'cuda:{self._dev}')
                _native.pad_rows(raw.data_ptr(), self.shape, self._coeffs.data_ptr(), row, device=self._dev,
                                 stream=_stream(self._dev))
            del raw

    # -- resident buffer access (used by the multi-GPU layer) -------------------------------------------
    @property
    def coefficients(self):
        """The resident (prefiltered, for filt_* modes) volume as a torch CUDA tensor view of shape `shape`
        (rows may be padded: the view is then non-contiguous)."""
        return self._coeffs[:, :, :self.shape[2]]

    @property
    def coefficient_buffer(self):
        """The resident buffer itself, shape (d0, d1, padded row); what the multi-GPU layer broadcasts."""
        return self._coeffs

    @classmethod
    def from_coefficients(cls, coeffs, interpolation: str = 'linear', width: int = None):
        """Wrap an already prepared coefficient buffer (e.g. one received by NCCL broadcast) without copying.
        `coeffs` is contiguous with shape (d0, d1, row); `width` <= row is the true extent of axis 2."""
        self = cls.__new__(cls)
        self.device = f'gpu:{coeffs.device.index}'
        self.interpolation = interpolation
        self._interp, _ = _INTERPOLATIONS[interpolation]
        self._dev = coeffs.device.index
        row = int(coeffs.shape[2])
        self.shape = (int(coeffs.shape[0]), int(coeffs.shape[1]), row if width is None else int(width))
        self._strides = (row, int(coeffs.shape[1]) * row)
        self.d_type = np.float32
        self._coeffs = coeffs
        return self

    # -- kernel family choice ---------------------------------------------------------------------------
    #: general-matrix transforms of one StaticVolume after which the texture family takes over.  Creating the CUDA
    #: array costs ~9 ms (cudaMalloc3DArray) + one 8 B/voxel upload; per 512^3 transform it then saves 0.3 ms
    #: (linear, 340 vs 193 Gvox/s) or 0.55 ms (bspline / filt_bspline, 66 vs 52 Gvox/s) over the brick kernels.
    TEXTURE_AFTER = 16

    def _z4_buffer(self, axis, stream):
        """The resident volume in the Z4 layout of `axis` (vt_resample_z4.cu), packed from the plain buffer on first
        use and kept: the second representation the reference keeps as a CUDA array (volume.py:37-50)."""
        z4 = getattr(self, '_z4', None)
        if z4 is None:
            z4 = self._z4 = {}
        buf = z4.get(axis)
        if buf is None:
            torch = _torch()
            buf = torch.empty(_native.z4_bytes(self.shape, axis) // 4, dtype=torch.float32, device=f'cuda:{self._dev}')
            _native.pack_z4(self._coeffs.data_ptr(), self.shape, buf.data_ptr(), axis, device=self._dev, stream=stream,
                            src_strides=self._strides)
            z4[axis] = buf
        return buf

    def plan(self, matrices):
        """Which kernel family `affine` / `affine_many` would run for these matrices on this volume (introspection):
        {'family': 'slice4', 'axis': m} for matrices that leave axis m alone, else 'brick' / 'gather' (or the texture
        family once the volume has switched to it)."""
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
        p = _native.launch_plan(self._coeffs.data_ptr(), self.shape, self.shape, m, self._interp, resident=True)
        if p['family'] != 'slice4' and getattr(self, '_tex', None) not in (None, False):
            p = {'family': 'texture'}
        return p

    def _launch(self, dst_ptr, m, flags, stream, z_range=None):
        """One launch set for the matrices `m` (K, 4, 4) on the resident volume.

        Matrices that leave one axis alone (rotations about an axis through the centre ...) run the slice4 kernels
        on the Z4 layout of that axis.

        Matrices of the slice family (rotations about axis 0 ...) always run the plane-marching kernels.  For the
        two interpolators that ARE the texture unit (linear, bspline / filt_bspline) a general matrix runs on a
        hardware texture object once this volume has seen TEXTURE_AFTER such transforms -- the reference keeps the
        same second copy (volume.py:37-50); until then, and always for *_simple, the shared-memory brick kernels."""
        axis = -1
        if _native.z4_wanted(self._interp, True) and not getattr(self, '_plain_only', False):  # (resident: always)
            axis = _native.z4_axis(self.shape, self.shape, m, self._interp)
        if axis >= 0:
            _native.affine_z4(self._z4_buffer(axis, stream).data_ptr(), axis, self.shape, dst_ptr, self.shape, m,
                              self._interp, flags, z_range=z_range, device=self._dev, stream=stream)
            return
        use_tex = False
        if self._interp in (_native.LINEAR, _native.CUBIC_TEX) and getattr(self, '_tex', None) is not False:
            if _native.affine_plan(self._coeffs.data_ptr(), self.shape, self.shape, m, self._interp) != 'slice':
                self._general = getattr(self, '_general', 0) + len(m)
                use_tex = getattr(self, '_tex', None) is not None or self._general >= self.TEXTURE_AFTER
        if use_tex and getattr(self, '_tex', None) is None:
            try:
                self._tex = _native.Texture(self._coeffs.data_ptr(), self.shape, self._strides, self._dev, stream)
            except RuntimeError:
                self._tex, use_tex = False, False  # extent beyond the 3-D texture limits / no memory: stay on bricks
        if use_tex:
            self._tex.affine(dst_ptr, self.shape, m, self._interp, flags, z_range=z_range, stream=stream)
        else:
            _native.affine(self._coeffs.data_ptr(), self.shape, dst_ptr, self.shape, m, self._interp, flags,
                           z_range=z_range, device=self._dev, stream=stream, src_strides=self._strides)

    # -- transforms ---------------------------------------------------------------------------------------
    def affine(self, transform_m: np.ndarray, profile: bool = False, output=None) -> Union[np.ndarray, None]:
        """volume.py:61-101."""
        torch = _torch()
        m = np.ascontiguousarray(transform_m, dtype=np.float32).reshape(4, 4)
        vout = None if output is None else _device_view(output, 'output')
        if vout is not None and vout.shape != self.shape:
            raise ValueError(f'output shape {vout.shape} does not match the volume shape {self.shape}')
        with torch.cuda.device(self._dev):
            if profile:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
            stream = _stream(self._dev)
            if vout is None:
                out_t = torch.empty(self.shape, dtype=torch.float32, device=f'cuda:{self._dev}')
                self._launch(out_t.data_ptr(), m[None], _native.OOB_ZERO, stream)
            else:
                self._launch(vout.ptr, m[None], _native.OOB_SKIP, stream)
            if profile:
                t1.record()
                t1.synchronize()
                print(f'transform finished in {t0.elapsed_time(t1):.3f}ms')
            return _native.download(out_t, stream) if vout is None else None

    def affine_many(self, matrices: Sequence[np.ndarray], output=None, zero_fill: bool = None):
        """Batched `affine`: K matrices -> K volumes, one launch per VT_MAX_BATCH matrices.

        output: None -> returns a new torch CUDA tensor of shape (K, *shape), zero where out of bounds;
                a (K, *shape) device array -> written in place (out-of-bounds voxels untouched unless
                zero_fill=True), returns None;
                a (K, *shape) float32 C-contiguous NUMPY array -> the K results land in host memory (K `.get()`s of
                volume.py:99, pipelined: the kernels of chunk c+1 run while chunk c crosses PCIe), every voxel
                overwritten (zero where out of bounds), returns None.  Give it page-locked memory
                (`voltools_b200.pinned_empty`) and the copies run at link rate; a pageable array is filled through
                two pinned staging buffers and a host memcpy.
        """
        torch = _torch()
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
        k = len(m)
        if isinstance(output, np.ndarray):
            return self._affine_many_host(m, output)
        with torch.cuda.device(self._dev):
            stream = _stream(self._dev)
            if output is None:
                out_t = torch.empty((k,) + self.shape, dtype=torch.float32, device=f'cuda:{self._dev}')
                ptr, flags = out_t.data_ptr(), _native.OOB_ZERO
            else:
                vout = _device_view(output, 'output')
                if vout.shape != (k,) + self.shape:
                    raise ValueError(f'output shape {vout.shape} does not match {(k,) + self.shape}')
                out_t, ptr = None, vout.ptr
                flags = _native.OOB_ZERO if zero_fill else _native.OOB_SKIP
            self._launch(ptr, m, flags, stream)
        return out_t

    HOST_CHUNK_BYTES = 128 << 20   # device staging per chunk of the host-output pipeline (three in flight)

    def _affine_many_host(self, m, out):
        """K results into the host array `out`: a ring of three device chunks; the kernels write chunk c on the
        caller's stream, a copy stream ships it while the kernels of chunk c+1 run."""
        torch = _torch()
        k = len(m)
        if out.shape != (k,) + self.shape or out.dtype != np.float32 or not out.flags['C_CONTIGUOUS'] \
                or not out.flags['WRITEABLE']:
            raise ValueError(f'output must be a writeable C-contiguous float32 array of shape {(k,) + self.shape}')
        if k == 0:
            return None
        host_t = torch.from_numpy(out)
        pinned = host_t.is_pinned()
        vol_bytes = 4 * int(np.prod(self.shape))
        per = int(max(1, min(k, self.HOST_CHUNK_BYTES // max(vol_bytes, 1))))
        n_chunks = -(-k // per)
        with torch.cuda.device(self._dev):
            cur = torch.cuda.current_stream(self._dev)
            cp = getattr(self, '_copy_stream', None)
            if cp is None:
                cp = self._copy_stream = torch.cuda.Stream(device=self._dev)
            slots = [torch.empty((per,) + self.shape, dtype=torch.float32, device=f'cuda:{self._dev}')
                     for _ in range(min(3, n_chunks))]
            slot_free = [None] * len(slots)
            stage = [] if pinned else [torch.empty((per,) + self.shape, dtype=torch.float32, pin_memory=True)
                                       for _ in range(min(2, n_chunks))]
            pending = []   # pageable destination: (copy-done event, staging index, k0, k1)

            def drain_one():
                ev, si, a, b = pending.pop(0)
                ev.synchronize()
                np.copyto(out[a:b], stage[si][:b - a].numpy())

            try:
                for c in range(n_chunks):
                    k0, k1 = c * per, min(k, (c + 1) * per)
                    s = c % len(slots)
                    if slot_free[s] is not None:
                        cur.wait_event(slot_free[s])
                    self._launch(slots[s].data_ptr(), m[k0:k1], _native.OOB_ZERO, cur.cuda_stream)
                    ready = torch.cuda.Event()
                    ready.record(cur)
                    cp.wait_event(ready)
                    if not pinned and len(pending) == len(stage):
                        drain_one()
                    with torch.cuda.stream(cp):
                        dst = host_t[k0:k1] if pinned else stage[c % len(stage)][:k1 - k0]
                        dst.copy_(slots[s][:k1 - k0], non_blocking=True)
                        done = torch.cuda.Event()
                        done.record(cp)
                    slot_free[s] = done
                    if not pinned:
                        pending.append((done, c % len(stage), k0, k1))
                while pending:
                    drain_one()
            finally:
                # also on an error path: no copy into `out` (or out of the ring) may still be running when the caller
                # gets control back and is free to release either
                cp.synchronize()
        return None

    # -- rotate-and-project (examples/projections.py:20-26: `transform(...).sum(axis=0)`, fused) ----------
    def _project(self, ptr, m, z_range, stream):
        torch = _torch()
        use_tex = getattr(self, '_tex', None) not in (None, False) and \
            self._interp in (_native.LINEAR, _native.CUBIC_TEX) and \
            _native.affine_plan(self._coeffs.data_ptr(), self.shape, self.shape, m, self._interp) != 'slice'
        if use_tex:
            self._tex.project(ptr, self.shape, m, self._interp, z_range=z_range, stream=stream)
            return
        ws = getattr(self, '_project_ws', None)
        if ws is None:
            ws = self._project_ws = torch.empty(_native.project_workspace_bytes(self.shape), dtype=torch.uint8,
                                                device=f'cuda:{self._dev}')
        _native.project(self._coeffs.data_ptr(), self.shape, ptr, self.shape, m, self._interp, z_range=z_range,
                        device=self._dev, stream=stream, src_strides=self._strides, workspace_ptr=ws.data_ptr(),
                        workspace_bytes=ws.numel())

    def project_many(self, matrices: Sequence[np.ndarray], output=None, z_range=None):
        """Projections along axis 0 of the volume transformed by each of K matrices: what the reference computes as
        `static_volume.affine(m).sum(axis=0)` per matrix, without writing (or reading back) the transformed volumes.

        Matrices that leave axis 0 alone -- every tilt `rotation=(i, 0, 0), rotation_order='sxyz'` of the reference's
        example -- cost ONE read of the resident volume plus a 2-D resample per matrix; other matrices run the
        resampling kernels with a register accumulator and one atomic add per thread.

        output: None -> returns a new torch CUDA tensor (K, d1, d2); a (K, d1, d2) device array -> overwritten,
        returns None.  z_range=(z0, z1) sums output planes z0..z1-1 only (partial projections of z-slabs add up).
        """
        torch = _torch()
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
        pshape = (len(m), self.shape[1], self.shape[2])
        with torch.cuda.device(self._dev):
            if output is None:
                out_t = torch.empty(pshape, dtype=torch.float32, device=f'cuda:{self._dev}')
                ptr = out_t.data_ptr()
            else:
                vout = _device_view(output, 'output')
                if vout.shape != pshape:
                    raise ValueError(f'output shape {vout.shape} does not match {pshape}')
                out_t, ptr = None, vout.ptr
            self._project(ptr, m, z_range, _stream(self._dev))
        return out_t

    def affine_project(self, transform_m: np.ndarray, output=None) -> Union[np.ndarray, None]:
        """`self.affine(transform_m).sum(axis=0)` fused: returns the (d1, d2) projection as numpy, like affine(); or
        overwrites the (d1, d2) device array `output` and returns None."""
        m = np.ascontiguousarray(transform_m, dtype=np.float32).reshape(1, 4, 4)
        if output is None:
            return self.project_many(m)[0].cpu().numpy()
        vout = _device_view(output, 'output')
        if vout.shape != self.shape[1:]:
            raise ValueError(f'output shape {vout.shape} does not match {self.shape[1:]}')
        with _torch().cuda.device(self._dev):
            self._project(vout.ptr, m, None, _stream(self._dev))
        return None

    def project(self, scale: Union[float, Tuple[float, float, float], np.ndarray] = None,
                shear: Union[float, Tuple[float, float, float], np.ndarray] = None,
                rotation: Union[Tuple[float, float, float], np.ndarray] = None,
                rotation_units: str = 'deg', rotation_order: str = 'rzxz',
                translation: Union[Tuple[float, float, float], np.ndarray] = None,
                center: Union[Tuple[float, float, float], np.ndarray] = None,
                output=None) -> Union[np.ndarray, None]:
        """`self.transform(...same arguments...).sum(axis=0)` fused (arguments as volume.py:103-113)."""
        if center is None:
            center = np.divide(np.subtract(self.shape, 1), 2, dtype=np.float32)
        if isinstance(scale, float):
            scale = (scale, scale, scale)
        if isinstance(shear, float):
            shear = (shear, shear, shear)
        m = transform_matrix(scale, shear, rotation, rotation_units, rotation_order, translation, center)
        return self.affine_project(m, output)

    def transform(self, scale: Union[float, Tuple[float, float, float], np.ndarray] = None,
                  shear: Union[float, Tuple[float, float, float], np.ndarray] = None,
                  rotation: Union[Tuple[float, float, float], np.ndarray] = None,
                  rotation_units: str = 'deg', rotation_order: str = 'rzxz',
                  translation: Union[Tuple[float, float, float], np.ndarray] = None,
                  center: Union[Tuple[float, float, float], np.ndarray] = None,
                  profile: bool = False,
                  output=None) -> Union[np.ndarray, None]:
        """volume.py:103-123."""
        if center is None:
            center = np.divide(np.subtract(self.shape, 1), 2, dtype=np.float32)
        if isinstance(scale, float):
            scale = (scale, scale, scale)
        if isinstance(shear, float):
            shear = (shear, shear, shear)
        m = transform_matrix(scale, shear, rotation, rotation_units, rotation_order, translation, center)
        return self.affine(m, profile, output)

    def translate(self, translation: Tuple[float, float, float], profile: bool = False,
                  output=None) -> Union[np.ndarray, None]:
        """volume.py:125-131."""
        return self.affine(translation_matrix(translation), profile, output)

    def shear(self, coefficients: Union[float, Tuple[float, float, float]], profile: bool = False,
              output=None) -> Union[np.ndarray, None]:
        """volume.py:133-143."""
        if isinstance(coefficients, float):
            coefficients = (coefficients, coefficients, coefficients)
        return self.affine(shear_matrix(coefficients), profile, output)

    def scale(self, coefficients: Union[float, Tuple[float, float, float]], profile: bool = False,
              output=None) -> Union[np.ndarray, None]:
        """volume.py:145-155."""
        if isinstance(coefficients, float):
            coefficients = (coefficients, coefficients, coefficients)
        return self.affine(scale_matrix(coefficients), profile, output)

    def rotate(self, rotation: Tuple[float, float, float], rotation_units: str = 'deg', rotation_order: str = 'rzxz',
               profile: bool = False, output=None) -> Union[np.ndarray, None]:
        """volume.py:157-165 (about the array origin, like the reference)."""
        m = rotation_matrix(rotation=rotation, rotation_units=rotation_units, rotation_order=rotation_order)
        return self.affine(m, profile, output)
