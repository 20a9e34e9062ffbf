"""Multi-GPU partitioning of the resampling path: one process per GPU, torch.distributed (NCCL) for the plumbing.

The reference has no multi-GPU layer at all (a single device string, voltools/utils/general.py:84-88).  The path
shards along the two axes that are independent by construction:

  * a batch of transforms of one volume (the README's 180-angle `StaticVolume` sweep, README.md:26-27):
    `sweep()` -- the root rank uploads and prefilters once, ONE broadcast ships the coefficient buffer to every
    rank over NVLink, then rank r resamples matrices `split_batch(K, world, r)`; outputs stay on the GPU that
    computed them.  No collective after the broadcast.
  * one large output volume: `zslab_affine()` -- same broadcast, then rank r produces output planes
    `split_slabs(d0, world, r)` with the z-range argument of the C ABI (every rank holds the whole coefficient
    volume: 4 GiB at 1024^3, trivial next to 180 GB of HBM, and it makes any matrix valid without computing
    per-slab input footprints).

Both take an `engine` so that the orchestration (partitioning, metadata + buffer broadcast, result placement) can be
exercised on CPU with the gloo backend and the oracle as the compute engine (tests/test_multigpu.py); the default
engine is the CUDA library and there is no CPU fallback in the product path.
"""
from typing import List, Sequence, Tuple

import numpy as np


def split_batch(n_items: int, world: int, rank: int) -> range:
    """Contiguous block of a batch of `n_items` for `rank`: sizes differ by at most one, earlier ranks get the
    larger blocks, every item belongs to exactly one rank."""
    base, extra = divmod(int(n_items), int(world))
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def split_slabs(d0: int, world: int, rank: int) -> Tuple[int, int]:
    """Output planes [z0, z1) of `rank` for z-slab sharding."""
    r = split_batch(d0, world, rank)
    return r.start, r.stop


class CudaEngine:
    """The product engine: StaticVolume on the rank's GPU."""

    def __init__(self, device_index: int):
        import torch
        self.torch = torch
        self.dev = device_index
        self.device = torch.device(f'cuda:{device_index}')

    def prepare(self, volume, interpolation):
        """Root only: upload + prefilter -> (resident buffer to broadcast, true width)."""
        from .volume import StaticVolume
        sv = StaticVolume(volume, interpolation=interpolation, device=f'gpu:{self.dev}')
        return sv.coefficient_buffer, sv.shape[2]

    def empty(self, shape):
        return self.torch.empty(shape, dtype=self.torch.float32, device=self.device)

    def resample_many(self, buffer, width, interpolation, matrices):
        from .volume import StaticVolume
        sv = StaticVolume.from_coefficients(buffer, interpolation, width)
        return sv.affine_many(matrices)

    def resample_slab(self, buffer, width, interpolation, matrix, z0, z1):
        from . import _native
        from .volume import StaticVolume
        sv = StaticVolume.from_coefficients(buffer, interpolation, width)
        d0, d1, d2 = sv.shape
        out = self.torch.empty((z1 - z0, d1, d2), dtype=self.torch.float32, device=self.device)
        if z1 > z0:
            # the C ABI addresses output plane z at d_dst + z*plane and touches only planes [z0, z1): handing it the
            # slab buffer shifted back by z0 planes makes it write the slab in place
            virtual_base = out.data_ptr() - z0 * d1 * d2 * 4
            with self.torch.cuda.device(self.dev):
                _native.affine(buffer.data_ptr(), sv.shape, virtual_base, sv.shape, matrix, sv._interp, _native.OOB_ZERO,
                               z_range=(z0, z1), device=self.dev,
                               stream=self.torch.cuda.current_stream(self.dev).cuda_stream, src_strides=sv._strides)
        return out


def _broadcast_buffer(engine, dist, group, src, buffer, meta):
    """Ships (shape of the resident buffer, true width) then the buffer itself from `src` to every rank."""
    box = [meta]
    dist.broadcast_object_list(box, src=src, group=group)
    buf_shape, width = box[0]
    if buffer is None:
        buffer = engine.empty(tuple(buf_shape))
    dist.broadcast(buffer, src=src, group=group)  # NCCL over NVLink / NVSwitch on a GPU node
    return buffer, width


def sweep(volume, matrices: Sequence[np.ndarray], interpolation: str = 'filt_bspline', src: int = 0, group=None,
          engine=None):
    """Batch of transforms of one volume, split across the ranks of `group`.

    volume: the samples on rank `src` (numpy / device array), ignored elsewhere (may be None).
    Returns (outputs, indices): `outputs[i]` is the volume for `matrices[indices[i]]`, resident on this rank.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    mats = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
    buffer, meta = None, None
    if rank == src:
        buffer, width = engine.prepare(volume, interpolation)
        meta = (tuple(buffer.shape), int(width))
    buffer, width = _broadcast_buffer(engine, dist, group, src, buffer, meta)
    mine = split_batch(len(mats), world, rank)
    out = engine.resample_many(buffer, width, interpolation, mats[mine.start:mine.stop]) if len(mine) else \
        engine.empty((0,) + tuple(buffer.shape[:2]) + (width,))
    return out, list(mine)


def zslab_affine(volume, matrix: np.ndarray, interpolation: str = 'filt_bspline', src: int = 0, group=None,
                 engine=None):
    """One transform of one (large) volume, the output split into z-slabs across the ranks of `group`.

    Returns (slab, (z0, z1)): this rank's output planes, resident on this rank.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    buffer, meta = None, None
    if rank == src:
        buffer, width = engine.prepare(volume, interpolation)
        meta = (tuple(buffer.shape), int(width))
    buffer, width = _broadcast_buffer(engine, dist, group, src, buffer, meta)
    z0, z1 = split_slabs(int(buffer.shape[0]), world, rank)
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(4, 4)
    return engine.resample_slab(buffer, width, interpolation, m, z0, z1), (z0, z1)


def gather_slabs(slab, group=None, dst: int = 0):
    """Convenience for tests / small volumes: concatenates the z-slabs on rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    shapes: List = [None] * world
    dist.all_gather_object(shapes, tuple(slab.shape), group=group)
    if rank != dst:
        if slab.numel():
            dist.send(slab.contiguous(), dst=dst, group=group)
        return None
    parts = []
    for r in range(world):
        if r == dst:
            parts.append(slab)
            continue
        buf = torch.empty(shapes[r], dtype=slab.dtype, device=slab.device)
        if buf.numel():
            dist.recv(buf, src=r, group=group)
        parts.append(buf)
    return torch.cat(parts, dim=0)
