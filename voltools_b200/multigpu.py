"""Multi-GPU partitioning of the resampling path: one process per GPU, torch.distributed (NCCL) for the plumbing.

The reference has no multi-GPU layer at all (a single device string, voltools/utils/general.py:84-88).  The path
shards along the two axes that are independent by construction:

  * a batch of transforms of one volume (the README's 180-angle `StaticVolume` sweep, README.md:26-27):
    `sweep()` -- the root rank uploads and prefilters once, ONE broadcast ships the coefficient buffer to every
    rank over NVLink, then rank r resamples matrices `split_strided(K, world, r)` (round robin); outputs stay on the GPU that
    computed them.  No collective after the broadcast.
  * one large output volume: `zslab_affine()` -- same broadcast, then rank r produces output planes
    `split_slabs(d0, world, r)` with the z-range argument of the C ABI (every rank holds the whole coefficient
    volume: 4 GiB at 1024^3, trivial next to 180 GB of HBM, and it makes any matrix valid without computing
    per-slab input footprints).

The prepare + broadcast step is pipelined in z-chunks (`stream_plan`): while NCCL ships the coefficient planes that
are already final, the root's compute stream prefilters the next chunk (vt_prefilter_planes_f32), so the root-side
cost is max(prefilter, broadcast) instead of their sum.

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


def split_strided(n_items: int, world: int, rank: int) -> range:
    """Round-robin share of a batch for `rank`: items rank, rank + world, ...  Used for rotation sweeps, where the
    cost of a transform varies smoothly with the angle (shared-memory bank conflicts peak around 45 degrees), so
    contiguous blocks of angles would leave the ranks unevenly loaded."""
    return range(int(rank), int(n_items), int(world))


def split_slabs(d0: int, world: int, rank: int) -> Tuple[int, int]:
    """Output planes [z0, z1) of `rank` for z-slab sharding."""
    r = split_batch(d0, world, rank)
    return r.start, r.stop


PREFILTER_LOOKAHEAD = 12  # planes of look-ahead of the Z prefilter (vt_prefilter_win.cu: K)


def stream_plan(d0: int, filtered: bool, chunks: int = None) -> List[Tuple[int, int, int, int]]:
    """z-chunks of the pipelined prepare + broadcast, identical on every rank: [(xy0, xy1, z0, z1)].
    Step i runs the XY prefilter of sample planes [xy0, xy1) and makes resident-buffer planes [z0, z1) final (the
    Z prefilter trails the XY passes by its look-ahead); [z0, z1) is what gets broadcast after step i."""
    d0 = int(d0)
    if chunks is None:
        chunks = max(1, min(8, d0 // 64))
    plan, done = [], 0
    for i in range(chunks):
        h0, h1 = d0 * i // chunks, d0 * (i + 1) // chunks
        zn = h1 if (not filtered or h1 == d0) else max(done, h1 - PREFILTER_LOOKAHEAD)
        plan.append((h0, h1, done, zn))
        done = zn
    return plan


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

    def describe(self, volume, interpolation):
        """Root only: (shape of the resident buffer, true width) without touching the data."""
        from . import _native
        d0, d1, d2 = (int(v) for v in volume.shape)
        return (d0, d1, _native.padded_row(d2)), d2

    def prepare_stream(self, volume, interpolation, buffer, plan):
        """Root only: returns step(i), which enqueues (on the current stream) the work that makes planes
        plan[i][2]:plan[i][3] of `buffer` final."""
        from . import _native
        from ._native import INTERPOLATIONS
        torch = self.torch
        _, filtered = INTERPOLATIONS[interpolation]
        raw = torch.as_tensor(volume, dtype=torch.float32, device=self.device).contiguous() \
            if not isinstance(volume, np.ndarray) else torch.from_numpy(np.ascontiguousarray(volume, np.float32)).to(self.device)
        shape = tuple(int(v) for v in raw.shape)
        row = int(buffer.shape[2])
        strides = (row, shape[1] * row)
        ws = torch.empty_like(buffer) if filtered else None

        def step(i):
            xy0, xy1, z0, z1 = plan[i]
            with torch.cuda.device(self.dev):
                if filtered:
                    _native.prefilter_planes(raw.data_ptr(), ws.data_ptr(), buffer.data_ptr(), shape, strides, (xy0, xy1),
                                             (z0, z1), self.dev, torch.cuda.current_stream(self.dev).cuda_stream)
                elif z1 > z0:
                    buffer[z0:z1, :, :shape[2]].copy_(raw[z0:z1])
                    if row != shape[2]:
                        buffer[z0:z1, :, shape[2]:].zero_()
        return step

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


def _prepare_and_broadcast(engine, dist, group, src, volume, interpolation):
    """Root: upload + prefilter; everyone: receive the resident buffer.  Pipelined in z-chunks: the broadcast of the
    planes that are final runs (async, on the communicator's stream) while the root prefilters the next chunk.
    Returns (buffer, width) on every rank."""
    rank = dist.get_rank(group)
    filtered = interpolation.startswith('filt')
    box = [engine.describe(volume, interpolation) if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    buf_shape, width = box[0]
    buffer = engine.empty(tuple(buf_shape))
    plan = stream_plan(buf_shape[0], filtered)
    step = engine.prepare_stream(volume, interpolation, buffer, plan) if rank == src else None
    works = []
    for i, (_, _, z0, z1) in enumerate(plan):
        if step is not None:
            step(i)
        if z1 > z0:
            # NCCL over NVLink / NVSwitch on a GPU node; ordered after the work enqueued so far on this stream
            works.append(dist.broadcast(buffer[z0:z1], src=src, group=group, async_op=True))
    for w in works:
        w.wait()
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
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation)
    mine = split_strided(len(mats), world, rank)
    out = engine.resample_many(buffer, width, interpolation, mats[mine.start::world]) if len(mine) else \
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
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation)
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
