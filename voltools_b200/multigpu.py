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

    def describe_shape(self, shape):
        from . import _native
        return (shape[0], shape[1], _native.padded_row(shape[2])), shape[2]

    def producer_stream(self):
        """Context: a side stream (ordered after the current one) for the prefilter steps and the broadcasts."""
        torch = self.torch
        if getattr(self, '_side', None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        self._side.wait_stream(torch.cuda.current_stream(self.device))
        return torch.cuda.stream(self._side)

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

    def _resident(self, buffer, interpolation, width, complete=True):
        """StaticVolume view of a received buffer (no copy).  complete=False: the buffer is still being filled
        (streaming), so no second layout may be derived from it: plain-layout kernels only."""
        from .volume import StaticVolume
        sv = StaticVolume.from_coefficients(buffer, interpolation, width)
        sv._plain_only = not complete
        return sv

    def resample_many(self, buffer, width, interpolation, matrices):
        return self._resident(buffer, interpolation, width).affine_many(matrices)

    def resample_many_range(self, buffer, width, interpolation, matrices, out, z0, z1):
        """Output planes [z0, z1) of every matrix into `out` (K, d0, d1, width); out-of-bounds voxels are zeroed."""
        from . import _native
        sv = self._resident(buffer, interpolation, width, complete=False)
        m = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
        with self.torch.cuda.device(self.dev):
            sv._launch(out.data_ptr(), m, _native.OOB_ZERO, self.torch.cuda.current_stream(self.dev).cuda_stream,
                       z_range=(z0, z1))

    def project_many(self, buffer, width, interpolation, matrices, z_range=None):
        """(K, d1, width) projections along axis 0 of the transformed volume (planes z_range of it), fused."""
        from .volume import StaticVolume
        sv = StaticVolume.from_coefficients(buffer, interpolation, width)
        return sv.project_many(matrices, z_range=z_range)

    def resample_slab(self, buffer, width, interpolation, matrix, z0, z1):
        from . import _native
        sv = self._resident(buffer, interpolation, width)
        d0, d1, d2 = sv.shape
        out = self.torch.empty((z1 - z0, d1, d2), dtype=self.torch.float32, device=self.device)
        if z1 > z0:
            # the C ABI addresses output plane z at d_dst + z*plane and touches only planes [z0, z1): handing it the
            # slab buffer shifted back by z0 planes makes it write the slab in place
            virtual_base = out.data_ptr() - z0 * d1 * d2 * 4
            m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(-1, 4, 4)
            with self.torch.cuda.device(self.dev):
                sv._launch(virtual_base, m, _native.OOB_ZERO, self.torch.cuda.current_stream(self.dev).cuda_stream,
                           z_range=(z0, z1))
        return out


def _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape=None, on_planes=None, chunks=None):
    """Root: upload + prefilter; everyone: receive the resident buffer.  Pipelined in z-chunks: the broadcast of the
    planes that are final runs (async, on the communicator's stream) while the root prefilters the next chunk.
    `shape`: the volume's shape if every rank knows it (saves the metadata broadcast, a host round trip).
    `on_planes(buffer, width, ready)`: called on every rank each time planes [0, ready) of the buffer have arrived
    (stream-ordered: work enqueued from it runs after them), so that consumers can start before the rest is there.
    Returns (buffer, width) on every rank."""
    rank = dist.get_rank(group)
    filtered = interpolation.startswith('filt')
    if shape is not None:
        buf_shape, width = engine.describe_shape(tuple(int(v) for v in shape))
    else:
        box = [engine.describe(volume, interpolation) if rank == src else None]
        dist.broadcast_object_list(box, src=src, group=group)
        buf_shape, width = box[0]
    buffer = engine.empty(tuple(buf_shape))
    plan = stream_plan(buf_shape[0], filtered, chunks)
    # The producer side (root: prefilter steps; everyone: the broadcasts) is enqueued from a side stream: a collective
    # is ordered after whatever its issuing stream already holds, so issuing it from the consumer's stream would make
    # broadcast i+1 wait for the resampling of chunk i.
    works = []
    with engine.producer_stream():
        step = engine.prepare_stream(volume, interpolation, buffer, plan) if rank == src else None
        for i, (_, _, z0, z1) in enumerate(plan):
            if step is not None:
                step(i)
            if z1 > z0:
                # NCCL over NVLink / NVSwitch on a GPU node
                works.append((dist.broadcast(buffer[z0:z1], src=src, group=group, async_op=True), z1))
    for w, z1 in works:
        w.wait()  # the consumer's stream waits for these planes; the host does not
        if on_planes is not None:
            on_planes(buffer, width, z1)
    return buffer, width


def streaming_margin(matrices: np.ndarray, interpolation: str):
    """If every matrix leaves axis 0 alone up to an integer shift t0 (the slice family of the kernels: rotations
    about axis 0 ...), output plane z only reads sampled planes z + t0 - m .. z + t0 + m (m = 0 linear, 1 cubic).
    Returns the number of sampled planes that must have arrived beyond output plane z (max t0 + m + 1), or None if
    some matrix mixes axis 0 with the others -- then nothing can be resampled before the whole volume is there."""
    m = np.asarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
    if not (np.all(m[:, 0, 0] == 1) and np.all(m[:, 0, 1:3] == 0) and np.all(m[:, 1:3, 0] == 0)):
        return None
    t0 = m[:, 0, 3]
    if not np.all(t0 == np.floor(t0)) or np.any(np.abs(t0) > 16384):
        return None
    return int(t0.max()) + (0 if interpolation == 'linear' else 1) + 1


def sweep(volume, matrices: Sequence[np.ndarray], interpolation: str = 'filt_bspline', src: int = 0, group=None,
          engine=None, shape=None, chunks=None, overlap: bool = False):
    """Batch of transforms of one volume, split across the ranks of `group`.

    volume: the samples on rank `src` (numpy / device array), ignored elsewhere (may be None).
    shape:  the volume's shape, if every rank knows it (skips a metadata broadcast).
    chunks: z-chunks of the pipelined prepare + broadcast (default: `stream_plan`'s choice).
    Returns (outputs, indices): `outputs[i]` is the volume for `matrices[indices[i]]`, resident on this rank.

    overlap=True: when all of this rank's matrices are of the slice family (the README's rotation sweep), the
    resampling is overlapped with the broadcast: as each z-chunk of coefficients arrives, the output planes it
    completes are produced for every matrix of the rank (z-range launches), so only the first chunk's latency is
    exposed.  Measured on B200s (256^3, 180 angles): the extra launches and warm-up planes of the z-range pieces cost
    more than the ~0.2 ms of broadcast they hide (2 GPUs: 445 vs 469 Gvox/s, 8 GPUs: 1355 vs 1483), so it is off by
    default; it pays when the broadcast is long against the resampling (few matrices per rank, large volumes).
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    mats = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
    mine = split_strided(len(mats), world, rank)
    my_mats = mats[mine.start::world]
    margin = streaming_margin(my_mats, interpolation) if (overlap and len(mine)) else None
    state = {'out': None, 'done': 0}

    def on_planes(buffer, width, ready):
        d0 = int(buffer.shape[0])
        if state['out'] is None:
            state['out'] = engine.empty((len(mine),) + tuple(buffer.shape[:2]) + (width,))
        upto = d0 if ready >= d0 else max(state['done'], min(d0, ready - margin))
        if upto > state['done']:
            engine.resample_many_range(buffer, width, interpolation, my_mats, state['out'], state['done'], upto)
            state['done'] = upto

    streaming = overlap and margin is not None and hasattr(engine, 'resample_many_range')
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape,
                                           on_planes if streaming else None, chunks)
    if streaming:
        return state['out'], list(mine)
    out = engine.resample_many(buffer, width, interpolation, my_mats) if len(mine) else \
        engine.empty((0,) + tuple(buffer.shape[:2]) + (width,))
    return out, list(mine)


def zslab_affine(volume, matrix: np.ndarray, interpolation: str = 'filt_bspline', src: int = 0, group=None,
                 engine=None, shape=None):
    """One transform of one (large) volume, the output split into z-slabs across the ranks of `group`.

    Returns (slab, (z0, z1)): this rank's output planes, resident on this rank.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape)
    z0, z1 = split_slabs(int(buffer.shape[0]), world, rank)
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(4, 4)
    return engine.resample_slab(buffer, width, interpolation, m, z0, z1), (z0, z1)


def project_sweep(volume, matrices: Sequence[np.ndarray], interpolation: str = 'filt_bspline', src: int = 0, group=None,
                  engine=None, shape=None, chunks=None):
    """Tilt series (examples/projections.py: `transform(rotation=...).sum(axis=0)` per angle), the matrices split across
    the ranks of `group` like `sweep`; the transformed volumes are never written (vt_project_strided_f32).
    Returns (projections, indices): `projections[i]` (d1, d2) belongs to `matrices[indices[i]]`, resident on this rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    mats = np.ascontiguousarray(matrices, dtype=np.float32).reshape(-1, 4, 4)
    mine = split_strided(len(mats), world, rank)
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape, None, chunks)
    if not len(mine):
        return engine.empty((0, int(buffer.shape[1]), width)), []
    return engine.project_many(buffer, width, interpolation, mats[mine.start::world]), list(mine)


def zslab_project(volume, matrix: np.ndarray, interpolation: str = 'filt_bspline', src: int = 0, group=None, engine=None,
                  shape=None):
    """ONE projection of one (large) volume under any matrix: rank r sums the output planes of its z-slab
    (`split_slabs`) with the fused kernels, then one all-reduce (NCCL, over NVLink) adds the partial images -- the only
    exchange step of the path besides the coefficient broadcast, d1 x d2 floats.  Returns the full (d1, d2) projection
    on every rank."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape)
    z0, z1 = split_slabs(int(buffer.shape[0]), world, rank)
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(1, 4, 4)
    part = engine.project_many(buffer, width, interpolation, m, z_range=(z0, z1))[0].contiguous()
    dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


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
