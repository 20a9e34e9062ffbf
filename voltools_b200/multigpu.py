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


def _rank_world(dist, group):
    """(rank, world) of this process; a process that never initialised torch.distributed is a world of one (every
    function here then degenerates to its single-GPU form: no collective is issued)."""
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


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

    def join_producer(self):
        """The current stream waits for everything enqueued on the producer stream so far."""
        if getattr(self, '_side', None) is not None:
            self.torch.cuda.current_stream(self.device).wait_stream(self._side)

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

    def to_device(self, volume):
        """The raw samples as a contiguous float32 tensor on this rank's GPU (root only)."""
        torch = self.torch
        if isinstance(volume, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(volume, np.float32)).to(self.device)
        return torch.as_tensor(volume, dtype=torch.float32, device=self.device).contiguous()

    def prefilter_slab(self, raw_slab, buffer, shape, interpolation, xy0, xy1, z0, z1):
        """`raw_slab` holds sample planes [xy0, xy1) of a volume of `shape`; makes planes [z0, z1) of the full-size
        resident `buffer` (padded rows) final.  The windowed prefilter treats xy1 as an artificial end 12 planes past z1
        (vt_prefilter_planes_f32), so a z-slab of coefficients costs its own planes plus a 12-plane halo of samples."""
        from . import _native
        from ._native import INTERPOLATIONS
        torch = self.torch
        _, filtered = INTERPOLATIONS[interpolation]
        d0, d1, d2 = (int(v) for v in shape)
        row = int(buffer.shape[2])
        if not filtered:
            if z1 > z0:
                buffer[z0:z1, :, :d2].copy_(raw_slab[z0 - xy0:z1 - xy0])
                if row != d2:
                    buffer[z0:z1, :, d2:].zero_()
            return
        ws = torch.empty((xy1 - xy0, d1, row), dtype=torch.float32, device=self.device)
        # the ABI addresses planes absolutely: hand it the slab buffers shifted back by xy0 planes
        raw_base = raw_slab.data_ptr() - xy0 * d1 * d2 * 4
        ws_base = ws.data_ptr() - xy0 * d1 * row * 4
        with torch.cuda.device(self.dev):
            _native.prefilter_planes(raw_base, ws_base, buffer.data_ptr(), (xy1, d1, d2), (row, d1 * row), (xy0, xy1),
                                     (z0, z1), self.dev, torch.cuda.current_stream(self.dev).cuda_stream)

    def timer(self):
        """A stream-ordered time stamp: returns a callable giving milliseconds since `since` (another stamp)."""
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record(self.torch.cuda.current_stream(self.device))
        return ev

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


def _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape=None, on_planes=None, chunks=None):  # noqa: E501
    """Root: upload + prefilter; everyone: receive the resident buffer.  Pipelined in z-chunks: the broadcast of the
    planes that are final runs (async, on the communicator's stream) while the root prefilters the next chunk.
    `shape`: the volume's shape if every rank knows it (saves the metadata broadcast, a host round trip).
    `on_planes(buffer, width, ready)`: called on every rank each time planes [0, ready) of the buffer have arrived
    (stream-ordered: work enqueued from it runs after them), so that consumers can start before the rest is there.
    Returns (buffer, width) on every rank."""
    rank, world = _rank_world(dist, group)
    filtered = interpolation.startswith('filt')
    if shape is not None:
        buf_shape, width = engine.describe_shape(tuple(int(v) for v in shape))
    else:
        box = [engine.describe(volume, interpolation) if rank == src else None]
        if world > 1:
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
            if z1 > z0 and world > 1:
                # NCCL over NVLink / NVSwitch on a GPU node
                works.append((dist.broadcast(buffer[z0:z1], src=src, group=group, async_op=True), z1))
    for w, z1 in works:
        w.wait()  # the consumer's stream waits for these planes; the host does not
        if on_planes is not None:
            on_planes(buffer, width, z1)
    if not works:  # a world of one: nothing was shipped, the consumer's stream just waits for the producer's
        join = getattr(engine, 'join_producer', None)
        if join is not None:
            join()
        if on_planes is not None:
            on_planes(buffer, width, int(buf_shape[0]))
    return buffer, width


def prepare_and_broadcast(engine, volume, interpolation, src: int = 0, group=None, shape=None, chunks=None):
    """Public form of the sweep's first phase: the root uploads + prefilters, one (z-chunk pipelined) NCCL broadcast
    ships the coefficient buffer; returns (buffer, width) on every rank -- wrap it with
    StaticVolume.from_coefficients(buffer, interpolation, width)."""
    import torch.distributed as dist
    return _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape, None, chunks)


# ----------------------------------------------------------------------------------------------------
# z-slab sharding with per-slab input footprints
# ----------------------------------------------------------------------------------------------------
def _clip_image_bbox(m64, obox, z_lo, z_hi):
    """Bounding box (float, input index space) of {M a : a in the output index box} intersected with the input slab
    z_lo <= p0 <= z_hi, or None if empty.  The image is a parallelepiped: its intersection with the slab is spanned by
    the vertices inside it and the points where its 12 edges cross the two bounding planes."""
    corners = np.array([[a0, a1, a2, 1.0] for a0 in obox[0] for a1 in obox[1] for a2 in obox[2]])
    p = corners @ m64[:3].T  # (8, 3)
    pts = [v for v in p if z_lo <= v[0] <= z_hi]
    for i in range(8):
        for j in range(i + 1, 8):
            if bin(i ^ j).count('1') != 1:
                continue
            d = p[j][0] - p[i][0]
            if d == 0.0:
                continue
            for zc in (z_lo, z_hi):
                t = (zc - p[i][0]) / d
                if 0.0 <= t <= 1.0:
                    pts.append(p[i] + t * (p[j] - p[i]))
    if not pts:
        return None
    pts = np.asarray(pts)
    return pts.min(axis=0), pts.max(axis=0)


def slab_footprint_boxes(matrix, shape, world: int, halo: int = 3):
    """Which part of the sampled volume each output z-slab reads, owner by owner.

    The sampled (coefficient) volume is owned in z-slabs I_r = split_slabs(d0, world, r); output slab q =
    split_slabs(d0, world, q) samples the affine image of its index box, a parallelepiped.  Returns {(r, q): (lo, hi)}:
    the integer box [lo, hi) inside I_r that contains every texel rank q's kernels can multiply into a result -- the
    bounding box of (image of slab q) intersected with I_r, grown by `halo` texels (filter support: taps reach
    floor(p) - 1 .. floor(p) + 2) and clipped to the volume; pairs with nothing to send are absent.  Pure host
    arithmetic, identical on every rank."""
    m64 = np.asarray(matrix, dtype=np.float64).reshape(4, 4)
    d = [int(v) for v in shape]
    boxes = {}
    for q in range(world):
        z0, z1 = split_slabs(d[0], world, q)
        if z1 <= z0:
            continue
        obox = ((z0, z1 - 1), (0, d[1] - 1), (0, d[2] - 1))
        for r in range(world):
            i0, i1 = split_slabs(d[0], world, r)
            if i1 <= i0:
                continue
            bb = _clip_image_bbox(m64, obox, i0 - halo - 1, i1 + halo)
            if bb is None:
                continue
            lo = np.floor(bb[0]).astype(np.int64) - halo
            hi = np.ceil(bb[1]).astype(np.int64) + halo + 1
            lo = np.maximum(lo, 0)
            hi = np.minimum(hi, d)
            lo[0], hi[0] = max(lo[0], i0), min(hi[0], i1)
            if np.all(hi > lo):
                boxes[(r, q)] = (tuple(int(v) for v in lo), tuple(int(v) for v in hi))
    return boxes


def footprint_block_size(buf_shape, world: int):
    """Block edge lengths (bz, by, bx) for `slab_footprint_blocks`, or None: blocks must tile the resident buffer
    exactly and must not straddle two owners' z-slabs."""
    d0, d1, row = (int(v) for v in buf_shape)
    pick = []
    for extent, cands, unit in ((d0, (32, 16, 8), world), (d1, (64, 32, 16), 1), (row, (64, 32, 16), 1)):
        b = next((c for c in cands if extent % (c * unit) == 0), None)
        if b is None:
            return None
        pick.append(b)
    return tuple(pick)


def slab_footprint_blocks(matrix, buf_shape, width: int, world: int, block, halo: int = 3):
    """Block-sparse form of `slab_footprint_boxes`: the resident buffer (d0, d1, row) is tiled by `block` =
    (bz, by, bx) texels; returns {(r, q): int64 array (n, 3) of block indices} -- the blocks owned by rank r (its z-slab
    of the sampled volume) that output slab q can sample.  A block, grown by `halo` texels of filter support, is needed
    iff its pre-image under the matrix meets the slab's output index box; the pre-image's extent along each output axis
    is attained at the block's corners (the inverse map is affine), so the test is exact.  An oblique slab of a 1024^3
    volume under BASELINE's full affine needs 20 % of the volume this way (its bounding boxes: 57 %)."""
    m64 = np.asarray(matrix, dtype=np.float64).reshape(4, 4)
    inv = np.linalg.inv(m64)
    d0, d1, row = (int(v) for v in buf_shape)
    bz, by, bx = block
    nb = (d0 // bz, d1 // by, row // bx)
    g = np.stack(np.meshgrid(*[np.arange(n) for n in nb], indexing='ij'), axis=-1).reshape(-1, 3)
    lo = g * np.array(block) - halo
    hi = lo + np.array(block) + 2 * halo - 1  # inclusive texel range the block's taps may serve
    mins = np.full((len(g), 3), np.inf)
    maxs = -mins
    for corner in range(8):
        sel = np.array([(corner >> 2) & 1, (corner >> 1) & 1, corner & 1])
        p = np.where(sel == 1, hi, lo).astype(np.float64)
        a = p @ inv[:3, :3].T + inv[:3, 3]
        mins = np.minimum(mins, a)
        maxs = np.maximum(maxs, a)
    out_hi = np.array([d0 - 1, d1 - 1, width - 1], dtype=np.float64)
    inside = np.all(maxs[:, 1:] >= -1.0, axis=1) & np.all(mins[:, 1:] <= out_hi[1:] + 1.0, axis=1)
    slab_planes = d0 // world
    owner = (g[:, 0] * bz) // slab_planes
    res = {}
    for q in range(world):
        z0, z1 = split_slabs(d0, world, q)
        need = inside & (maxs[:, 0] >= z0 - 1.0) & (mins[:, 0] <= z1)
        for r in range(world):
            sel = g[need & (owner == r)]
            if len(sel):
                res[(r, q)] = sel.astype(np.int64)
    return res


def footprint_fraction(boxes, shape, world: int, block=None):
    """Largest share of the volume any rank has to RECEIVE from the others under `boxes` (or block lists)."""
    vol = float(np.prod([int(v) for v in shape]))
    worst = 0.0
    for q in range(world):
        if block is None:
            got = sum(float(np.prod(np.subtract(hi, lo))) for (r, qq), (lo, hi) in boxes.items() if qq == q and r != q)
        else:
            got = sum(float(len(idx)) * float(np.prod(block)) for (r, qq), idx in boxes.items() if qq == q and r != q)
        worst = max(worst, got / vol)
    return worst


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
    rank, world = _rank_world(dist, group)
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


_FOOTPRINT_PLANS = {}


def zslab_affine(volume, matrix: np.ndarray, interpolation: str = 'filt_bspline', src: int = 0, group=None,
                 engine=None, shape=None, timings: dict = None, footprint='auto'):
    """One transform of one (large) volume, the output split into z-slabs across the ranks of `group`.

    The raw volume lives on rank `src`.  Two ways of getting every rank what its slab samples:
      * footprint path (needs `shape` on every rank): the root scatters raw z-slabs (+ 12 planes of prefilter halo),
        every rank prefilters ITS slab (vt_prefilter_planes_f32), then the ranks exchange, pairwise and all at once,
        only the boxes of coefficients each output slab reads from each owner (`slab_footprint_boxes`: the affine
        image of the slab, owner by owner, + filter halo).  No rank ever holds more than its slab's input footprint;
        the root sends (world-1)/world of the volume once instead of broadcasting all of it.
      * broadcast path: the root prefilters everything and one NCCL broadcast ships the whole coefficient volume --
        used when the footprints are nearly the whole volume anyway (`footprint='auto'`: largest receive share >= 0.9
        minus the rank's own slab) or when the shape is not known everywhere.
    timings (optional dict): filled with 'distribute_ms' / 'resample_ms' callables (CUDA events on the engine's stream;
    call them after a synchronize) and 'info'.
    Returns (slab, (z0, z1)): this rank's output planes, resident on this rank.
    """
    import torch.distributed as dist
    rank, world = _rank_world(dist, group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(4, 4)
    stamp = getattr(engine, 'timer', None)
    t0 = stamp() if (timings is not None and stamp) else None
    info = {'path': 'broadcast'}
    use_fp = bool(footprint) and shape is not None and world > 1 and hasattr(engine, 'prefilter_slab')
    if use_fp:
        shape = tuple(int(v) for v in shape)
        buf_shape, width = engine.describe_shape(shape)
        block = footprint_block_size(buf_shape, world)
        # the plan is a pure function of (matrix, shape, world): a few ms of host arithmetic, memoised
        key = (m.tobytes(), tuple(buf_shape), width, world, block)
        plan = _FOOTPRINT_PLANS.get(key)
        if plan is None:
            if block is not None:
                boxes = slab_footprint_blocks(m, buf_shape, width, world, block)
                frac = footprint_fraction(boxes, buf_shape, world, block)
            else:  # the blocks do not tile this shape: one bounding box per (owner, slab) pair
                boxes = slab_footprint_boxes(m, shape, world)
                frac = footprint_fraction(boxes, shape, world)
            if len(_FOOTPRINT_PLANS) >= 16:
                _FOOTPRINT_PLANS.clear()
            plan = _FOOTPRINT_PLANS[key] = (boxes, frac)
        boxes, frac = plan
        info['largest_receive_share'] = frac
        info['footprint_block'] = block
        if footprint == 'auto' and frac + 1.0 / world >= 0.9:
            use_fp = False
    if use_fp:
        info['path'] = 'footprint'
        buffer, width = _scatter_prefilter_exchange(engine, dist, group, src, volume, interpolation, shape, boxes, block)
    else:
        buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape)
    t1 = stamp() if t0 is not None else None
    z0, z1 = split_slabs(int(buffer.shape[0]), world, rank)
    slab = engine.resample_slab(buffer, width, interpolation, m, z0, z1)
    if t0 is not None:
        t2 = stamp()
        timings['distribute_ms'] = lambda: t0.elapsed_time(t1)
        timings['resample_ms'] = lambda: t1.elapsed_time(t2)
    if timings is not None:
        timings['info'] = info
    return slab, (z0, z1)


def _scatter_prefilter_exchange(engine, dist, group, src, volume, interpolation, shape, boxes, block=None):
    """Footprint path of `zslab_affine`; returns (full-size resident buffer of which only this rank's footprint is
    filled, width)."""
    rank, world = _rank_world(dist, group)
    filtered = interpolation.startswith('filt')
    d0, d1, d2 = shape
    halo = PREFILTER_LOOKAHEAD if filtered else 0
    buf_shape, width = engine.describe_shape(shape)
    buffer = engine.empty(tuple(buf_shape))

    def raw_range(r):
        i0, i1 = split_slabs(d0, world, r)
        return max(0, i0 - halo), min(d0, i1 + halo)

    # 1) raw z-slabs (+ prefilter halo) leave the root: contiguous plane ranges, no packing
    xy0, xy1 = raw_range(rank)
    ops = []
    if rank == src:
        raw = engine.to_device(volume)
        raw_slab = raw[xy0:xy1]
        for r in range(world):
            a, b = raw_range(r)
            if r != src and b > a:
                ops.append(dist.P2POp(dist.isend, raw[a:b], _global_rank(dist, group, r), group))
    else:
        raw_slab = engine.empty((xy1 - xy0, d1, d2))
        if xy1 > xy0:
            ops.append(dist.P2POp(dist.irecv, raw_slab, _global_rank(dist, group, src), group))
    works = dist.batch_isend_irecv(ops) if ops else []
    if rank != src:
        for w in works:
            w.wait()
    # 2) every rank prefilters its own slab straight into its place in the full-size buffer (the root while its sends
    #    are still leaving: they only read `raw`)
    i0, i1 = split_slabs(d0, world, rank)
    if i1 > i0:
        engine.prefilter_slab(raw_slab, buffer, shape, interpolation, xy0, xy1, i0, i1)
    if rank == src:
        for w in works:
            w.wait()
    del raw_slab
    # 3) pairwise exchange of what each output slab reads from each owner, all pairs in one group
    ops, sends, recvs = [], [], []
    if block is not None:
        import torch
        bz, by, bx = block
        D0, D1, ROW = (int(v) for v in buffer.shape)
        # (Zb, Yb, Xb, bz, by, bx) view of the buffer: indexing its first three axes gathers / scatters whole blocks
        blocks = buffer.view(D0 // bz, bz, D1 // by, by, ROW // bx, bx).permute(0, 2, 4, 1, 3, 5)

        def index(idx):
            t = torch.from_numpy(idx).to(buffer.device, non_blocking=True)
            return t[:, 0], t[:, 1], t[:, 2]
    for q in range(world):
        bxs = boxes.get((rank, q))
        if q != rank and bxs is not None:
            if block is not None:
                t = blocks[index(bxs)]  # (n, bz, by, bx), contiguous
            else:
                lo, hi = bxs
                t = buffer[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].contiguous()
            sends.append(t)
            ops.append(dist.P2POp(dist.isend, t, _global_rank(dist, group, q), group))
    for r in range(world):
        bxs = boxes.get((r, rank))
        if r != rank and bxs is not None:
            if block is not None:
                t = engine.empty((len(bxs),) + tuple(block))
            else:
                t = engine.empty(tuple(h - l for l, h in zip(*bxs)))
            recvs.append((t, bxs))
            ops.append(dist.P2POp(dist.irecv, t, _global_rank(dist, group, r), group))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()
    for t, bxs in recvs:
        if block is not None:
            blocks[index(bxs)] = t
        else:
            lo, hi = bxs
            buffer[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].copy_(t)
    del sends, recvs
    return buffer, width


def _global_rank(dist, group, r):
    return r if group is None else dist.get_global_rank(group, r)


def project_sweep(volume, matrices: Sequence[np.ndarray], interpolation: str = 'filt_bspline', src: int = 0, group=None,
                  engine=None, shape=None, chunks=None):
    """Tilt series (examples/projections.py: `transform(rotation=...).sum(axis=0)` per angle), the matrices split across
    the ranks of `group` like `sweep`; the transformed volumes are never written (vt_project_strided_f32).
    Returns (projections, indices): `projections[i]` (d1, d2) belongs to `matrices[indices[i]]`, resident on this rank."""
    import torch.distributed as dist
    rank, world = _rank_world(dist, group)
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
    rank, world = _rank_world(dist, group)
    if engine is None:
        import torch
        engine = CudaEngine(torch.cuda.current_device())
    buffer, width = _prepare_and_broadcast(engine, dist, group, src, volume, interpolation, shape)
    z0, z1 = split_slabs(int(buffer.shape[0]), world, rank)
    m = np.ascontiguousarray(matrix, dtype=np.float32).reshape(1, 4, 4)
    part = engine.project_many(buffer, width, interpolation, m, z_range=(z0, z1))[0].contiguous()
    if world > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part


def gather_slabs(slab, group=None, dst: int = 0):
    """Convenience for tests / small volumes: concatenates the z-slabs on rank `dst` (None elsewhere)."""
    import torch
    import torch.distributed as dist
    rank, world = _rank_world(dist, group)
    if world == 1:
        return slab
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
