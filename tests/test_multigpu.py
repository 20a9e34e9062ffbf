"""CPU suite for the multi-GPU layer: world_size-2 gloo processes, the oracle standing in for the CUDA engine.

What is tested is the host-side orchestration of voltools_b200/multigpu.py (partitioning of matrices and z-slabs,
metadata + buffer broadcast from the root, result placement); the kernels themselves are covered by the -m gpu suite.
"""
import os
import tempfile

import numpy as np
import pytest

import oracle
from voltools_b200 import multigpu
from voltools_b200.utils import transform_matrix


def test_stream_plan_covers_the_volume_in_order():
    for d0 in (1, 13, 64, 100, 250, 512, 1024):
        for filtered in (False, True):
            for chunks in (None, 1, 3, 8):
                plan = multigpu.stream_plan(d0, filtered, chunks)
                assert plan[0][0] == 0 and plan[-1][1] == d0 and plan[0][2] == 0 and plan[-1][3] == d0
                for (a0, a1, z0, z1), (b0, b1, y0, y1) in zip(plan, plan[1:]):
                    assert a1 == b0 and z1 == y0 and z0 <= z1
                for xy0, xy1, z0, z1 in plan:
                    # the Z pass may only touch planes whose XY passes are done, 12 planes of look-ahead included
                    assert z1 <= xy1 and (not filtered or z1 == d0 or z1 <= xy1 - multigpu.PREFILTER_LOOKAHEAD or z1 == z0)


def test_streaming_margin():
    c = np.array([10, 10, 10], np.float32)
    rot = [transform_matrix(rotation=(0, a, 0), center=c) for a in (0, 30, 45)]
    assert multigpu.streaming_margin(rot, 'linear') == 1 and multigpu.streaming_margin(rot, 'filt_bspline') == 2
    shifted = rot + [transform_matrix(rotation=(0, 10, 0), center=c, translation=(-3, 0.5, 0))]   # t0 = +3
    assert multigpu.streaming_margin(shifted, 'bspline') == 5
    assert multigpu.streaming_margin([transform_matrix(rotation=(10, 20, 30), center=c)], 'linear') is None
    assert multigpu.streaming_margin([transform_matrix(translation=(0.5, 0, 0))], 'linear') is None


def test_partitions_cover_everything_once():
    for n in (0, 1, 2, 7, 180, 181, 1024):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                blk = multigpu.split_batch(n, world, r)
                seen += list(blk)
                assert len(blk) in (n // world, n // world + 1)
            assert seen == list(range(n))
            strided = sorted(i for r in range(world) for i in multigpu.split_strided(n, world, r))
            assert strided == list(range(n))
            z = [multigpu.split_slabs(n, world, r) for r in range(world)]
            assert z[0][0] == 0 and z[-1][1] == n and all(z[i][1] == z[i + 1][0] for i in range(world - 1))


def test_slab_footprints_cover_every_tap():
    """Every texel a slab's kernels can multiply into a result lies in a box / block its rank receives (or owns)."""
    rng = np.random.default_rng(5)
    for shape, world in (((32, 32, 64), 2), ((64, 32, 32), 4), ((24, 20, 28), 3)):
        c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
        for trial in range(6):
            m = transform_matrix(rotation=tuple(rng.uniform(-40, 40, 3)), rotation_order='sxyz', center=c,
                                 scale=tuple(rng.uniform(0.8, 1.25, 3)), translation=tuple(rng.uniform(-4, 4, 3)))
            boxes = multigpu.slab_footprint_boxes(m, shape, world)
            block = multigpu.footprint_block_size(shape, world)
            blocks = multigpu.slab_footprint_blocks(m, shape, shape[2], world, block) if block else None
            assert (block is None) == (shape == (24, 20, 28))
            a = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing='ij'), -1).reshape(-1, 3).astype(np.float32)
            p = a @ m[:3, :3].T.astype(np.float32) + m[:3, 3]  # index-space sample points (float32 like the kernels)
            ok = np.all((p + 0.5 >= 0) & (p + 0.5 < np.array(shape)), axis=1)
            base = np.floor(p).astype(int)
            for q in range(world):
                z0, z1 = multigpu.split_slabs(shape[0], world, q)
                sel = ok & (a[:, 0] >= z0) & (a[:, 0] < z1)
                have = np.zeros(shape, bool)
                for (r, qq), bx in boxes.items():
                    if qq == q:
                        lo, hi = bx
                        have[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
                have_b = np.zeros(shape, bool)
                if blocks is not None:
                    for (r, qq), idx in blocks.items():
                        if qq == q:
                            for iz, iy, ix in idx:
                                have_b[iz * block[0]:(iz + 1) * block[0], iy * block[1]:(iy + 1) * block[1],
                                       ix * block[2]:(ix + 1) * block[2]] = True
                for dz in (-1, 0, 1, 2):
                    for dy in (-1, 0, 1, 2):
                        for dx in (-1, 0, 1, 2):
                            t = base[sel] + np.array([dz, dy, dx])
                            inside = np.all((t >= 0) & (t < np.array(shape)), axis=1)
                            t = t[inside]
                            assert have[t[:, 0], t[:, 1], t[:, 2]].all(), (shape, world, trial, q, 'boxes')
                            if blocks is not None:
                                assert have_b[t[:, 0], t[:, 1], t[:, 2]].all(), (shape, world, trial, q, 'blocks')
            # an owner only sends what it owns
            for (r, q), (lo, hi) in boxes.items():
                i0, i1 = multigpu.split_slabs(shape[0], world, r)
                assert i0 <= lo[0] < hi[0] <= i1


class OracleEngine:
    """CPU stand-in for CudaEngine (test infrastructure): same interface, oracle arithmetic, torch CPU tensors."""

    def prepare(self, volume, interpolation):
        import torch
        v = oracle.prefilter(volume) if interpolation.startswith('filt') else np.asarray(volume, np.float32)
        return torch.from_numpy(np.ascontiguousarray(v)), v.shape[2]

    def describe(self, volume, interpolation):
        return tuple(int(v) for v in volume.shape), int(volume.shape[2])

    def describe_shape(self, shape):
        return tuple(shape), shape[2]

    def producer_stream(self):
        import contextlib
        return contextlib.nullcontext()

    def resample_many_range(self, buffer, width, interpolation, matrices, out, z0, z1):
        import torch
        v = buffer.numpy()[:, :, :width]
        for k, m in enumerate(matrices):
            full = oracle.affine(v, m, self._mode(interpolation), z_range=(z0, z1))
            out[k, z0:z1] = torch.from_numpy(full[z0:z1].copy())

    def prepare_stream(self, volume, interpolation, buffer, plan):
        coef, _ = self.prepare(volume, interpolation)

        def step(i):
            _, _, z0, z1 = plan[i]
            buffer[z0:z1].copy_(coef[z0:z1])
        return step

    def empty(self, shape):
        import torch
        return torch.empty(shape, dtype=torch.float32)

    def to_device(self, volume):
        import torch
        return torch.from_numpy(np.ascontiguousarray(volume, np.float32))

    def prefilter_slab(self, raw_slab, buffer, shape, interpolation, xy0, xy1, z0, z1):
        import torch
        v = raw_slab.numpy()
        # the sub-volume's ends stand in for the true ends 12 planes away: |pole|^12 = 1.4e-7 of the range
        coef = oracle.prefilter(v) if interpolation.startswith('filt') else v
        buffer[z0:z1].copy_(torch.from_numpy(np.ascontiguousarray(coef[z0 - xy0:z1 - xy0])))

    @staticmethod
    def _mode(interpolation):  # the buffer already holds coefficients
        return interpolation.replace('filt_', '')

    def resample_many(self, buffer, width, interpolation, matrices):
        import torch
        v = buffer.numpy()[:, :, :width]
        return torch.from_numpy(np.stack([oracle.affine(v, m, self._mode(interpolation)) for m in matrices]))

    def project_many(self, buffer, width, interpolation, matrices, z_range=None):
        import torch
        v = buffer.numpy()[:, :, :width]
        z0, z1 = (0, v.shape[0]) if z_range is None else z_range
        return torch.from_numpy(np.stack([
            oracle.affine(v, m, self._mode(interpolation), z_range=(z0, z1))[z0:z1].astype(np.float64).sum(axis=0)
            .astype(np.float32) for m in matrices]))

    def resample_slab(self, buffer, width, interpolation, matrix, z0, z1):
        import torch
        v = buffer.numpy()[:, :, :width]
        full = oracle.affine(v, matrix, self._mode(interpolation), z_range=(z0, z1))
        return torch.from_numpy(full[z0:z1].copy())


def _worker(rank, world, init_file, outdir):
    import torch
    import torch.distributed as dist
    dist.init_process_group('gloo', init_method=f'file://{init_file}', rank=rank, world_size=world)
    try:
        shape = (13, 16, 18)
        rng = np.random.default_rng(42)
        vol = rng.random(shape, dtype=np.float32)
        c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
        mats = [transform_matrix(rotation=(0, a, 0), center=c) for a in range(0, 180, 36)]
        eng = OracleEngine()
        # only the root has the samples
        out, idx = multigpu.sweep(vol if rank == 0 else None, mats, 'filt_bspline', src=0, engine=eng)
        np.savez(os.path.join(outdir, f'sweep_{rank}.npz'), out=out.numpy(), idx=np.array(idx))
        # a taller volume in 3 z-chunks: resampling overlapped with the chunked broadcast (shape known everywhere)
        tall = np.random.default_rng(43).random((40, 12, 14), dtype=np.float32)
        ct = np.divide(np.subtract(tall.shape, 1), 2, dtype=np.float32)
        mats_t = [transform_matrix(rotation=(0, a, 0), center=ct, translation=(t, 0.5, 0)) for a, t in
                  ((0, 0), (30, 2), (77, -3), (120, 0), (45, 1))]
        out_t, idx_t = multigpu.sweep(tall if rank == 0 else None, mats_t, 'filt_bspline', src=0, engine=eng,
                                      shape=tall.shape, chunks=3, overlap=True)
        np.savez(os.path.join(outdir, f'tall_{rank}.npz'), out=out_t.numpy(), idx=np.array(idx_t))
        m = transform_matrix(rotation=(20, 30, 40), translation=(1, -2, 0.5), center=c)
        slab, (z0, z1) = multigpu.zslab_affine(vol if rank == 0 else None, m, 'bspline_simple', src=0, engine=eng)
        full = multigpu.gather_slabs(slab, dst=0)
        np.savez(os.path.join(outdir, f'slab_{rank}.npz'), slab=slab.numpy(), z=np.array([z0, z1]),
                 full=full.numpy() if full is not None else np.zeros(0))
        # footprint path: raw z-slabs scattered, per-rank prefilter, block-sparse / boxed exchange of what each slab reads
        big = np.random.default_rng(44).random((64, 32, 48), dtype=np.float32)
        cb = np.divide(np.subtract(big.shape, 1), 2, dtype=np.float32)
        for tag, mm, shp, v in (('blocks', transform_matrix(rotation=(5, 8, -6), rotation_order='sxyz', center=cb,
                                                            translation=(1.5, -2, 1)), big.shape, big),
                                ('boxes', m, shape, vol)):
            info = {}
            slab_f, zr = multigpu.zslab_affine(v if rank == 0 else None, mm, 'filt_bspline', src=0, engine=eng, shape=shp,
                                               footprint=True, timings=info)
            np.savez(os.path.join(outdir, f'fp_{tag}_{rank}.npz'), slab=slab_f.numpy(), z=np.array(zr),
                     path=np.array([info['info']['path']]), block=np.array(info['info']['footprint_block'] or (0, 0, 0)))
        # rotate-and-project: a tilt series split across the ranks, and one projection summed from z-slabs (all-reduce)
        tilts = [transform_matrix(rotation=(a, 0, 0), rotation_order='sxyz', center=c) for a in (-60, -20, 15, 50, 72)]
        proj, pidx = multigpu.project_sweep(vol if rank == 0 else None, tilts, 'filt_bspline', src=0, engine=eng)
        whole = multigpu.zslab_project(vol if rank == 0 else None, m, 'bspline_simple', src=0, engine=eng)
        np.savez(os.path.join(outdir, f'proj_{rank}.npz'), proj=proj.numpy(), idx=np.array(pidx), whole=whole.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sweep_and_zslab_two_ranks_gloo():
    import torch.multiprocessing as mp
    world = 2
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, 'rendezvous')
        mp.spawn(_worker, args=(world, init_file, d), nprocs=world, join=True)
        shape = (13, 16, 18)
        vol = np.random.default_rng(42).random(shape, dtype=np.float32)
        c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
        mats = [transform_matrix(rotation=(0, a, 0), center=c) for a in range(0, 180, 36)]
        seen = []
        for r in range(world):
            z = np.load(os.path.join(d, f'sweep_{r}.npz'))
            for o, i in zip(z['out'], z['idx']):
                assert np.array_equal(o, oracle.affine(vol, mats[int(i)], 'filt_bspline'))
                seen.append(int(i))
        assert sorted(seen) == list(range(len(mats)))
        tall = np.random.default_rng(43).random((40, 12, 14), dtype=np.float32)
        ct = np.divide(np.subtract(tall.shape, 1), 2, dtype=np.float32)
        mats_t = [transform_matrix(rotation=(0, a, 0), center=ct, translation=(t, 0.5, 0)) for a, t in
                  ((0, 0), (30, 2), (77, -3), (120, 0), (45, 1))]
        seen = []
        for r in range(world):
            z = np.load(os.path.join(d, f'tall_{r}.npz'))
            for o, i in zip(z['out'], z['idx']):
                assert np.array_equal(o, oracle.affine(tall, mats_t[int(i)], 'filt_bspline')), (r, int(i))
                seen.append(int(i))
        assert sorted(seen) == list(range(len(mats_t)))
        m = transform_matrix(rotation=(20, 30, 40), translation=(1, -2, 0.5), center=c)
        want = oracle.affine(vol, m, 'bspline_simple')
        z0 = np.load(os.path.join(d, 'slab_0.npz'))
        z1 = np.load(os.path.join(d, 'slab_1.npz'))
        assert tuple(z0['z']) == (0, 7) and tuple(z1['z']) == (7, 13)
        assert np.array_equal(z0['slab'], want[0:7]) and np.array_equal(z1['slab'], want[7:13])
        assert np.array_equal(z0['full'], want)
        # footprint path: every rank only ever held its slab's input footprint; results agree with the one-GPU answer to
        # the windowed prefilter's 1e-6 of the coefficient range
        big = np.random.default_rng(44).random((64, 32, 48), dtype=np.float32)
        cb = np.divide(np.subtract(big.shape, 1), 2, dtype=np.float32)
        mb = transform_matrix(rotation=(5, 8, -6), rotation_order='sxyz', center=cb, translation=(1.5, -2, 1))
        for tag, mm, v in (('blocks', mb, big), ('boxes', m, vol)):
            ref = oracle.affine(v, mm, 'filt_bspline')
            tol = 2e-6 * float(np.ptp(oracle.prefilter(v)))
            for r in range(world):
                z = np.load(os.path.join(d, f'fp_{tag}_{r}.npz'))
                assert str(z['path'][0]) == 'footprint'
                assert (tuple(z['block']) != (0, 0, 0)) == (tag == 'blocks')
                a, b = (int(t) for t in z['z'])
                assert np.abs(z['slab'] - ref[a:b]).max() <= tol, (tag, r)
        tilts = [transform_matrix(rotation=(a, 0, 0), rotation_order='sxyz', center=c) for a in (-60, -20, 15, 50, 72)]
        seen = []
        for r in range(world):
            z = np.load(os.path.join(d, f'proj_{r}.npz'))
            for p, i in zip(z['proj'], z['idx']):
                ref = oracle.affine(vol, tilts[int(i)], 'filt_bspline').astype(np.float64).sum(axis=0)
                assert np.allclose(p, ref, rtol=0, atol=1e-5 * shape[0])
                seen.append(int(i))
            # the all-reduced projection is complete and identical on every rank
            assert np.allclose(z['whole'], want.astype(np.float64).sum(axis=0), rtol=0, atol=1e-5 * shape[0])
        assert sorted(seen) == list(range(len(tilts)))


def test_world_of_one_needs_no_process_group():
    """Without torch.distributed initialised every entry point degenerates to its single-GPU form (no collective)."""
    import torch.distributed as dist
    assert not dist.is_initialized()
    shape = (13, 16, 18)
    vol = np.random.default_rng(42).random(shape, dtype=np.float32)
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    eng = OracleEngine()
    mats = [transform_matrix(rotation=(0, a, 0), center=c) for a in (0, 40, 95)]
    out, idx = multigpu.sweep(vol, mats, 'filt_bspline', engine=eng)
    assert idx == [0, 1, 2]
    for o, m in zip(out.numpy(), mats):
        assert np.array_equal(o, oracle.affine(vol, m, 'filt_bspline'))
    m = transform_matrix(rotation=(20, 30, 40), translation=(1, -2, 0.5), center=c)
    info = {}
    slab, (z0, z1) = multigpu.zslab_affine(vol, m, 'bspline_simple', engine=eng, shape=shape, timings=info)
    assert (z0, z1) == (0, 13) and info['info']['path'] == 'broadcast'
    assert np.array_equal(slab.numpy(), oracle.affine(vol, m, 'bspline_simple'))
    assert multigpu.gather_slabs(slab) is slab
    buf, width = multigpu.prepare_and_broadcast(eng, vol, 'filt_bspline', shape=shape)
    assert width == 18 and np.array_equal(buf.numpy(), oracle.prefilter(vol))
    whole = multigpu.zslab_project(vol, m, 'bspline_simple', engine=eng)
    assert np.allclose(whole.numpy(), oracle.affine(vol, m, 'bspline_simple').astype(np.float64).sum(axis=0), atol=1e-4)
