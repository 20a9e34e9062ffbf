"""Multi-GPU correctness on real GPUs (run under torchrun): sharded sweep / z-slab results == single-GPU results."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import multigpu  # noqa: E402

rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
shape = (140, 70, 90)  # deep enough for the pipelined prepare + broadcast to use several z-chunks
vol = np.random.default_rng(3).random(shape, dtype=np.float32)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
mats = [vt.utils.transform_matrix(rotation=(0, a, 0), center=c) for a in range(0, 180, 12)]
for mode in ('filt_bspline', 'linear', 'bspline_simple'):
    outs, idx = multigpu.sweep(vol if rank == 0 else None, mats, mode, overlap=(mode != 'linear'))
    sv = vt.StaticVolume(vol, interpolation=mode, device=f'gpu:{local}')
    ref = sv.affine_many([mats[i] for i in idx])
    # streamed prefilter (z-chunks) vs one-shot prefilter: same coefficients to ~1e-7 of their range
    tol = 1e-6 * float(sv.coefficients.max() - sv.coefficients.min()) if mode.startswith('filt') else 0.0
    assert float((outs - ref).abs().max()) <= tol, (mode, rank, float((outs - ref).abs().max()))
    m = vt.utils.transform_matrix(rotation=(20, 30, 40), translation=(1, -2, 0.5), center=c)
    slab, (z0, z1) = multigpu.zslab_affine(vol if rank == 0 else None, m, mode)
    full = sv.affine_many([m])[0]
    assert float((slab - full[z0:z1]).abs().max()) <= tol, (mode, rank, z0, z1)
    whole = multigpu.gather_slabs(slab)
    if rank == 0:
        assert float((whole - full).abs().max()) <= tol
    # rotate-and-project: a tilt series split across the ranks; one projection summed from z-slabs by all-reduce
    tilts = [vt.utils.transform_matrix(rotation=(a, 0, 0), rotation_order='sxyz', center=c) for a in range(-60, 60, 15)]
    proj, pidx = multigpu.project_sweep(vol if rank == 0 else None, tilts, mode)
    scale = (float(sv.coefficients.max() - sv.coefficients.min())) * shape[0]
    ref_p = sv.project_many([tilts[i] for i in pidx])
    assert float((proj - ref_p).abs().max()) <= 2e-6 * scale, (mode, rank, 'project_sweep')
    whole_p = multigpu.zslab_project(vol if rank == 0 else None, m, mode)
    assert float((whole_p - full.double().sum(dim=0).float()).abs().max()) <= 2e-6 * scale, (mode, rank, 'zslab_project')
# z-slab sharding with per-slab input footprints: block-sparse exchange (a shape the blocks tile) and boxed exchange
for fshape, mkw in (((64 * world, 64, 128), dict(rotation=(5, 8, -6), rotation_order='sxyz', translation=(1.5, -2, 1))),
                    ((64 * world, 128, 64), dict(rotation=(30, 45, 60), scale=(1.1, 0.9, 1.05), translation=(2.5, -1, 0))),
                    ((140, 70, 90), dict(rotation=(20, 30, 40), translation=(1, -2, 0.5)))):
    fvol = np.random.default_rng(5).random(fshape, dtype=np.float32)
    fc = np.divide(np.subtract(fshape, 1), 2, dtype=np.float32)
    fm = vt.utils.transform_matrix(center=fc, **mkw)
    for mode in ('filt_bspline', 'linear'):
        info = {}
        slab, (z0, z1) = multigpu.zslab_affine(fvol if rank == 0 else None, fm, mode, shape=fshape, footprint=True, timings=info)
        sv = vt.StaticVolume(fvol, interpolation=mode, device=f'gpu:{local}')
        full = sv.affine_many([fm])[0]
        tol = 2e-6 * float(sv.coefficients.max() - sv.coefficients.min()) if mode.startswith('filt') else 0.0
        err = float((slab - full[z0:z1]).abs().max())
        assert info['info']['path'] == 'footprint', info
        assert err <= tol, (fshape, mode, rank, err, tol, info['info'])
        torch.cuda.synchronize()
        if rank == 0:
            print('footprint path', fshape, mode, info['info'], 'distribute %.3f ms resample %.3f ms' %
                  (info['distribute_ms'](), info['resample_ms']()))
dist.barrier()
if rank == 0:
    print(f'multi-GPU check OK on {world} ranks')
dist.destroy_process_group()
