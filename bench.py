#!/usr/bin/env python
"""bench.py -- Gvoxels/s of the affine volume-resampling hot path on N B200s (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Default run = the headline workload plus the secondary tables, all in the ONE JSON line:
  headline  BASELINE configs[2] (README.md:26-27 of the reference): StaticVolume 256^3 float32 'filt_bspline',
            180-angle sweep rotation=(0,i,0) 'rzxz' about the centre.  One step = the whole job: prefilter on rank 0,
            [N > 1: one NCCL broadcast of the coefficient volume,] the 180 matrices dealt round-robin to the ranks, every
            rank resamples its share through StaticVolume.affine_many (public API, eager calls).  Strong scaling: the
            180 transforms are the fixed job; `value` = 180 * 256^3 voxels / step time (CUDA events, max over ranks).
  modes_512   every interpolation mode at 512^3 through affine(output=): rot45 (configs[1]'s matrix), the full affine
              G6 (configs[3]), and the mean over random 3-axis rotations (the reference's tests/benchmark.py:52-54 set);
              ms, Gvox/s, fraction of the 8 (16 for filt_*) B/voxel roofline, in-bounds fraction.  N = 1 only.
  zslab_1024  configs[4]: 1024^3 filt_bspline, z-slab sharding, distribution and resampling timed separately
              (a mild rotation and the full affine G6).
  cfg1        configs[1]: transform() of 250^3 'filt_bspline' rot45, eager calls, batch of 8 volumes.  N = 1 only.
Single pieces: --workload sweep|modes|zslab|cfg1|project|cpu_baselines.

`e2e` is the headline job through the same public API with HOST buffers: the volume starts in pinned host memory
(H2D inside the timed region) and every output volume is returned to the host (`sv.transform(...)` -> numpy).
`roofline` is for the kernel with the largest share of the headline step (per-launch CUDA events recorded by the
library, vt_profile_*), against MEASURED_PEAKS.json.  `cpu_baseline` / `--impl reference` time the unmodified reference
package's CPU path (voltools.transform(device='cpu') -> scipy.ndimage.affine_transform), staged under oracle/_ref/py,
on this host's cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = 'Gvoxels/s'
ROT45 = dict(rotation=(0, 45, 0), rotation_order='rzxz')
FULL_AFFINE = dict(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60), rotation_order='rzxz',
                   translation=(5.5, -3.25, 2.0))
MILD = dict(rotation=(5, 8, -6), rotation_order='sxyz', translation=(3.5, -2.0, 1.0))
SWEEP = dict(n=256, angles=180, interpolation='filt_bspline',
             name="configs[2]: StaticVolume 256^3 float32 'filt_bspline', 180-angle sweep rotation=(0,i,0) rzxz "
                  '(prefilter on rank 0 + one NCCL broadcast + angles dealt round-robin; step = the whole sweep)')
CFG1 = dict(n=250, batch=8, interpolation='filt_bspline', kw=ROT45)


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def inbounds_fraction(shape, m):
    """Fraction of output voxels whose sample point lies inside the source (transforms.py:276-278); float64 on a
    strided sub-grid, informational."""
    step = max(1, min(shape) // 64)
    idx = [np.arange(0, s, step, dtype=np.float64) for s in shape]
    a = np.stack(np.meshgrid(*idx, indexing='ij'), axis=-1).reshape(-1, 3)
    p = a @ np.asarray(m, np.float64)[:3, :3].T + np.asarray(m, np.float64)[:3, 3] + 0.5
    ok = np.all((p >= 0) & (p < np.asarray(shape, np.float64)), axis=1)
    return float(ok.mean())


# ------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, uuid):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', uuid, f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def rows(self):
        """Samples delivered so far."""
        if self.p is None:
            return 1 << 30  # no nvidia-smi here: nothing to wait for
        try:
            return sum(1 for r in Path(self.f.name).read_text().splitlines() if r.count(',') >= 6)
        except OSError:
            return 0

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(',') for r in Path(self.f.name).read_text().strip().splitlines() if r.count(',') >= 6]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.strip().lower().startswith('active'):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def clock_sampler(torch, dev):
    uuid = str(getattr(torch.cuda.get_device_properties(dev), 'uuid', dev))
    return ClockSampler(uuid if uuid.startswith('GPU-') or uuid.isdigit() else 'GPU-' + uuid)


# ------------------------------------------------------------------------------------------------------------
# reference CPU path (the unmodified reference package, staged by oracle/build_ref.py under oracle/_ref/py)
# ------------------------------------------------------------------------------------------------------------
def _ref_voltools():
    import contextlib
    import io
    import oracle
    path = oracle.ref_python_path()
    if path is None:
        return None
    sys.path.insert(0, path)
    try:
        with contextlib.redirect_stdout(io.StringIO()):  # the reference prints a cupy-missing warning on import
            import voltools as ref_vt
    finally:
        sys.path.remove(path)
    return ref_vt


def _ref_worker(job):
    seed, shape, interpolation, kw = job
    ref_vt = _ref_voltools()
    vol = np.random.default_rng(seed).random(shape, dtype=np.float32)
    t0 = time.perf_counter()
    out = ref_vt.transform(vol, interpolation=interpolation, device='cpu', **kw)
    dt = time.perf_counter() - t0
    return dt, float(out[shape[0] // 2, shape[1] // 2, shape[2] // 2])


def _sweep_kw(i):
    return dict(rotation=(0, i, 0), rotation_order='rzxz')


def cpu_baseline_leg(cfg):
    """Reference CPU path on a bounded sample of the headline job, one core (scipy.ndimage is single-threaded): a few
    of the 180 angles on the full 256^3 volume (~3.6 s each)."""
    ref_vt = _ref_voltools()
    if ref_vt is None:
        return {'value': None, 'unit': METRIC, 'cores': 0, 'kind': 'reference', 'sample': 'reference package not staged'}
    n = cfg['n']
    angles = (0, 45, 90, 135)
    dt = sum(_ref_worker((0, (n, n, n), cfg['interpolation'], _sweep_kw(a)))[0] for a in angles)
    out = {'value': len(angles) * n ** 3 / dt / 1e9, 'unit': METRIC, 'cores': 1, 'kind': 'reference',
           'sample': f"{len(angles)} of the sweep's {cfg['angles']} angles {angles} on the full {n}^3 volume through the "
                     f"unmodified reference voltools.transform(device='cpu') = scipy.ndimage.affine_transform(order=3, "
                     f'prefilter=True); {dt:.2f} s on 1 of {os.cpu_count()} host cores (the reference has no resident '
                     'CPU volume: every call prefilters again)'}
    # scipy.ndimage.affine_transform directly (tests/benchmark.py:60), order 1 and 3, 100^3 (BASELINE configs[0] shape)
    try:
        from scipy import ndimage
        from voltools_b200.utils import transform_matrix
        v = np.random.default_rng(0).random((100, 100, 100), dtype=np.float32)
        m = transform_matrix(center=(49.5, 49.5, 49.5), **ROT45)
        extra = {}
        for order in (1, 3):
            t0 = time.perf_counter()
            ndimage.affine_transform(v, m, order=order)
            extra[f'scipy_affine_transform_order{order}_100^3_gvox_s'] = 1e6 / (time.perf_counter() - t0) / 1e9
        t0 = time.perf_counter()
        ref_vt.transform(v, interpolation='linear', device='cpu', **ROT45)
        extra['reference_cpu_linear_100^3_gvox_s (BASELINE configs[0])'] = 1e6 / (time.perf_counter() - t0) / 1e9
        out['extra'] = extra
    except Exception as e:  # informational only
        out['extra'] = {'error': repr(e)}
    return out


def run_cpu_baselines(args):
    """SURVEY 8(d) CPU baselines on this host: the unmodified reference's CPU path (voltools.transform(device='cpu'))
    and scipy.ndimage.affine_transform (order 1 / 3, tests/benchmark.py:60) at 100^3, 250^3, 256^3 and 512^3, rot45
    matrix, one core each (SciPy ndimage is single-threaded).  1024^3 is extrapolated (x8 voxels of the 512^3 time)."""
    from scipy import ndimage
    from voltools_b200.utils import transform_matrix
    ref_vt = _ref_voltools()
    out = {'cores_used': 1, 'host_cores': os.cpu_count(), 'rows': {}}
    for n in (100, 250, 256, 512):
        v = np.random.default_rng(0).random((n, n, n), dtype=np.float32)
        c = np.divide(np.subtract(v.shape, 1), 2, dtype=np.float32)
        m = transform_matrix(center=c, **ROT45)
        row = {}
        for label, fn in (('scipy_order1', lambda: ndimage.affine_transform(v, m, order=1)),
                          ('scipy_order3', lambda: ndimage.affine_transform(v, m, order=3)),
                          ('reference_cpu_linear', (lambda: ref_vt.transform(v, interpolation='linear', device='cpu', **ROT45))
                           if ref_vt else None),
                          ('reference_cpu_filt_bspline',
                           (lambda: ref_vt.transform(v, interpolation='filt_bspline', device='cpu', **ROT45))
                           if ref_vt else None)):
            if fn is None:
                continue
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            row[label] = {'s': dt, 'gvox_s': n ** 3 / dt / 1e9}
        out['rows'][f'{n}^3'] = row
    out['rows']['1024^3 (extrapolated: 8 x the 512^3 time)'] = {
        k: {'s': 8 * r['s'], 'gvox_s': r['gvox_s']} for k, r in out['rows']['512^3'].items()}
    return out


def reference_arm(args, cfg):
    """--impl reference: the reference's own CPU implementation of the headline job, all the host parallelism it can
    use (one process per angle; each call is single-threaded inside SciPy).  Bounded sample per step: every process
    transforms a z-slab of the 256^3 volume for its own angle (rotation about axis 0: every output plane costs the same)."""
    import multiprocessing as mp
    rank, _, world = dist_env()
    if rank != 0:
        return
    if _ref_voltools() is None:
        print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/_ref/py (reference package) not staged'}))
        return
    n = cfg['n']
    procs = max(1, min(os.cpu_count() or 1, 32))
    t_probe, _ = _ref_worker((0, (16, n, n), cfg['interpolation'], _sweep_kw(45)))
    per_plane = t_probe / 16
    budget = float(os.environ.get('VT_BENCH_REF_BUDGET_S', 150.0)) / (args.steps + args.warmup)  # CPU seconds per step
    planes = int(max(8, min(n, budget / per_plane)))
    shape = (planes, n, n)
    jobs = [(1000, shape, cfg['interpolation'], _sweep_kw((i * cfg['angles']) // procs)) for i in range(procs)]
    with mp.get_context('fork').Pool(procs) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, jobs)
        dt = time.perf_counter() - t0
    vox = procs * planes * n * n * args.steps
    value = vox / dt / 1e9
    sample = (f'{procs} processes (of {os.cpu_count()} host cores) x one {planes}x{n}x{n} z-slab of the {n}^3 volume per '
              f"step, each at its own angle of the sweep, unmodified reference voltools.transform(interpolation="
              f"'{cfg['interpolation']}', device='cpu')")
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': METRIC, 'cores': procs, 'kind': 'reference', 'sample': sample},
        'e2e': {'value': value, 'unit': METRIC, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def timed(torch, fn, steps, warmup, barrier):
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / 1e3


def med_ms(torch, fn, steps, warmup, flush=None):
    ts = []
    for it in range(warmup + steps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if it >= warmup:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def run_sweep(args, torch, vt, dev, barrier, reduce_max, with_e2e=True):
    """The headline: BASELINE configs[2]."""
    from voltools_b200 import _native, multigpu
    rank, _, world = dist_env()
    cfg = SWEEP
    n = args.size if args.workload == 'sweep' and args.size != 512 else cfg['n']
    shape = (n, n, n)
    interp = cfg['interpolation']
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    mats = [vt.utils.transform_matrix(center=c, **_sweep_kw(i)) for i in range(cfg['angles'])]
    vol = torch.rand(shape, device=f'cuda:{dev}', generator=torch.Generator(f'cuda:{dev}').manual_seed(7)) \
        if rank == 0 else None
    mine = multigpu.split_strided(len(mats), world, rank)
    my_mats = np.stack([mats[i] for i in mine]) if len(mine) else np.zeros((0, 4, 4), np.float32)
    out = torch.empty((len(mine),) + shape, device=f'cuda:{dev}')
    eng = multigpu.CudaEngine(dev)

    def step():
        if world > 1:
            buf, width = multigpu.prepare_and_broadcast(eng, vol, interp, src=0, shape=shape)
            sv = vt.StaticVolume.from_coefficients(buf, interp, width)
        else:
            sv = vt.StaticVolume(vol, interpolation=interp, device=f'gpu:{dev}')
        if len(mine):
            sv.affine_many(my_mats, output=out, zero_fill=True)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    # nvidia-smi (100 ms period) needs samples under this load: ~0.6 s of steps before the timed region.  The step holds a
    # collective, so EVERY rank must run the same number of them: the count comes from a max-reduced time, never from a
    # rank's own clock.
    barrier()
    t_probe = time.perf_counter()
    step()
    torch.cuda.synchronize()
    step_s = reduce_max(time.perf_counter() - t_probe)
    n_load = int(min(400, max(2, 0.4 / max(step_s, 1e-4))))
    clocks = clock_sampler(torch, dev)
    for _ in range(8):  # chunks of ~0.4 s until the sampler has delivered a few rows (nvidia-smi can take a second to start)
        for _ in range(n_load):
            step()
        torch.cuda.synchronize()
        if reduce_max(0.0 if clocks.rows() >= 3 else 1.0) == 0.0:  # collective: every rank leaves the loop together
            break
    l0 = _native.launch_count()
    sec = reduce_max(timed(torch, step, args.steps, 0, barrier))
    launches = _native.launch_count() - l0
    clk = clocks.stop()
    vox_step = len(mats) * n ** 3
    value = vox_step * args.steps / sec / 1e9

    # per-kernel durations over K more steps (CUDA events inside the library, on the launch stream)
    _native.profile_enable(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    prof = _native.profile_read()
    _native.profile_enable(False)
    peak, peak_src = peaks()
    kernels = {}
    total_ms = sum(ms for ms, _ in prof.values()) or 1.0
    my_vox = len(mine) * n ** 3
    for name, (ms, cnt) in prof.items():
        # voxels this kernel processed on this rank over the K steps: the resampling kernels cover the rank's share of
        # the matrices, the prefilter / pack kernels one volume per step
        vox = my_vox * args.steps if ('cubic' in name or 'linear' in name) else n ** 3 * args.steps
        ach = 8.0 * vox / (ms * 1e-3) / 1e9
        kernels[name] = {'launches_per_step': cnt / args.steps, 'avg_ms': ms / cnt, 'share': ms / total_ms,
                         'voxels_per_launch': vox / cnt, 'achieved_gbs': ach, 'frac': ach / peak}
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    traffic = None
    tfile = ROOT / 'profiles' / 'traffic.json'
    if dom and tfile.exists():
        t = json.loads(tfile.read_text()).get(dom, {})
        # measured per voxel on a 32-matrix launch of the same kernel and shape, scaled to this run's average launch
        traffic = t['dram_bytes_per_voxel'] * kernels[dom]['voxels_per_launch'] if 'dram_bytes_per_voxel' in t \
            else t.get('dram_bytes_per_launch')
    roofline = None
    if dom:
        inb = float(np.mean([inbounds_fraction(shape, m) for m in mats[::6]]))
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kernels[dom]['achieved_gbs'], 'peak': peak, 'unit': 'GB/s',
                    'frac': kernels[dom]['frac'], 'traffic': traffic, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': 8.0 * kernels[dom]['voxels_per_launch'],
                    'inbounds_fraction': inb,
                    'step_frac': 8.0 * vox_step * args.steps / sec / 1e9 / peak / world,
                    'note': 'achieved = 8 B/voxel (4 B compulsory read + 4 B write) x voxels per launch / average launch '
                            'duration on rank 0; out-of-bounds voxels (1 - inbounds_fraction) are counted although they are '
                            'neither read nor written; ncu sees the algorithmic bytes at DRAM (the 64 MiB volume is fetched '
                            'about once per matrix, 4 B/voxel of writes); step_frac = the same for the whole step per GPU',
                    'kernels': kernels}

    line = {
        'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['name'].replace('256^3', f'{n}^3'), 'voxels_per_step': vox_step,
                   'parallelism': f'dp{world} over the matrix batch',
                   'l2': 'outputs (180 x 64 MiB per step) exceed L2 and stream through it; no flush between steps is needed',
                   'call': "StaticVolume(vol, 'filt_bspline').affine_many(matrices, output=device array) -- eager calls, "
                           'no CUDA graph', 'timing': 'CUDA events around K steps, barrier + synchronize on both sides, '
                                                      'max over ranks'},
        'clocks': {'sm_mhz': clk['sm_mhz'], 'sm_max_mhz': clk['sm_max_mhz'], 'reasons': clk['reasons'],
                   'samples': clk['samples']},
        'gpu_launches': int(launches),
        'roofline': roofline,
    }
    if with_e2e:
        line['e2e'] = sweep_e2e(args, torch, vt, dev, barrier, reduce_max, shape, interp, mats, mine, vol, out)
    return line


def sweep_e2e(args, torch, vt, dev, barrier, reduce_max, shape, interp, mats, mine, vol, out_dev):
    """The headline job end to end through the public API with HOST buffers: the volume starts in pinned host memory on
    rank 0, every output volume ends in host memory (`sv.affine(m)` returns numpy, like the reference's .get())."""
    from voltools_b200 import multigpu
    rank, _, world = dist_env()
    n = shape[0]
    h_vol = None
    if rank == 0:
        h_vol = torch.empty(shape, dtype=torch.float32).pin_memory()
        h_vol.copy_(vol)
        h_vol = h_vol.numpy()
    eng = multigpu.CudaEngine(dev)
    check = {}
    my_mats = np.stack([mats[i] for i in mine]) if len(mine) else np.zeros((0, 4, 4), np.float32)
    try:   # the rank's results land here, step after step (page-locked once, like the pinned input volume)
        h_out = vt.pinned_empty((len(mine),) + tuple(shape))
    except RuntimeError as e:
        h_out = None
        print(f'[bench] no page-locked result buffer ({e}); e2e falls back to per-angle calls', file=sys.stderr)

    def resident():
        if world > 1:
            buf, width = multigpu.prepare_and_broadcast(eng, h_vol, interp, src=0, shape=shape)
            return vt.StaticVolume.from_coefficients(buf, interp, width)
        return vt.StaticVolume(h_vol, interpolation=interp, device=f'gpu:{dev}')

    def step_batched():
        sv = resident()
        sv.affine_many(my_mats, output=h_out)   # returns when the last result is in host memory

    def step_per_call():
        sv = resident()
        for k, i in enumerate(mine):
            res = sv.affine(mats[i])  # numpy (pinned staging inside the library)
            if k == len(mine) // 2:
                check['res'], check['k'] = res, k

    def timed(step):
        steps = max(1, min(args.steps, 3))
        step()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        sec = reduce_max(time.perf_counter() - t0)
        barrier()
        return steps, sec

    vox = len(mats) * n ** 3
    raw = None
    if h_out is not None and len(mine):   # the link itself, right now: plain copies of one device volume into h_out
        d_one = torch.empty(shape, device=f'cuda:{dev}')
        h_t = torch.from_numpy(h_out)
        reps = min(len(mine), 16)
        for timed_pass in (False, True):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(reps):
                h_t[k].copy_(d_one, non_blocking=True)
            torch.cuda.synchronize()
            raw = reps * n ** 3 * 4 / (time.perf_counter() - t0) / 1e9
        del d_one, h_t
    steps, sec = timed(step_per_call)
    if 'res' in check:  # the host path must agree with the device path
        err = float((torch.from_numpy(check['res']).to(f'cuda:{dev}') - out_dev[check['k']]).abs().max())
        assert err <= 1e-5 * 16, f'host path differs from device path: {err}'
    per_call = {'value': vox * steps / sec / 1e9, 'ms_per_step': sec / steps * 1e3,
                'call': "StaticVolume(host volume, 'filt_bspline') then sv.affine(m) -> numpy for each of the rank's angles "
                        '(the reference README loop: one synchronous result at a time)'}
    phases, batched = None, None
    # (collective decision: a rank without its page-locked buffer must not leave the others alone in a barrier)
    if reduce_max(0.0 if h_out is not None else 1.0) == 0.0:
        steps_b, sec_b = timed(step_batched)
        for k in sorted({0, len(mine) // 2, len(mine) - 1} if len(mine) else ()):
            err = float((torch.from_numpy(h_out[k]).to(f'cuda:{dev}') - out_dev[k]).abs().max())
            assert err <= 1e-5 * 16, f'host results differ from device results: {err} (angle {mine[k]})'
        batched = {'value': vox * steps_b / sec_b / 1e9, 'ms_per_step': sec_b / steps_b * 1e3,
                   'call': "StaticVolume(host volume, 'filt_bspline').affine_many(matrices, output=page-locked numpy "
                           "array): the rank's results stream to host memory while the next kernels run"}
        # one more (untimed) step, phase by phase, on this rank
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sv = resident()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sv.affine_many(my_mats, output=h_out)
        t2 = time.perf_counter()
        phases = {'raw_d2h_GBps_rank0_before': raw, 'resident_volume_ms': (t1 - t0) * 1e3,
                  'affine_many_to_host_ms': (t2 - t1) * 1e3,
                  'd2h_GBps_rank0': len(mine) * n ** 3 * 4 / max(t2 - t1, 1e-9) / 1e9}
        barrier()
    # both forms are public-API calls for the same job with host buffers on both ends; the line's e2e is the faster one
    # (one GPU: the batched form, by the overlap; several GPUs sharing the host's PCIe root: whichever the link favours)
    best = batched if batched is not None and batched['value'] >= per_call['value'] else per_call
    return {'value': best['value'], 'unit': METRIC, 'h2d_bytes_per_step': n ** 3 * 4,
            'd2h_bytes_per_step': vox * 4, 'steps': steps, 'ms_per_step': best['ms_per_step'],
            'call': best['call'] + ' (H2D of the volume, prefilter, broadcast and D2H of every output volume inside the '
                                   'timed region; wall clock, max over ranks)',
            'per_call': per_call, 'batched': batched, 'phases': phases,
            'bound': 'PCIe D2H: 180 x 64 MiB per step (raw pinned copies on this host: 57 GB/s for one GPU, 127 GB/s '
                     'for four at once -- tools/e2e_probe.py)'}


def random_rotation_mats(vt, n, count):
    """The reference benchmark's matrices (tests/benchmark.py:52-54): 100 random 3-axis rotations, order 'sxyz', about
    centre = size/2 -- here the first `count` of a seeded set."""
    rots = np.random.default_rng(1).uniform(-180, 180, (100, 3))[:count]
    c = (n / 2, n / 2, n / 2)
    return [vt.utils.transform_matrix(rotation=tuple(r), rotation_units='deg', rotation_order='sxyz', center=c) for r in rots]


def run_modes(args, torch, vt, dev, n=None, steps=None, warmup=None, random_count=8):
    """All five modes at n^3 through affine(output=): kernel-level Gvox/s for rot45, the full affine and the reference
    benchmark's random rotations; L2 flushed between iterations, median."""
    n = n or args.size
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    shape = (n, n, n)
    peak, _ = peaks()
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    mats = {'rot45': [vt.utils.transform_matrix(center=c, **ROT45)],
            'full_affine': [vt.utils.transform_matrix(center=c, **FULL_AFFINE)],
            'random_rotations': random_rotation_mats(vt, n, random_count)}
    src = torch.rand(shape, device=f'cuda:{dev}')
    dst = torch.zeros(shape, device=f'cuda:{dev}')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{dev}')
    res = {}
    for mode in vt.AVAILABLE_INTERPOLATIONS:
        for mname, ms_list in mats.items():
            per = []
            for m in ms_list:
                k = steps if len(ms_list) == 1 else max(2, steps // 3)
                per.append(med_ms(torch, lambda: vt.affine(src, m, interpolation=mode, output=dst, device=f'gpu:{dev}'),
                                  k, warmup if len(ms_list) == 1 else 1, flush))
            ms = float(np.mean(per))
            bytes_per_vox = 16 if mode.startswith('filt') else 8
            res[f'{mode}/{mname}'] = {'ms': ms, 'gvox_s': n ** 3 / ms / 1e6,
                                      'roofline_frac': bytes_per_vox * n ** 3 / (ms * 1e-3) / 1e9 / peak,
                                      'inbounds_fraction': float(np.mean([inbounds_fraction(shape, m) for m in ms_list])),
                                      'matrices': len(ms_list)}
    # resident volumes (StaticVolume: prefilter / layouts paid once): the same three matrix classes
    for mode in vt.AVAILABLE_INTERPOLATIONS:
        sv = vt.StaticVolume(src, interpolation=mode, device=f'gpu:{dev}')
        for mname, ms_list in mats.items():
            per = [med_ms(torch, lambda: sv.affine(m, output=dst), max(2, steps // 2), 2, flush) for m in ms_list]
            ms = float(np.mean(per))
            res[f'static/{mode}/{mname}'] = {'ms': ms, 'gvox_s': n ** 3 / ms / 1e6,
                                             'roofline_frac': 8 * n ** 3 / (ms * 1e-3) / 1e9 / peak}
        del sv
    return res


def run_cfg1(args, torch, vt, dev):
    """configs[1]: transform() of 250^3 'filt_bspline' rot45, a batch of 8 independent device-resident volumes per step
    through the public API -- eager calls on one stream (`value`), and the same calls on two streams replayed from a CUDA
    graph (labelled)."""
    cfg = CFG1
    n, batch, interp, kw = cfg['n'], cfg['batch'], cfg['interpolation'], cfg['kw']
    shape = (n, n, n)
    device = f'gpu:{dev}'
    gen = torch.Generator(device=f'cuda:{dev}').manual_seed(1234)
    vols = [torch.rand(shape, generator=gen, device=f'cuda:{dev}', dtype=torch.float32) for _ in range(batch)]
    outs = [torch.zeros(shape, device=f'cuda:{dev}', dtype=torch.float32) for _ in range(batch)]

    def step():
        for v, o in zip(vols, outs):
            vt.transform(v, interpolation=interp, output=o, device=device, **kw)

    sec = timed(torch, step, args.steps, 3, lambda: None)
    res = {'workload': "configs[1]: transform() 250^3 float32 'filt_bspline' rotation=(0,45,0) rzxz, batch of 8 volumes "
                       '(inputs larger than L2), eager calls on one stream',
           'value': batch * n ** 3 * args.steps / sec / 1e9, 'unit': METRIC, 'ms_per_volume': sec / args.steps / batch * 1e3,
           'roofline_frac_16B_per_voxel': 16.0 * batch * n ** 3 * args.steps / sec / 1e9 / peaks()[0]}
    # labelled extra: two streams + CUDA graph replay of the same calls
    try:
        side = [torch.cuda.Stream(device=dev) for _ in range(2)]

        def step2():
            main_stream = torch.cuda.current_stream(dev)
            for s in side:
                s.wait_stream(main_stream)
            for i, (v, o) in enumerate(zip(vols, outs)):
                with torch.cuda.stream(side[i % 2]):
                    vt.transform(v, interpolation=interp, output=o, device=device, **kw)
            for s in side:
                main_stream.wait_stream(s)

        step2()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step2()
        g.replay()
        torch.cuda.synchronize()
        sec_g = timed(torch, g.replay, args.steps, 2, lambda: None)
        res['graph_replay_2_streams_value'] = batch * n ** 3 * args.steps / sec_g / 1e9
    except Exception as e:  # informational
        res['graph_replay_2_streams_value'] = None
        res['graph_error'] = repr(e)
        torch.cuda.synchronize()
    # host path: pinned and pageable numpy in -> numpy out, one thread
    h_vol = torch.empty(shape, dtype=torch.float32).pin_memory()
    h_vol.copy_(vols[0])
    h_out = torch.empty(shape, dtype=torch.float32).pin_memory()
    pinned_in, pinned_out = h_vol.numpy(), h_out.numpy()
    pageable_in = np.array(pinned_in, copy=True)
    for label, call in (('pinned', lambda: vt.transform(pinned_in, interpolation=interp, output=pinned_out, device=device, **kw)),
                        ('pageable', lambda: vt.transform(pageable_in, interpolation=interp, device=device, **kw))):
        call()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            call()
        torch.cuda.synchronize()
        res[f'e2e_{label}_numpy_value'] = 5 * n ** 3 / (time.perf_counter() - t0) / 1e9
    return res


def run_project(args, torch, vt, dev):
    """Rotate-and-project: StaticVolume.project_many against the reference's way of doing it
    (transform into a device array, then sum over axis 0), per mode, for a tilt about axis 0 (the reference example's
    matrix family) and for a general matrix.  Gvox/s counts the voxels of the transformed volume that is summed."""
    n = args.size
    shape = (n, n, n)
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    mats = {'tilt30_axis0': vt.utils.transform_matrix(rotation=(30, 0, 0), rotation_order='sxyz', center=c),
            'full_affine': vt.utils.transform_matrix(center=c, **FULL_AFFINE)}
    src = torch.rand(shape, device=f'cuda:{dev}')
    dst = torch.zeros(shape, device=f'cuda:{dev}')
    proj = torch.zeros((1, n, n), device=f'cuda:{dev}')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{dev}')
    res = {}
    for mode in ('linear', 'filt_bspline', 'filt_bspline_simple'):
        sv = vt.StaticVolume(src, interpolation=mode, device=f'gpu:{dev}')
        for mname, m in mats.items():
            def unfused():
                dst.zero_()
                sv.affine(m, output=dst)
                return dst.sum(dim=0)
            ms_u = med_ms(torch, unfused, args.steps, args.warmup, flush)
            ms_f = med_ms(torch, lambda: sv.project_many([m], output=proj), args.steps, args.warmup, flush)
            err = float((proj[0].double() - unfused().double()).abs().max()) / (float(dst.max() - dst.min()) * n)
            res[f'{mode}/{mname}'] = {'fused_ms': ms_f, 'transform_then_sum_ms': ms_u, 'speedup': ms_u / ms_f,
                                      'fused_gvox_s': n ** 3 / ms_f / 1e6, 'max_err_of_scale': err}
    return res


def run_zslab(args, torch, vt, dev, barrier, reduce_max, n=None, steps=None):
    """configs[4]: one large filt_bspline transform, the output split into z-slabs across the ranks.  The raw volume
    starts on rank 0; distribution (prefilter + getting every rank the input footprint of its slab) and resampling are
    timed separately with CUDA events (max over ranks each)."""
    from voltools_b200 import _native, multigpu
    rank, _, world = dist_env()
    n = n or args.size
    steps = steps or args.steps
    shape = (n, n, n)
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    vol = torch.rand(shape, device=f'cuda:{dev}', generator=torch.Generator(f'cuda:{dev}').manual_seed(7)) \
        if rank == 0 else None
    eng = multigpu.CudaEngine(dev)
    res = {}
    for label, kw in (('mild_rotation', MILD), ('full_affine', FULL_AFFINE)):
        m = vt.utils.transform_matrix(center=c, **kw)
        phases = []

        def step():
            t = {}
            multigpu.zslab_affine(vol, m, 'filt_bspline', src=0, engine=eng, shape=shape, timings=t)
            phases.append(t)

        for _ in range(2):
            step()
        phases.clear()
        l0 = _native.launch_count()
        sec = reduce_max(timed(torch, step, steps, 0, barrier))
        launches = _native.launch_count() - l0
        dist_ms = reduce_max(statistics.median(p['distribute_ms']() for p in phases))
        res_ms = reduce_max(statistics.median(p['resample_ms']() for p in phases))
        info = phases[-1].get('info', {})
        res[label] = {'value': n ** 3 * steps / sec / 1e9, 'unit': METRIC, 'ms_per_step': sec / steps * 1e3,
                      'distribute_ms': dist_ms, 'resample_ms': res_ms, 'gpu_launches': int(launches),
                      'inbounds_fraction': inbounds_fraction(shape, m), **info}
    res['workload'] = (f'configs[4]: {n}^3 filt_bspline, raw volume on rank 0, output z-slabs across {world} rank(s); '
                       'strong scaling')
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='all',
                    choices=['all', 'sweep', 'modes', 'zslab', 'cfg1', 'project', 'cpu_baselines'])
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='headline only (no modes_512 / zslab_1024 / cfg1 tables)')
    args = ap.parse_args()
    if args.impl == 'reference':
        reference_arm(args, SWEEP)
        return
    if args.workload == 'cpu_baselines':
        print(json.dumps({'metric': METRIC, 'workload': 'CPU baselines (SURVEY 8d), rot45, one core',
                          'cpu_baselines': run_cpu_baselines(args)}))
        return
    import torch
    import voltools_b200 as vt
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback)')
    dev = local_rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    if world > 1:
        import datetime
        # (a mismatched collective must fail in minutes, not hold N GPUs for the watchdog's default 10)
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{dev}'), timeout=datetime.timedelta(seconds=180))

        def barrier():
            dist.barrier()

        def reduce_max(x):
            t = torch.tensor([x], dtype=torch.float64, device=f'cuda:{dev}')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
    else:
        def barrier():
            pass

        def reduce_max(x):
            return x

    try:
        if args.workload == 'modes':
            res = run_modes(args, torch, vt, dev)
            if rank == 0:
                print(json.dumps({'metric': METRIC, 'workload': f'modes {args.size}^3', 'modes': res}))
        elif args.workload == 'project':
            res = run_project(args, torch, vt, dev)
            if rank == 0:
                print(json.dumps({'metric': METRIC, 'workload': f'rotate-and-project {args.size}^3', 'project': res}))
        elif args.workload == 'cfg1':
            res = run_cfg1(args, torch, vt, dev)
            if rank == 0:
                print(json.dumps({'metric': METRIC, 'cfg1': res}))
        elif args.workload == 'zslab':
            res = run_zslab(args, torch, vt, dev, barrier, reduce_max)
            if rank == 0:
                print(json.dumps({'metric': METRIC, 'n_gpus': world, 'zslab': res}))
        else:
            line = run_sweep(args, torch, vt, dev, barrier, reduce_max)
            if args.workload == 'all' and not args.no_extras:
                torch.cuda.empty_cache()
                line['zslab_1024'] = run_zslab(args, torch, vt, dev, barrier, reduce_max, n=1024, steps=3)
                torch.cuda.empty_cache()
                if world == 1:
                    line['modes_512'] = run_modes(args, torch, vt, dev, n=512, steps=6, warmup=2)
                    torch.cuda.empty_cache()
                    line['cfg1'] = run_cfg1(args, torch, vt, dev)
            if rank == 0:
                if world == 1 and not args.no_cpu_baseline:
                    line['cpu_baseline'] = cpu_baseline_leg(SWEEP)
                print(json.dumps(line))
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
