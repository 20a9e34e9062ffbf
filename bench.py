#!/usr/bin/env python
"""bench.py -- Gvoxels/s of the affine volume-resampling hot path on N B200s (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`, inputs per SURVEY.md section 8d):
  cfg1    (default) configs[1]: transform() of 250^3 float32 volumes, interpolation='filt_bspline' (prefilter +
          8-fetch cubic), rotation=(0,45,0) 'rzxz' about the centre.  One step = one pass over a batch of 8
          independent volumes (500 MB in + 500 MB out, larger than the 126 MB L2), each through the public
          voltools_b200.transform(...).  N > 1: every rank owns its own batch (independent objects, no
          collective) -> weak scaling.
  affine512  configs[3]: 512^3 bspline_simple, full affine G6, output= device array.
  sweep   configs[2]: StaticVolume 256^3 filt_bspline, 180-angle sweep; rank 0 prefilters, one NCCL broadcast of
          the coefficient volume, the angles are split across ranks (strong scaling).
  modes   all five interpolation modes at 512^3 (rot45 and full affine) -> reported under "modes" (1 GPU).
  cpu_baselines  reference CPU path and scipy.ndimage.affine_transform at 100^3 .. 512^3 on this host (no GPU work).
  project rotate-and-project at 512^3: StaticVolume.project_many vs transform + sum(axis=0) (1 GPU).

`value` is device-resident throughput (inputs already in HBM, CUDA events); `e2e` is the same batch through the
same public call with pinned HOST arrays in and out (H2D + kernels + D2H inside the timed region).
`roofline` is for the kernel with the largest share of the step, from per-launch CUDA events recorded by the
library (vt_profile_*), against MEASURED_PEAKS.json.  `cpu_baseline` times the unmodified reference package's
CPU path (voltools.transform(device='cpu') -> scipy.ndimage.affine_transform), staged under oracle/_ref/py, on
this host.  `--impl reference` runs that CPU path alone as the reference arm.
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
E2E_THREADS = 2  # host threads issuing the e2e leg's independent transform() calls (PCIe duplex overlap)
ROT45 = dict(rotation=(0, 45, 0), rotation_order='rzxz')
FULL_AFFINE = dict(scale=(1.1, 0.9, 1.05), shear=(0.05, -0.03, 0.02), rotation=(30, 45, 60), rotation_order='rzxz',
                   translation=(5.5, -3.25, 2.0))
CFG1 = dict(n=250, batch=8, interpolation='filt_bspline', kw=ROT45,
            name="configs[1]: transform() 250^3 float32 interpolation='filt_bspline' rotation=(0,45,0) rzxz; "
                 "step = batch of 8 independent volumes")


def peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


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


def cpu_baseline_leg(cfg):
    """Reference CPU path on ONE full volume of the workload, one core (scipy.ndimage is single-threaded)."""
    ref_vt = _ref_voltools()
    if ref_vt is None:
        return {'value': None, 'unit': METRIC, 'cores': 0, 'kind': 'reference', 'sample': 'reference package not staged'}
    n = cfg['n']
    dt, _ = _ref_worker((0, (n, n, n), cfg['interpolation'], cfg['kw']))
    out = {'value': n ** 3 / dt / 1e9, 'unit': METRIC, 'cores': 1, 'kind': 'reference',
           'sample': f"1 of the step's {cfg['batch']} volumes ({n}^3) through the unmodified reference "
                     f"voltools.transform(device='cpu') = scipy.ndimage.affine_transform(order=3, prefilter=True); "
                     f'{dt:.2f} s on 1 of {os.cpu_count()} host cores'}
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
    """--impl reference: the reference's own CPU implementation, all the host parallelism it can use (one process
    per independent volume of the batch; each call is single-threaded inside SciPy)."""
    import multiprocessing as mp
    rank, _, world = dist_env()
    if rank != 0:
        return
    if _ref_voltools() is None:
        print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/_ref/py (reference package) not staged'}))
        return
    n = cfg['n']
    procs = max(1, min(os.cpu_count() or 1, cfg['batch']))
    # bounded sample: a z-slab of `planes` planes of each volume (in-plane rotation about axis 0: every output
    # plane costs the same), sized so that (steps + warmup) steps fit in ~150 s
    t_probe, _ = _ref_worker((0, (16, n, n), cfg['interpolation'], cfg['kw']))
    per_plane = t_probe / 16
    budget = 150.0 / (args.steps + args.warmup)
    planes = int(max(8, min(n, budget / per_plane)))
    shape = (planes, n, n)
    jobs = [(1000 + i, shape, cfg['interpolation'], cfg['kw']) for i in range(procs)]
    with mp.get_context('fork').Pool(procs) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, jobs)
        dt = time.perf_counter() - t0
    vox = procs * planes * n * n * args.steps
    value = vox / dt / 1e9
    sample = (f'{procs} processes x one {planes}x{n}x{n} z-slab of a {n}^3 volume per step, unmodified reference '
              f"voltools.transform(interpolation='{cfg['interpolation']}', device='cpu')")
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
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


def run_cfg1(args, cfg, torch, vt, dev, barrier, reduce_max):
    from voltools_b200 import _native
    n, batch, interp, kw = cfg['n'], cfg['batch'], cfg['interpolation'], cfg['kw']
    shape = (n, n, n)
    device = f'gpu:{dev}'
    rank, _, world = dist_env()
    gen = torch.Generator(device=f'cuda:{dev}').manual_seed(1234 + rank)
    vols = [torch.rand(shape, generator=gen, device=f'cuda:{dev}', dtype=torch.float32) for _ in range(batch)]
    outs = [torch.zeros(shape, device=f'cuda:{dev}', dtype=torch.float32) for _ in range(batch)]

    # The batch's volumes are independent: the calls are issued round-robin on `args.streams` CUDA streams (the API is
    # stream-ordered on torch's current stream), so one volume's small kernels (250 CTAs do not fill 148 SMs x 2-3
    # resident CTAs) overlap the next volume's.  The timed region starts and ends on the default stream, which the
    # side streams fork from and join back into every step.
    side = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else []

    def step_on(streams):
        if not streams:
            for v, o in zip(vols, outs):
                vt.transform(v, interpolation=interp, output=o, device=device, **kw)
            return
        main_stream = torch.cuda.current_stream(dev)  # the capture stream while a CUDA graph is being recorded
        for s in streams:
            s.wait_stream(main_stream)
        for i, (v, o) in enumerate(zip(vols, outs)):
            with torch.cuda.stream(streams[i % len(streams)]):
                vt.transform(v, interpolation=interp, output=o, device=device, **kw)
        for s in streams:
            main_stream.wait_stream(s)

    def step_eager():
        step_on(side)

    for _ in range(max(3, args.warmup)):
        step_eager()
    torch.cuda.synchronize()
    # A step is 24 kernels of 30-80 us behind 8 Python calls: the host barely keeps ahead of the GPU (and falls behind
    # with several ranks sharing the host's cores).  The same calls are therefore recorded once into a CUDA graph --
    # same kernels, same streams, same buffers -- and the timed steps replay it.  --no-graph times the eager calls.
    graph, launches_per_step = None, None
    if not args.no_graph:
        try:
            l_cap = _native.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step_eager()
            launches_per_step = _native.launch_count() - l_cap
            g.replay()
            torch.cuda.synchronize()
            graph = g
        except Exception as e:  # capture not possible here: eager calls
            print(f'bench: CUDA graph capture failed ({e!r}); timing eager calls', file=sys.stderr)
            torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            step_eager()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    uuid = str(getattr(torch.cuda.get_device_properties(dev), 'uuid', dev))
    clocks = ClockSampler(uuid if uuid.startswith('GPU-') or uuid.isdigit() else 'GPU-' + uuid)
    # keep the GPU under the same load for >= 0.5 s before the timed region so that nvidia-smi (100 ms period)
    # gets samples of the clocks this workload runs at; the sampler stays on through the timed region
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.6:
        step()
        torch.cuda.synchronize()
    l0 = _native.launch_count()
    sec = timed(torch, step, args.steps, 0, barrier)
    launches = launches_per_step * args.steps if graph is not None else _native.launch_count() - l0
    clk = clocks.stop()
    sec = reduce_max(sec)
    vox_step = batch * n ** 3
    value = world * vox_step * args.steps / sec / 1e9
    single_stream_value = value
    if side:
        sec1 = reduce_max(timed(torch, lambda: step_on([]), args.steps, 2, barrier))
        single_stream_value = world * vox_step * args.steps / sec1 / 1e9
    eager_value = value
    if graph is not None:
        sec2 = reduce_max(timed(torch, step_eager, args.steps, 2, barrier))
        eager_value = world * vox_step * args.steps / sec2 / 1e9

    # per-kernel durations over the same K steps (CUDA events inside the library, on the launch stream), with the
    # calls on ONE stream so that each kernel is timed alone
    _native.profile_enable(True)
    for _ in range(args.steps):
        step_on([])
    torch.cuda.synchronize()
    prof = _native.profile_read()
    _native.profile_enable(False)
    peak, peak_src = peaks()
    kernels = {}
    total_ms = sum(ms for ms, _ in prof.values()) or 1.0
    for name, (ms, cnt) in prof.items():
        avg = ms / cnt
        ach = 8.0 * n ** 3 / (avg * 1e-3) / 1e9
        kernels[name] = {'launches_per_step': cnt / args.steps, 'avg_ms': avg, 'share': ms / total_ms,
                         'achieved_gbs': ach, 'frac': ach / peak}
    dom = max(prof, key=lambda k: prof[k][0]) if prof else None
    traffic = None
    tfile = ROOT / 'profiles' / 'traffic.json'
    if dom and tfile.exists():
        traffic = json.loads(tfile.read_text()).get(dom, {}).get('dram_bytes_per_launch')
    roofline = None
    if dom:
        roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': kernels[dom]['achieved_gbs'], 'peak': peak,
                    'unit': 'GB/s', 'frac': kernels[dom]['frac'], 'traffic': traffic, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': 8 * n ** 3,
                    'step_frac': 16.0 * vox_step * args.steps / sec / 1e9 / peak,
                    'note': 'achieved = 8 B/voxel x 250^3 voxels / average launch duration; step_frac = 16 B/voxel '
                            '(prefilter 8 + resample 8) x voxels per step / step time / peak',
                    'kernels': kernels}

    # end to end: pinned host arrays in and out through the same public call
    h_vols = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(batch)]
    h_outs = [torch.empty(shape, dtype=torch.float32).pin_memory() for _ in range(batch)]
    for h, v in zip(h_vols, vols):
        h.copy_(v)
    np_vols, np_outs = [h.numpy() for h in h_vols], [h.numpy() for h in h_outs]

    def one_e2e(i):
        vt.transform(np_vols[i], interpolation=interp, output=np_outs[i], device=device, **kw)

    # The batch's volumes are independent, so the public call is made from E2E_THREADS host threads (ctypes releases
    # the GIL inside the library): one call's upload overlaps another's download on the full-duplex PCIe link.  The
    # single-thread figure (calls strictly one after another) is reported next to it.
    from concurrent.futures import ThreadPoolExecutor
    e2e_steps = max(2, min(args.steps, 5))
    e2e_by_threads = {}
    for nthreads in (1, E2E_THREADS):
        pool = ThreadPoolExecutor(nthreads) if nthreads > 1 else None

        def step_e2e():
            if pool is None:
                for i in range(batch):
                    one_e2e(i)
            else:
                list(pool.map(one_e2e, range(batch)))

        step_e2e()
        step_e2e()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_by_threads[nthreads] = reduce_max(time.perf_counter() - t0)
        barrier()
        if pool is not None:
            pool.shutdown()
    e2e_sec = e2e_by_threads[E2E_THREADS]
    e2e_value = world * vox_step * e2e_steps / e2e_sec / 1e9
    # the host path must agree with the device path
    err = float((torch.from_numpy(np_outs[0]).to(f'cuda:{dev}') - outs[0]).abs().max())
    assert err <= 1e-5 * 16, f'host path differs from device path: {err}'

    line = {
        'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': cfg['name'], 'voxels_per_step_per_gpu': vox_step, 'parallelism': f'dp{world}',
                   'l2': 'inputs larger than L2: 8 distinct volumes per step (500 MB in + 500 MB out per GPU)',
                   'call': "voltools_b200.transform(vol, rotation=(0,45,0), rotation_order='rzxz', "
                           "interpolation='filt_bspline', output=out, device='gpu:X')",
                   'cuda_streams': max(1, args.streams), 'cuda_graph': graph is not None,
                   'eager_value': eager_value, 'single_stream_eager_value': single_stream_value},
        'clocks': {'sm_mhz': clk['sm_mhz'], 'sm_max_mhz': clk['sm_max_mhz'], 'reasons': clk['reasons'],
                   'samples': clk['samples']},
        'e2e': {'value': e2e_value, 'unit': METRIC, 'h2d_bytes_per_step': vox_step * 4, 'd2h_bytes_per_step': vox_step * 4,
                'steps': e2e_steps, 'ms_per_step': e2e_sec / e2e_steps * 1e3,
                'call': f'same call with pinned numpy arrays for volume and output, issued from {E2E_THREADS} host threads '
                        '(independent volumes)', 'host_threads': E2E_THREADS,
                'single_thread_value': world * vox_step * e2e_steps / e2e_by_threads[1] / 1e9},
        'gpu_launches': int(launches),
        'roofline': roofline,
    }
    return line


def run_modes(args, torch, vt, dev):
    """All five modes at 512^3 (north-star target shape): kernel-only Gvox/s, rot45 and full affine."""
    from voltools_b200 import _native
    n = args.size
    shape = (n, n, n)
    peak, _ = peaks()
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    mats = {'rot45': vt.utils.transform_matrix(center=c, **ROT45),
            'full_affine': vt.utils.transform_matrix(center=c, **FULL_AFFINE)}
    src = torch.rand(shape, device=f'cuda:{dev}')
    dst = torch.zeros(shape, device=f'cuda:{dev}')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{dev}')
    res = {}
    for mode in vt.AVAILABLE_INTERPOLATIONS:
        for mname, m in mats.items():
            ts = []
            for it in range(args.warmup + args.steps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                vt.affine(src, m, interpolation=mode, output=dst, device=f'gpu:{dev}')
                e1.record()
                e1.synchronize()
                if it >= args.warmup:
                    ts.append(e0.elapsed_time(e1))
            ms = statistics.median(ts)
            bytes_per_vox = 16 if mode.startswith('filt') else 8
            res[f'{mode}/{mname}'] = {'ms': ms, 'gvox_s': n ** 3 / ms / 1e6,
                                      'roofline_frac': bytes_per_vox * n ** 3 / (ms * 1e-3) / 1e9 / peak}
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

    def med(fn):
        ts = []
        for it in range(args.warmup + args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            if it >= args.warmup:
                ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    res = {}
    for mode in ('linear', 'filt_bspline', 'filt_bspline_simple'):
        sv = vt.StaticVolume(src, interpolation=mode, device=f'gpu:{dev}')
        for mname, m in mats.items():
            def unfused():
                dst.zero_()
                sv.affine(m, output=dst)
                return dst.sum(dim=0)
            ms_u = med(unfused)
            ms_f = med(lambda: sv.project_many([m], output=proj))
            err = float((proj[0].double() - unfused().double()).abs().max()) / (float(dst.max() - dst.min()) * n)
            res[f'{mode}/{mname}'] = {'fused_ms': ms_f, 'transform_then_sum_ms': ms_u, 'speedup': ms_u / ms_f,
                                      'fused_gvox_s': n ** 3 / ms_f / 1e6, 'max_err_of_scale': err}
    return res


def run_sweep(args, torch, vt, dev, barrier, reduce_max):
    """configs[2]: StaticVolume 256^3 filt_bspline, 180-angle sweep rotation=(0,i,0); root prefilters, one broadcast of
    the coefficient buffer (NCCL), angles split across ranks.  Strong scaling: the 180 angles are the fixed job."""
    from voltools_b200 import _native, multigpu
    import torch.distributed as dist
    rank, _, world = dist_env()
    n = args.size if args.size != 512 else 256
    shape = (n, n, n)
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    mats = [vt.utils.transform_matrix(rotation=(0, i, 0), rotation_order='rzxz', center=c) for i in range(180)]
    vol = torch.rand(shape, device=f'cuda:{dev}', generator=torch.Generator(f'cuda:{dev}').manual_seed(7)) \
        if rank == 0 else None
    mine = multigpu.split_strided(len(mats), world, rank)
    out = torch.empty((len(mine),) + shape, device=f'cuda:{dev}')
    eng = multigpu.CudaEngine(dev)

    def step():
        if world > 1:
            outs, _ = multigpu.sweep(vol, mats, 'filt_bspline', src=0, engine=eng, shape=shape)
        else:
            buf, width = eng.prepare(vol, 'filt_bspline')
            sv = vt.StaticVolume.from_coefficients(buf, 'filt_bspline', width)
            sv.affine_many(mats, output=out, zero_fill=True)

    l0 = _native.launch_count()
    sec = reduce_max(timed(torch, step, args.steps, max(3, args.warmup), barrier))
    launches = _native.launch_count() - l0
    value = len(mats) * n ** 3 * args.steps / sec / 1e9
    peak, _ = peaks()
    return {'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'configs[2]: StaticVolume {n}^3 filt_bspline, 180-angle sweep rotation=(0,i,0), '
                                   'prefilter on rank 0 + 1 NCCL broadcast + angles split across ranks',
                       'parallelism': f'dp{world}', 'l2': 'outputs (180 volumes) exceed L2; the coefficient volume is '
                                                         'L2-resident by design'},
            'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'note': '8 B/voxel/matrix', 'frac': 8.0 * value / peak / world, 'peak': peak}}


def run_zslab(args, torch, vt, dev, barrier, reduce_max):
    """configs[4]: one large filt_bspline full-affine transform, output z-slabs split across ranks."""
    from voltools_b200 import _native, multigpu
    rank, _, world = dist_env()
    n = args.size
    shape = (n, n, n)
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    m = vt.utils.transform_matrix(center=c, **FULL_AFFINE)
    vol = torch.rand(shape, device=f'cuda:{dev}', generator=torch.Generator(f'cuda:{dev}').manual_seed(7)) \
        if rank == 0 else None
    eng = multigpu.CudaEngine(dev)

    def step():
        if world > 1:
            multigpu.zslab_affine(vol, m, 'filt_bspline', src=0, engine=eng, shape=shape)
        else:
            buf, width = eng.prepare(vol, 'filt_bspline')
            eng.resample_slab(buf, width, 'filt_bspline', m, 0, n)

    l0 = _native.launch_count()
    sec = reduce_max(timed(torch, step, args.steps, max(3, args.warmup), barrier))
    launches = _native.launch_count() - l0
    value = n ** 3 * args.steps / sec / 1e9
    return {'metric': METRIC, 'value': value, 'unit': METRIC, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': sec / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'configs[4]: {n}^3 filt_bspline full affine, prefilter on rank 0 + NCCL broadcast of '
                                   'the coefficients + output z-slabs across ranks', 'parallelism': f'dp{world}'},
            'gpu_launches': int(launches)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg1', choices=['cfg1', 'modes', 'sweep', 'zslab', 'project', 'cpu_baselines'])
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='time eager calls instead of replaying a captured CUDA graph')
    ap.add_argument('--streams', type=int, default=2, help="CUDA streams the step's independent volumes are issued on")
    ap.add_argument('--batch', type=int, default=None, help='volumes per step (default 8; smaller only for profiling)')
    args = ap.parse_args()
    cfg = dict(CFG1)
    if args.batch:
        cfg['batch'] = args.batch
        cfg['name'] = cfg['name'].replace('batch of 8', f'batch of {args.batch}')
    if args.impl == 'reference':
        reference_arm(args, cfg)
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{dev}'))

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

    if args.workload == 'modes':
        res = run_modes(args, torch, vt, dev)
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'workload': f'modes {args.size}^3', 'modes': res}))
        return
    if args.workload == 'project':
        res = run_project(args, torch, vt, dev)
        if rank == 0:
            print(json.dumps({'metric': METRIC, 'workload': f'rotate-and-project {args.size}^3', 'project': res}))
        return
    if args.workload in ('sweep', 'zslab'):
        if world == 1:  # single-process: the helpers still want a process group for get_rank()
            import torch.distributed as dist
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            os.environ.setdefault('MASTER_PORT', '29533')
            dist.init_process_group('nccl', rank=0, world_size=1, device_id=torch.device(f'cuda:{dev}'))
        line = (run_sweep if args.workload == 'sweep' else run_zslab)(args, torch, vt, dev, barrier, reduce_max)
        if rank == 0:
            print(json.dumps(line))
        import torch.distributed as dist
        dist.destroy_process_group()
        return
    line = run_cfg1(args, cfg, torch, vt, dev, barrier, reduce_max)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_leg(cfg)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
