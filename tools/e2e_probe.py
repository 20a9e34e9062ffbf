"""Where the N-GPU end-to-end (host results) number of bench.py goes: concurrent device->host traffic of all ranks, layer
by layer.   usage: torchrun --nproc-per-node N tools/e2e_probe.py      (1 rank works too)

 A  raw      one pinned 64 MiB buffer per rank, 16 async copies, all ranks at once (and each rank alone): the platform
 B  download _native.download (fresh torch pinned buffer per call -- the product's `.get()`)
 C  affine   sv.affine(m) -> numpy, the call bench.py's e2e makes
 D  raw, after binding the rank's threads to the GPU's local CPUs (when sysfs / NVML expose them)
Every phase: barrier, wall clock, max over ranks; fixed iteration counts (no rank-local loop bounds)."""
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    import datetime
    dist.init_process_group('nccl', timeout=datetime.timedelta(seconds=90), device_id=torch.device(f'cuda:{local}'))
N, REPS = 256, 16
BYTES = N ** 3 * 4


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather_times(sec):
    t = torch.tensor([sec], dtype=torch.float64, device='cuda')
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o) for o in out]
    return [sec]


def timed(name, fn, solo=False):
    fn()  # warm
    if solo:
        per = []
        for r in range(world):
            barrier()
            t0 = time.perf_counter()
            if r == rank:
                fn()
                torch.cuda.synchronize()
            per.append(time.perf_counter() - t0)
        ts = gather_times(per[rank])
        if rank == 0:
            print(f'{name:34s} alone: ' + ' '.join(f'{REPS * BYTES / t / 1e9:5.1f}' for t in ts) + ' GB/s per rank', flush=True)
        return
    barrier()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    ts = gather_times(time.perf_counter() - t0)
    if rank == 0:
        print(f'{name:34s} all at once: ' + ' '.join(f'{REPS * BYTES / t / 1e9:5.1f}' for t in ts) +
              f' GB/s per rank; aggregate {world * REPS * BYTES / max(ts) / 1e9:6.1f} GB/s', flush=True)


if rank == 0:
    print(subprocess.run('nvidia-smi topo -m | head -24; lscpu | grep -i -E "numa|socket|^CPU\\(s\\)"; '
                         'for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; '
                         'then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done; free -g | head -2',
                         shell=True, capture_output=True, text=True).stdout, flush=True)
    print('allowed cpus', len(os.sched_getaffinity(0)), flush=True)

d = torch.rand((N, N, N), device='cuda')
o = torch.empty((N, N, N), dtype=torch.float32).pin_memory()


def raw():
    for _ in range(REPS):
        o.copy_(d, non_blocking=True)


def download():
    for _ in range(REPS):
        _native.download(d, torch.cuda.current_stream().cuda_stream)


sv = vt.StaticVolume(d, interpolation='filt_bspline', device=f'gpu:{local}')
c = np.divide(np.subtract((N, N, N), 1), 2, dtype=np.float32)
mats = [vt.utils.transform_matrix(rotation=(0, a, 0), rotation_order='rzxz', center=c) for a in range(REPS)]


def affine():
    for m in mats:
        sv.affine(m)


timed('A raw pinned copies', raw, solo=True)
timed('A raw pinned copies', raw)
timed('B _native.download', download)
timed('C sv.affine -> numpy', affine)
timed('C sv.affine -> numpy', affine, solo=True)

# D: bind to the GPU's CPUs
cpus = None
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
except Exception as e:  # noqa: BLE001
    if rank == 0:
        print('nvml affinity unavailable:', e)
ts = gather_times(float(len(cpus) if cpus else 0))
if rank == 0:
    print('local cpus per rank (NVML):', ts, flush=True)
if cpus and cpus & os.sched_getaffinity(0) and len(cpus) < len(os.sched_getaffinity(0)):
    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0))
    o = torch.empty((N, N, N), dtype=torch.float32).pin_memory()  # first touch on the bound CPUs
    o.fill_(0)
    timed('D raw, bound to local cpus', raw)
    timed('D sv.affine, bound to local cpus', affine)
elif rank == 0:
    print('D skipped: the GPU-local CPU set is the whole machine (no NUMA exposed)', flush=True)

# E: the e2e step of bench.py, phase by phase
if world > 1:
    from voltools_b200 import multigpu
    eng = multigpu.CudaEngine(local)
    h_vol = d.cpu().numpy() if rank == 0 else None
    for it in range(2):
        barrier()
        t0 = time.perf_counter()
        buf, width = multigpu.prepare_and_broadcast(eng, h_vol, 'filt_bspline', src=0, shape=(N, N, N))
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sv2 = vt.StaticVolume.from_coefficients(buf, 'filt_bspline', width)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for m in mats:
            sv2.affine(m)
        t3 = time.perf_counter()
        a, b, cc = gather_times(t1 - t0), gather_times(t2 - t1), gather_times(t3 - t2)
        if rank == 0:
            print(f'E it{it}: prepare+broadcast ms ' + ' '.join(f'{x * 1e3:.1f}' for x in a) + ' | from_coefficients ms ' +
                  ' '.join(f'{x * 1e3:.1f}' for x in b) + f' | {REPS} x affine->numpy ms ' + ' '.join(f'{x * 1e3:.1f}' for x in cc), flush=True)
    barrier()
    dist.destroy_process_group()
