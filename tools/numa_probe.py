import os, sys, time, subprocess
print(subprocess.run('nvidia-smi topo -m | head -14; lscpu | grep -i -E "numa|socket|model name|^CPU\\(s\\)"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /proc/self/status | grep -i -E "cpus_allowed_list|mems_allowed_list"', shell=True, capture_output=True, text=True).stdout)
import torch
def bw(cpus):
    os.sched_setaffinity(0, cpus)
    n = 250**3
    h = torch.empty(n, dtype=torch.float32).pin_memory(); h.fill_(1.0)
    o = torch.empty(n, dtype=torch.float32).pin_memory(); o.fill_(0.0)
    d = torch.empty(n, device='cuda'); d2 = torch.ones(n, device='cuda')
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = []
    for mode in ('h2d', 'd2h', 'both'):
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10):
                if mode in ('h2d', 'both'):
                    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
                if mode in ('d2h', 'both'):
                    with torch.cuda.stream(s2): o.copy_(d2, non_blocking=True)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        res.append(f'{mode} {n*4/dt/1e9:.1f} GB/s')
    print(sorted(cpus)[:1], sorted(cpus)[-1:], ' | '.join(res), flush=True)
allc = sorted(os.sched_getaffinity(0))
print('allowed cpus', allc)
half = len(allc)//2
bw(set(allc[:half])); bw(set(allc[half:])); bw(set(allc))
