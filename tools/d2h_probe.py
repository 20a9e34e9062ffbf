"""Device->host rate into ONE 64 MiB page-locked buffer against a stream into a large (11 GiB) page-locked destination
-- what the e2e leg of bench.py does.   usage: python tools/d2h_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402

N, K = 256, 180
d = torch.rand((2, N, N, N), device='cuda')
small = torch.empty((N, N, N), dtype=torch.float32).pin_memory()


def rate(name, fn, nbytes):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'{name:60s} {nbytes / dt / 1e9:6.1f} GB/s', flush=True)


rate('same 64 MiB buffer x 180', lambda: [small.copy_(d[0], non_blocking=True) for _ in range(K)], K * N ** 3 * 4)
t0 = time.perf_counter()
big = vt.pinned_empty((K, N, N, N))
print(f'pinned_empty of {big.nbytes / 2**30:.2f} GiB (cudaHostRegister): {time.perf_counter() - t0:.2f} s', flush=True)
big_t = torch.from_numpy(big)
rate('180 x 64 MiB into the registered 11 GiB array', lambda: [big_t[k].copy_(d[0], non_blocking=True) for k in range(K)], K * N ** 3 * 4)
rate('90 x 128 MiB into the registered 11 GiB array', lambda: [big_t[2 * k:2 * k + 2].copy_(d, non_blocking=True) for k in range(K // 2)], K * N ** 3 * 4)
del big_t, big
t0 = time.perf_counter()
chunks = [torch.empty((16, N, N, N), dtype=torch.float32, pin_memory=True) for _ in range(4)]   # 4 x 1 GiB cudaHostAlloc
print(f'4 x 1 GiB cudaHostAlloc: {time.perf_counter() - t0:.2f} s', flush=True)
rate('64 x 64 MiB into 4 GiB of cudaHostAlloc memory', lambda: [chunks[k // 16][k % 16].copy_(d[0], non_blocking=True) for k in range(64)], 64 * N ** 3 * 4)
s2 = torch.cuda.Stream()


def two_streams():
    for k in range(32):
        chunks[0][k % 16].copy_(d[0], non_blocking=True)
        with torch.cuda.stream(s2):
            chunks[2][k % 16].copy_(d[1], non_blocking=True)


rate('two streams, 64 x 64 MiB into 2 GiB', two_streams, 64 * N ** 3 * 4)
import subprocess
print(subprocess.run('cat /sys/kernel/mm/transparent_hugepage/enabled; free -g | head -2', shell=True, capture_output=True, text=True).stdout)
