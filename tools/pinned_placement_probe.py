"""Does the device->host rate depend on WHICH page-locked buffer is the destination (physical placement the guest cannot
see)?  16 separate 1 GiB buffers (8 cudaHostAlloc, 8 registered numpy), 16 x 64 MiB copies into each, three rounds.
usage: python tools/pinned_placement_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from voltools_b200 import _native  # noqa: E402

N = 256
d = torch.rand((N, N, N), device='cuda')
bufs = []
for i in range(8):
    bufs.append(('alloc', torch.empty((16, N, N, N), dtype=torch.float32, pin_memory=True)))
_native.PINNED_CACHE_LIMIT = 0
for i in range(8):
    a = _native.pinned_empty((16, N, N, N))
    bufs.append(('reg', torch.from_numpy(a)))
for rnd in range(3):
    row = []
    for kind, b in bufs:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(16):
            b[k].copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        row.append(f'{kind}:{16 * N ** 3 * 4 / (time.perf_counter() - t0) / 1e9:.1f}')
    print(f'round {rnd}: ' + ' '.join(row), flush=True)
