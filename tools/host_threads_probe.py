"""Timeline of the host path when transform() is called from T threads (VT_HOST_DEBUG=1 prints one line per call)."""
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = 8
shape = (n, n, n)
vs = [torch.rand(shape).pin_memory().numpy() for _ in range(B)]
os_ = [torch.empty(shape).pin_memory().numpy() for _ in range(B)]
kw = dict(rotation=(0, 45, 0), rotation_order='rzxz')
pool = ThreadPoolExecutor(T)
t00 = time.perf_counter()


def one(i):
    t0 = time.perf_counter()
    vt.transform(vs[i], interpolation='filt_bspline', output=os_[i], device='gpu:0', **kw)
    t1 = time.perf_counter()
    return i, (t0 - t00) * 1e3, (t1 - t00) * 1e3


for rep in range(3):
    t00 = time.perf_counter()
    res = list(pool.map(one, range(B)))
    tot = (time.perf_counter() - t00) * 1e3
    print(f'rep {rep}: {tot:.3f} ms for {B} volumes ({tot / B:.3f} ms each)')
    for i, a, b in res:
        print(f'   call {i}: host start {a:.3f} end {b:.3f} ms')
