"""PCIe reality check for the host path: pinned H2D / D2H / duplex bandwidth and the per-call time of the host API."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import voltools_b200 as vt  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250
shape = (n, n, n)
h_in = torch.rand(shape).pin_memory()
h_out = torch.empty(shape).pin_memory()
d_a = torch.empty(shape, device='cuda')
d_b = torch.rand(shape, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
mb = n ** 3 * 4 / 1e6


def wall(fn, it=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it * 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn in (('H2D', h2d), ('D2H', d2h), ('duplex', both)):
    ms = wall(fn)
    print(f'{name}: {ms:.3f} ms for {mb:.1f} MB each way -> {mb / ms:.1f} GB/s per direction')
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
v, o = h_in.numpy(), h_out.numpy()
for mode in ('linear', 'filt_bspline'):
    for kw, label in ((dict(rotation=(0, 45, 0), rotation_order='rzxz'), 'rot45 (streams)'),
                      (dict(rotation=(30, 45, 60), rotation_order='rzxz'), 'general (waits for the volume)')):
        ms = wall(lambda: vt.transform(v, interpolation=mode, output=o, device='gpu:0', **kw), it=5)
        print(f'{n}^3 {mode} {label}: {ms:.3f} ms per call -> {n ** 3 / ms / 1e6:.2f} Gvox/s')

from voltools_b200 import _native as N  # noqa: E402
ctx = N.HostContext(0)
m = vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c)
for interp, pre, label in ((0, False, 'linear'), (1, True, 'filt_bspline')):
    ms = wall(lambda: ctx.affine(v, o, m, interp, pre), it=5)
    print(f'{n}^3 {label} rot45 direct vt_host_affine_f32: {ms:.3f} ms')
t0 = time.perf_counter()
for _ in range(100):
    vt.utils.transform_matrix(rotation=(0, 45, 0), rotation_order='rzxz', center=c)
print(f'transform_matrix: {(time.perf_counter() - t0) * 10:.3f} ms')

# independent volumes through transform() from T host threads (each thread has its own host context)
from concurrent.futures import ThreadPoolExecutor  # noqa: E402
B = 8
vs = [torch.rand(shape).pin_memory().numpy() for _ in range(B)]
os_ = [torch.empty(shape).pin_memory().numpy() for _ in range(B)]
kw = dict(rotation=(0, 45, 0), rotation_order='rzxz')
for T in (1, 2, 3, 4):
    pool = ThreadPoolExecutor(T)

    def step():
        list(pool.map(lambda i: vt.transform(vs[i], interpolation='filt_bspline', output=os_[i], device='gpu:0', **kw),
                      range(B)))
    ms = wall(step, it=4) / B
    print(f'{n}^3 filt_bspline rot45, {B} volumes from {T} host threads: {ms:.3f} ms per volume -> '
          f'{n ** 3 / ms / 1e6:.2f} Gvox/s')
    pool.shutdown()
