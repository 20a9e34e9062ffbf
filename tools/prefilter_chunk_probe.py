"""Does feeding the windowed prefilter in z-chunks keep the XY-filtered intermediate in L2?  Whole prefilter (L2 flushed
before each run) against the streaming form at several chunk depths.   usage: python tools/prefilter_chunk_probe.py [512 1024]"""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from voltools_b200 import _native  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [512]
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def med(fn, reps=7):
    ts = []
    for it in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


for n in sizes:
    shape = (n, n, n)
    src = torch.rand(shape, device='cuda')
    row = _native.padded_row(n)
    strides = (row, n * row)
    dst = torch.empty((n, n, row), device='cuda')
    ws = torch.empty((n, n, row), device='cuda')
    ref = torch.empty((n, n, row), device='cuda')
    _native.prefilter(src.data_ptr(), shape, 0, st, dst_ptr=ref.data_ptr(), dst_strides=strides)
    t = med(lambda: _native.prefilter(src.data_ptr(), shape, 0, st, dst_ptr=dst.data_ptr(), dst_strides=strides))
    print(f'{n}^3 whole prefilter: {t:.3f} ms', flush=True)

    def chunked(c):
        xy_done = 0
        for z0 in range(0, n, c):
            z1 = min(n, z0 + c)
            xy_end = n if z1 == n else min(n, z1 + 12)
            _native.prefilter_planes(src.data_ptr(), ws.data_ptr(), dst.data_ptr(), shape, strides, (xy_done, xy_end), (z0, z1), 0, st)
            xy_done = xy_end

    for c in (8, 16, 32, 64, 128):
        dst.zero_()
        chunked(c)
        torch.cuda.synchronize()
        err = float((dst - ref).abs().max())
        t = med(lambda: chunked(c))
        print(f'{n}^3 streaming form, {c:3d} planes per chunk ({2 * -(-n // c)} launches): {t:.3f} ms   max |d| vs whole {err:.1e}', flush=True)
