"""Per-kernel times of the windowed prefilter at a few sizes (through the public binding, with the workspace)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from voltools_b200 import _native  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [250, 256, 512]
st = torch.cuda.current_stream().cuda_stream
for n in sizes:
    shape = (n, n, n)
    src = torch.rand(shape, device='cuda')
    row = _native.padded_row(n)
    dst = torch.empty((n, n, row), device='cuda')
    strides = (row, n * row)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3):
        _native.prefilter(src.data_ptr(), shape, 0, st, dst_ptr=dst.data_ptr(), dst_strides=strides)
    _native.profile_enable(True)
    for _ in range(10):
        flush.zero_()
        _native.prefilter(src.data_ptr(), shape, 0, st, dst_ptr=dst.data_ptr(), dst_strides=strides)
    torch.cuda.synchronize()
    prof = _native.profile_read()
    _native.profile_enable(False)
    tot = sum(v[0] / v[1] for v in prof.values())
    print(f'{n}^3 prefilter (L2 flushed): ' + ' '.join(f'{k}={v[0] / v[1] * 1e3:.1f}us' for k, v in prof.items())
          + f' total={tot * 1e3:.1f}us -> {n ** 3 / tot / 1e6:.0f} Gvox/s ({8 * n ** 3 / tot / 1e6 / 6549.1 * 100:.1f}% of 8 B/vox)')
