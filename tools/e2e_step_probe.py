"""Phases of bench.py's e2e step on one GPU: building the resident volume from a pinned host array, and
affine_many(output=<page-locked array>) at several chunk sizes.   usage: python tools/e2e_step_probe.py"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402

N, K = 256, 180
shape = (N, N, N)
h_vol = vt.pinned_empty(shape)
h_vol[:] = np.random.default_rng(0).random(shape, dtype=np.float32)
c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
mats = np.stack([vt.utils.transform_matrix(rotation=(0, a, 0), rotation_order='rzxz', center=c) for a in range(K)])
h_out = vt.pinned_empty((K,) + shape)
GB = K * N ** 3 * 4 / 1e9


def wall(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


t = wall(lambda: vt.StaticVolume(h_vol, interpolation='filt_bspline', device='gpu:0'))
print(f'StaticVolume(pinned host volume): {t * 1e3:.2f} ms', flush=True)
sv = vt.StaticVolume(h_vol, interpolation='filt_bspline', device='gpu:0')
dev_out = torch.empty((K,) + shape, device='cuda')
t = wall(lambda: sv.affine_many(mats, output=dev_out))
print(f'affine_many -> device: {t * 1e3:.2f} ms', flush=True)
ht = torch.from_numpy(h_out)
t = wall(lambda: ht.copy_(dev_out, non_blocking=True))
print(f'one 11 GiB copy device -> host: {t * 1e3:.2f} ms = {GB / t:.1f} GB/s', flush=True)
for mb in (64, 128, 256, 1024):
    vt.StaticVolume.HOST_CHUNK_BYTES = mb << 20
    t = wall(lambda: sv.affine_many(mats, output=h_out))
    print(f'affine_many -> page-locked host, {mb} MiB chunks: {t * 1e3:.2f} ms = {GB / t:.1f} GB/s', flush=True)
with torch.cuda.stream(torch.cuda.Stream()):
    vt.StaticVolume.HOST_CHUNK_BYTES = 128 << 20
    t = wall(lambda: sv.affine_many(mats, output=h_out))
    print(f'the same on a side stream (not the legacy default stream), 128 MiB chunks: {t * 1e3:.2f} ms = {GB / t:.1f} GB/s', flush=True)


def whole():
    s = vt.StaticVolume(h_vol, interpolation='filt_bspline', device='gpu:0')
    s.affine_many(mats, output=h_out)


t = wall(whole)
print(f'whole step: {t * 1e3:.2f} ms = {GB / t:.1f} GB/s', flush=True)
assert np.array_equal(h_out[90], sv.affine(mats[90]))
