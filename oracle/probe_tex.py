"""Hardware texture-unit probes (TEST INFRASTRUCTURE ONLY): pins the oracle's tex3D<float> emulation.

Run on the GPU box:  python oracle/probe_tex.py   -> gpurun_out/golden/tex_probe2.npz
Every fetch goes through oracle/_ref/libvt_ref_gpu.so's probe kernel with the reference's own texture
descriptor (voltools/transforms.py:184-192: float32 3-D CUDA array, border, linear filter, unnormalised).

A volume holding a single 1.0 makes the fetch return the filter WEIGHT of that texel, so sweeping the
sample point over the 1/256 grid (and finer) reads the hardware's weight function directly.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402


def grid(levels_a, levels_b, fixed, order):
    a, b = np.meshgrid(levels_a, levels_b, indexing='ij')
    cols = {order[0]: a.ravel(), order[1]: b.ravel(), order[2]: np.full(a.size, fixed)}
    return np.stack([cols['x'], cols['y'], cols['z']], axis=1).astype(np.float32)


def main():
    out = ROOT / 'gpurun_out' / 'golden'
    out.mkdir(parents=True, exist_ok=True)
    data = {}
    delta = np.zeros((4, 4, 4), np.float32)
    delta[1, 1, 1] = 1.0
    lv = (1.5 + np.arange(-256, 257) / 256.0)  # two cells around the texel centre, on the 1/256 grid
    for name, order in (('xy', 'xyz'), ('xz', 'xzy'), ('yz', 'yzx')):
        xyz = grid(lv, lv, 1.5, order)
        data[f'delta_{name}'] = oracle.tex3d_ref_gpu(delta, xyz).reshape(len(lv), len(lv))
    # off-centre third coordinate
    for name, order in (('xy_z1.75', 'xyz'),):
        xyz = grid(lv, lv, 1.75, order)
        data['delta_xy_zq'] = oracle.tex3d_ref_gpu(delta, xyz).reshape(len(lv), len(lv))
    lc = (1.5 + np.arange(-32, 33) / 32.0)
    a, b, c = np.meshgrid(lc, lc, lc, indexing='ij')
    xyz = np.stack([a.ravel(), b.ravel(), c.ravel()], axis=1).astype(np.float32)
    data['delta_xyz'] = oracle.tex3d_ref_gpu(delta, xyz).reshape(65, 65, 65)
    # fine 1-D sweeps along each axis (1/4096 steps) through the delta, other coordinates at the centre
    k = np.arange(0, 8193) / 4096.0 + 0.5
    for ax, nm in enumerate('xyz'):
        xyz = np.full((len(k), 3), 1.5, np.float32)
        xyz[:, ax] = k.astype(np.float32)
        data[f'delta_fine_{nm}'] = oracle.tex3d_ref_gpu(delta, xyz)
    # fine 1-D sweeps along x with y, z off-centre (weights multiply)
    xyz = np.full((len(k), 3), 1.5, np.float32)
    xyz[:, 0] = k.astype(np.float32)
    xyz[:, 1] = 1.5 + 77 / 256.0
    xyz[:, 2] = 1.5 - 45 / 256.0
    data['delta_fine_x_off'] = oracle.tex3d_ref_gpu(delta, xyz)
    # value scaling: is the result weight * value in float32?
    for v in (0.7310586, 123.456, 1e-3):
        d2 = delta * np.float32(v)
        data[f'scaled_{v}'] = oracle.tex3d_ref_gpu(d2, grid(lv[::4], lv[::4], 1.5, 'xyz'))
    # two texels along x: (1-a)*A + a*B or A + a*(B-A)?
    two = np.zeros((4, 4, 8), np.float32)
    two[1, 1, 2], two[1, 1, 3] = 0.3141592, 0.9182817
    kk = np.arange(0, 4097) / 4096.0 + 2.5
    xyz = np.full((len(kk), 3), 1.5, np.float32)
    xyz[:, 0] = kk.astype(np.float32)
    data['two_x'] = oracle.tex3d_ref_gpu(two, xyz)
    data['two_vals'] = np.array([0.3141592, 0.9182817], np.float32)
    # random volume, random coordinates ON the 1/256 grid (no conversion ambiguity) and off it
    rng = np.random.default_rng(99)
    rv = rng.random((9, 10, 11), dtype=np.float32)
    n = 100000
    ongrid = np.stack([rng.integers(-384, 12 * 256 + 128, n), rng.integers(-384, 11 * 256 + 128, n),
                       rng.integers(-384, 10 * 256 + 128, n)], axis=1) / 256.0
    ongrid = ongrid.astype(np.float32)
    offgrid = np.stack([rng.uniform(-1.5, 12.5, n), rng.uniform(-1.5, 11.5, n), rng.uniform(-1.5, 10.5, n)],
                       axis=1).astype(np.float32)
    data['rand_vol'] = rv
    data['rand_on_xyz'] = ongrid
    data['rand_on_out'] = oracle.tex3d_ref_gpu(rv, ongrid)
    data['rand_off_xyz'] = offgrid
    data['rand_off_out'] = oracle.tex3d_ref_gpu(rv, offgrid)
    # large-coordinate behaviour (512-wide axis): does the conversion happen in float32 at large magnitude?
    big = np.zeros((2, 2, 1024), np.float32)
    big[0, 0, 1000] = 1.0
    kb = (np.arange(0, 8193) / 4096.0 + 999.5).astype(np.float32)
    xyz = np.full((len(kb), 3), 0.5, np.float32)
    xyz[:, 0] = kb
    data['big_x'] = kb
    data['big_out'] = oracle.tex3d_ref_gpu(big, xyz)
    np.savez_compressed(out / 'tex_probe2.npz', **data)
    print('wrote', out / 'tex_probe2.npz', {k: v.shape for k, v in data.items()})


if __name__ == '__main__':
    main()
