"""GPU probe of the slice4 (Z4 layout) family: parity against the slice / gather families for all three march axes,
and kernel timings over angles at the given sizes.   usage: python tools/z4_probe.py [256 512] [--quick]"""
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import voltools_b200 as vt  # noqa: E402
from voltools_b200 import _native  # noqa: E402

dev = 0
MODES = {'linear': _native.LINEAR, 'cubic_tex': _native.CUBIC_TEX, 'cubic_simple': _native.CUBIC_SIMPLE}
ORDERS = {0: ('rzxz', lambda a: (0, a, 0)), 1: ('ryzy', lambda a: (a, 0, 0)), 2: ('rzxz', lambda a: (a, 0, 0))}


def z4_of(src, axis):
    buf = torch.empty(_native.z4_bytes(src.shape, axis) // 4, dtype=torch.float32, device=src.device)
    _native.pack_z4(src.data_ptr(), src.shape, buf.data_ptr(), axis, device=dev, stream=torch.cuda.current_stream().cuda_stream)
    return buf


def mat_for(shape, axis, angle, shift=(0, 0, 0)):
    c = np.divide(np.subtract(shape, 1), 2, dtype=np.float32)
    order, rot = ORDERS[axis]
    return vt.utils.transform_matrix(rotation=rot(angle), rotation_order=order, center=c, translation=shift)


def parity():
    torch.manual_seed(0)
    worst = 0.0
    for shape in ((40, 44, 48), (33, 50, 37), (64, 64, 64), (70, 9, 130)):
        src = torch.rand(shape, device=f'cuda:{dev}')
        for axis in range(3):
            z4 = z4_of(src, axis)
            for angle, shift in ((0, (0, 0, 0)), (17, (0, 0, 0)), (45, (0, 0, 0)), (90, (0, 0, 0)), (133, (2, -3, 5))):
                m = mat_for(shape, axis, angle, shift)
                ax = _native.z4_axis(shape, shape, m, 0)
                assert ax >= 0 and (angle == 0 or ax == axis), (axis, angle, ax, m)
                for name, interp in MODES.items():
                    for flags in (_native.OOB_ZERO, _native.OOB_SKIP):
                        want = torch.full(shape, -7.0, device=src.device)
                        got = torch.full(shape, -7.0, device=src.device)
                        _native.affine(src.data_ptr(), shape, want.data_ptr(), shape, m, interp, flags | _native.KERNEL_GATHER,
                                       device=dev, stream=torch.cuda.current_stream().cuda_stream)
                        _native.affine_z4(z4.data_ptr(), axis, shape, got.data_ptr(), shape, m, interp, flags, device=dev,
                                          stream=torch.cuda.current_stream().cuda_stream)
                        torch.cuda.synchronize()
                        e = float((got - want).abs().max())
                        skipped_same = bool(((got == -7.0) == (want == -7.0)).all())
                        worst = max(worst, e)
                        if e > 2e-6 or not skipped_same:
                            print('MISMATCH', shape, axis, angle, name, flags, e, skipped_same)
                            return False
                # z-range launch
                got = torch.zeros(shape, device=src.device)
                want = torch.zeros(shape, device=src.device)
                z0, z1 = shape[0] // 3, shape[0] - 5
                _native.affine(src.data_ptr(), shape, want.data_ptr(), shape, m, 1, _native.KERNEL_GATHER, z_range=(z0, z1),
                               device=dev, stream=torch.cuda.current_stream().cuda_stream)
                _native.affine_z4(z4.data_ptr(), axis, shape, got.data_ptr(), shape, m, 1, 0, z_range=(z0, z1), device=dev,
                                  stream=torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                e = float((got - want).abs().max())
                if e > 2e-6:
                    print('MISMATCH z-range', shape, axis, angle, e)
                    return False
        # axis 0 against the plain-layout slice family: bit-identical
        z4 = z4_of(src, 0)
        if shape[2] % 4 == 0:
            for angle in (0, 17, 45):
                m = mat_for(shape, 0, angle)
                for name, interp in MODES.items():
                    a = torch.zeros(shape, device=src.device)
                    b = torch.zeros(shape, device=src.device)
                    _native.affine(src.data_ptr(), shape, a.data_ptr(), shape, m, interp, _native.OOB_ZERO | _native.KERNEL_SLICE,
                                   device=dev, stream=torch.cuda.current_stream().cuda_stream)
                    _native.affine_z4(z4.data_ptr(), 0, shape, b.data_ptr(), shape, m, interp, _native.OOB_ZERO, device=dev,
                                      stream=torch.cuda.current_stream().cuda_stream)
                    torch.cuda.synchronize()
                    if not torch.equal(a, b):
                        print('not bit-identical to the slice family', shape, angle, name, float((a - b).abs().max()))
    # prefilter writing Z4 directly against prefilter + pack
    for shape in ((40, 44, 48), (130, 50, 37), (64, 64, 64), (250, 30, 250)):
        src = torch.rand(shape, device=f'cuda:{dev}')
        plain = torch.empty(shape, device=src.device)
        _native.prefilter(src.data_ptr(), shape, dev, torch.cuda.current_stream().cuda_stream, dst_ptr=plain.data_ptr())
        want = z4_of(plain, 0)
        got = torch.full_like(want, 3.0)
        ws = torch.empty(shape, device=src.device)
        _native.prefilter_z4(src.data_ptr(), shape, got.data_ptr(), ws.data_ptr(), ws.numel() * 4, dev,
                             torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        e = float((got - want).abs().max()) / float(plain.max() - plain.min())
        worst = max(worst, e)
        if e > 1e-6:
            print('MISMATCH prefilter_z4', shape, e)
            return False
    print('parity ok, worst', worst)
    return True


def timing(n, quick):
    shape = (n, n, n)
    src = torch.rand(shape, device=f'cuda:{dev}')
    dst = torch.zeros(shape, device=f'cuda:{dev}')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{dev}')
    st = torch.cuda.current_stream().cuda_stream

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

    for axis in ((0, 2) if quick else (0, 1, 2)):
        z4 = z4_of(src, axis)
        for name, interp in MODES.items():
            row = []
            for angle in ((0, 45) if quick else (0, 10, 25, 35, 45, 60, 90, 120)):
                m = mat_for(shape, axis, angle)
                plan = _native.z4_plan(shape, m, interp, axis)
                t_new = med(lambda: _native.affine_z4(z4.data_ptr(), axis, shape, dst.data_ptr(), shape, m, interp, 0, device=dev, stream=st))
                t_old = med(lambda: _native.affine(src.data_ptr(), shape, dst.data_ptr(), shape, m, interp, 0, device=dev, stream=st))
                row.append(f'{angle}:{n ** 3 / t_new / 1e6:.0f}({n ** 3 / t_old / 1e6:.0f};wf{plan["wavefronts"][0]:.2f},s{plan["shapes"][0]})')
            print(f'{n}^3 axis {axis} {name:13s} Gvox/s z4(plain) ' + ' '.join(row), flush=True)
    # pack and prefilter
    z4 = z4_of(src, 0)
    ws = torch.empty(shape, device=src.device)
    t_pack = med(lambda: _native.pack_z4(src.data_ptr(), shape, z4.data_ptr(), 0, device=dev, stream=st))
    t_pack2 = med(lambda: _native.pack_z4(src.data_ptr(), shape, z4.data_ptr(), 2, device=dev, stream=st))
    t_pf4 = med(lambda: _native.prefilter_z4(src.data_ptr(), shape, z4.data_ptr(), ws.data_ptr(), ws.numel() * 4, dev, st))
    t_pf = med(lambda: _native.prefilter(src.data_ptr(), shape, dev, st, dst_ptr=dst.data_ptr()))
    print(f'{n}^3 pack axis0 {t_pack:.3f} ms, pack axis2 {t_pack2:.3f} ms, prefilter->z4 {t_pf4:.3f} ms, prefilter plain {t_pf:.3f} ms')
    # batched sweep: 32 matrices per launch, like StaticVolume.affine_many
    if n <= 256:
        mats = np.stack([mat_for(shape, 0, a) for a in range(0, 180, 6)])
        out = torch.empty((len(mats),) + shape, device=src.device)
        for name, interp in MODES.items():
            t_new = med(lambda: _native.affine_z4(z4.data_ptr(), 0, shape, out.data_ptr(), shape, mats, interp, 1, device=dev, stream=st), 5)
            t_old = med(lambda: _native.affine(src.data_ptr(), shape, out.data_ptr(), shape, mats, interp, 1, device=dev, stream=st), 5)
            print(f'{n}^3 sweep of {len(mats)} angles {name:13s}: z4 {len(mats) * n ** 3 / t_new / 1e6:.0f} Gvox/s, plain {len(mats) * n ** 3 / t_old / 1e6:.0f} Gvox/s')


if __name__ == '__main__':
    quick = '--quick' in sys.argv
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [256, 512]
    if '--no-parity' in sys.argv or parity():
        for n in sizes:
            timing(n, quick)
