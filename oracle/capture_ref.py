"""Capture the reference's own CUDA source for its `transform` / prefilter kernels.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Runs only where `/root/reference` exists
(the build container).  Nothing from the reference is copied into git history: the captured
strings are written under the git-ignored `oracle/_ref/`.

The reference assembles its kernels as Python f-strings and hands them to `cupy.RawKernel`
(voltools/transforms.py:232-287 and :290-299).  cupy is not installed here, so a stub `cupy`
module records the `code=` argument of every `RawKernel(...)` instead of compiling it.
"""
import sys
import types
from pathlib import Path

REF = Path('/root/reference')
OUT = Path(__file__).resolve().parent / '_ref'


def _fake_cupy(captured):
    cp = types.ModuleType('cupy')

    class RawKernel:  # records what the reference would have JIT-compiled
        def __init__(self, code=None, name=None, options=()):
            captured.append((name, code, tuple(options)))

    class _Runtime:
        @staticmethod
        def getDeviceCount():
            return 0

    cp.RawKernel = RawKernel
    cp.cuda = types.SimpleNamespace(runtime=_Runtime())
    return cp


def capture():
    """Returns {kernel_file_stem: source} for the 3 transform bodies + the prefilter module."""
    if not REF.exists():
        raise RuntimeError('/root/reference is not present; the reference oracle can only be built in the '
                           'build container (prebuilt oracle/_ref travels to the GPU box)')
    captured = []
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == 'cupy' or k.startswith('voltools')}
    for k in saved:
        sys.modules.pop(k, None)
    sys.modules['cupy'] = _fake_cupy(captured)
    sys.path.insert(0, str(REF))
    try:
        import voltools.transforms as rt  # the reference, unmodified
        out = {}
        for interp, stem in (('linear', 'transform_linear'), ('bspline', 'transform_cubic'),
                             ('bspline_simple', 'transform_cubic_simple')):
            del captured[:]
            rt._get_transform_kernel(interp)
            (name, code, _), = captured
            assert name == 'transform'
            out[stem] = code
        # prefilter module: volume arg is only touched after the three RawKernel() calls; a dummy
        # object that fails on `.strides` lets us record the sources without any device.
        del captured[:]
        try:
            rt._bspline_prefilter(object())
        except AttributeError:
            pass
        names = [c[0] for c in captured]
        assert names == ['SamplesToCoefficients3DX', 'SamplesToCoefficients3DY', 'SamplesToCoefficients3DZ'], names
        out['prefilter'] = captured[0][1]
        return out
    finally:
        sys.path.remove(str(REF))
        for k in [k for k in sys.modules if k == 'cupy' or k.startswith('voltools')]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


if __name__ == '__main__':
    OUT.mkdir(exist_ok=True)
    for stem, code in capture().items():
        (OUT / f'{stem}.cu').write_text(code)
        print('captured', stem, len(code), 'chars')
