"""Compile voltools_b200/csrc/*.cu into voltools_b200/libvoltools_b200.so for sm_100a (in-tree)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
SO = PKG / 'libvoltools_b200.so'
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _compile(src, verbose):
    obj = HERE / (src.stem + '.o')
    deps = [src, HERE / 'vt_common.cuh', PKG.parent / 'include' / 'voltools_b200.h']
    if obj.exists() and all(obj.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return obj, ''
    extra = os.environ.get('VT_NVCC_EXTRA', '').split()  # tuning experiments: extra -D flags
    r = subprocess.run(['nvcc', *NVCC_FLAGS, *extra, '-c', '-o', str(obj), str(src)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}')
    return obj, r.stderr


def build(verbose=False, force=False):
    srcs = sorted(HERE.glob('*.cu'))
    if force:
        for o in HERE.glob('*.o'):
            o.unlink()
    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or not SO.exists() or any(o.stat().st_mtime > SO.stat().st_mtime for o in objs):
        r = subprocess.run(['nvcc', '-shared', '-o', str(SO), *[str(o) for o in objs]],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return SO


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='--force' in sys.argv))
