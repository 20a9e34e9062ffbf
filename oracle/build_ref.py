"""Build the reference-backed oracles into the git-ignored oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

  oracle/_ref/transform_{linear,cubic,cubic_simple}.cubin, prefilter.cubin
        the reference's own kernel sources (captured from voltools/transforms.py by capture_ref.py),
        compiled unmodified for sm_100a with nvcc against /root/reference/voltools/kernels
  oracle/_ref/libvt_ref_gpu.so    oracle/ref_harness.cu: replays the reference's CuPy host calls around them
  oracle/_ref/libvt_ref_host.so   oracle/ref_host.cu: the reference's __host__ prefilter code on the CPU
  oracle/_ref/py/voltools         the unmodified reference package (pip --target), for the CPU baseline
                                  (`device='cpu'` -> scipy.ndimage.affine_transform)

Only runs where /root/reference exists; the GPU box uses the prebuilt files (they travel with gpurun).
"""
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / '_ref'
REF = Path('/root/reference')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def run(cmd):
    print('+', ' '.join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True)


def build(force=False):
    if not REF.exists():
        print('oracle/build_ref.py: /root/reference absent; keeping prebuilt oracle/_ref')
        return False
    OUT.mkdir(exist_ok=True)
    sys.path.insert(0, str(HERE))
    import capture_ref
    inc = REF / 'voltools' / 'kernels'
    for stem, code in capture_ref.capture().items():
        src = OUT / f'{stem}.cu'
        if not src.exists() or src.read_text() != code:
            src.write_text(code)
        cubin = OUT / f'{stem}.cubin'
        if force or not cubin.exists() or cubin.stat().st_mtime < src.stat().st_mtime:
            # same options the reference passes to RawKernel: just the include path (no fast-math)
            run(['nvcc', *ARCH, '-I', inc, '-cubin', '-o', cubin, src])
    so = OUT / 'libvt_ref_gpu.so'
    src = HERE / 'ref_harness.cu'
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        run(['nvcc', *ARCH, '-O2', '-shared', '-Xcompiler', '-fPIC', '-o', so, src, '-lcuda'])
    so = OUT / 'libvt_ref_host.so'
    src = HERE / 'ref_host.cu'
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        run(['nvcc', '-O2', '-shared', '-Xcompiler', '-fPIC', '-I', inc, '-o', so, src])
    py = OUT / 'py'
    if force or not (py / 'voltools' / '__init__.py').exists():
        tmp = Path('/tmp/_voltools_ref_src')
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(REF, tmp)  # pip wants to write egg-info next to setup.py; /root/reference is read-only
        shutil.rmtree(py, ignore_errors=True)
        run([sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '-q',
             '--find-links', '/opt/wheelhouse', '--target', py, tmp])
        shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == '__main__':
    build(force='--force' in sys.argv)
