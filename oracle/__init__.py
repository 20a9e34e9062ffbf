"""oracle/ -- CPU restatement of the reference's GPU resampling path + reference-backed checkers.

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs, never by voltools_b200/ (the product fails loudly without its CUDA library).
"""
from .oracle import (  # noqa: F401
    INTERP_FN, TEX_RN, TEX_TRUNC, TEX_EXACT, TEX_HW, build, affine, prefilter, prefilter_line, tex3d_many, transform_ref_gpu,
    prefilter_ref_gpu, tex3d_ref_gpu, ref_gpu_available, ref_host_available, ref_host_prefilter_line,
    ref_host_bspline, ref_python_path, constants,
)
