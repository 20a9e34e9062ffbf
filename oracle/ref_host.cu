/*
 * ref_host.cu -- exposes the reference's own `__host__ __device__` prefilter / B-spline code on the CPU.
 * TEST INFRASTRUCTURE ONLY (oracle).  The two headers are included from where they lie under
 * /root/reference/voltools/kernels (build_ref.py passes -I); nothing is copied.  Compiled for the HOST
 * by nvcc into oracle/_ref/libvt_ref_host.so, so the CPU test-suite can pin oracle/vt_oracle.c against
 * the reference's real code without a GPU.
 */
#include "helper_math.h"
#include "bspline.h"

extern "C" void ref_host_prefilter_line(float *c, unsigned n, int step_bytes)
{
    ConvertToInterpolationCoefficients(c, n, step_bytes); /* bspline.h:30-54 */
}

extern "C" float ref_host_bspline(float t) { return bspline(t); } /* bspline.h:114-122 */
