// vt_common.cuh -- shared device helpers for the resampling kernels (sm_100a).
//
// Everything numerical in here mirrors, operation for operation, what the reference's kernels compile to
// (SASS of its `transform` kernels, see DESIGN.md "float recipe"); explicit _rn intrinsics pin the rounding
// sequence so the compiler cannot contract or reassociate it differently.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/voltools_b200.h"

#define VT_CUDA(x)                                  \
    do {                                            \
        cudaError_t e_ = (x);                       \
        if (e_ != cudaSuccess) return 1000 + (int)e_; \
    } while (0)

struct VtMat {
    float r[3][4];  // rows 0..2 of the row-major 4x4 (output index -> input index)
};

struct VtResampleParams {
    const float *src;
    float *dst;
    int s0, s1, s2;
    int o0, o1, o2;
    long long dst_batch_stride;
    int z_begin, z_end;
    int n_mats;
    unsigned flags;
    VtMat mats[VT_MAX_BATCH];
};

void vt_count_launch(int n = 1);

// ---------------------------------------------------------------------------------------------------
// coordinate recipe: voltools/transforms.py:264-274 as compiled:
//   t = a1*M[r][1]; t = fma(a0, M[r][0], t); t = fma(a2, M[r][2], t); t = M[r][3] + t; p = t + 0.5
// split so that the (a0, a1)-only part can be hoisted out of a run along a2.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float vt_row_base(const float *row, float a0, float a1)
{
    return __fmaf_rn(a0, row[0], __fmul_rn(a1, row[1]));
}
__device__ __forceinline__ float vt_row_finish(const float *row, float base, float a2)
{
    return __fadd_rn(__fadd_rn(row[3], __fmaf_rn(a2, row[2], base)), 0.5f);
}

// ---------------------------------------------------------------------------------------------------
// texture-unit emulation: unnormalised coordinate -> (base texel, alpha), alpha with 8 fractional bits.
// RULE 0: round-to-nearest conversion to 1.8 fixed point, 1: truncating, 2: exact float32 fraction.
// ---------------------------------------------------------------------------------------------------
template <int RULE>
__device__ __forceinline__ void vt_tex_fix(float x, int &i, float &alpha)
{
    if (RULE == 2) {
        const float xb = __fadd_rn(x, -0.5f);
        const float fl = floorf(xb);
        i = (int)fl;
        alpha = __fsub_rn(xb, fl);
    } else {
        const float s = __fmul_rn(x, 256.0f);
        int X = (RULE == 0) ? __float2int_rn(s) : __float2int_rd(s);
        X -= 128;
        i = X >> 8;
        alpha = (float)(X & 255) * (1.0f / 256.0f);
    }
}

// ---------------------------------------------------------------------------------------------------
// cubic B-spline pieces
// ---------------------------------------------------------------------------------------------------
// bspline() of voltools/kernels/bspline.h:114-122, in its compiled operation order
__device__ __forceinline__ float vt_bspline(float t)
{
    t = fabsf(t);
    const float a = __fsub_rn(2.0f, t);
    if (t < 1.0f) return __fmaf_rn(a, __fmul_rn(__fmul_rn(t, -0.5f), t), 2.0f / 3.0f);
    if (t < 2.0f) return __fdiv_rn(__fmul_rn(__fmul_rn(a, a), a), 6.0f);
    return 0.0f;
}

// bspline_weights() + g0/g1/h0/h1 of helper_interpolation.h:11-20 / bspline.h:102-112, one axis.
// h0/h1 must be bit-identical to the reference: they are quantised to 1/256 by the texture unit, so a
// one-ulp difference can move a weight by a whole quantisation step.
__device__ __forceinline__ void vt_ruijters(float coord, float &g0, float &g1, float &h0, float &h1)
{
    const float cg = __fadd_rn(coord, -0.5f);
    const float idx = floorf(cg);
    const float f = __fsub_rn(cg, idx);
    const float o = __fsub_rn(1.0f, f);
    const float sq = __fmul_rn(f, f), osq = __fmul_rn(o, o);
    const float sixth = 1.0f / 6.0f, twothirds = 2.0f / 3.0f;
    const float w1 = __fmaf_rn(__fmul_rn(sq, -0.5f), __fsub_rn(2.0f, f), twothirds);
    const float w2 = __fmaf_rn(__fmul_rn(osq, -0.5f), __fsub_rn(2.0f, o), twothirds);
    const float w3 = __fmul_rn(f, __fmul_rn(sq, sixth));
    g0 = __fmaf_rn(o, __fmul_rn(osq, sixth), w1);
    g1 = __fadd_rn(w2, w3);
    h0 = __fadd_rn(idx, __fadd_rn(__fdiv_rn(w1, g0), -0.5f));
    h1 = __fadd_rn(idx, __fadd_rn(__fdiv_rn(w3, g1), 1.5f));
}
