// vt_common.cuh -- shared device helpers for the resampling kernels (sm_100a).
//
// Everything numerical in here mirrors, operation for operation, what the reference's kernels compile to
// (SASS of its `transform` kernels, see DESIGN.md "float recipe"); explicit _rn intrinsics pin the rounding
// sequence so the compiler cannot contract or reassociate it differently.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
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
    long long src_row, src_plane;  // element strides of the source: row (axis 1) and plane (axis 0); dense = s2, s1*s2
    int o0, o1, o2;
    long long dst_batch_stride;
    int z_begin, z_end;
    int n_mats;
    unsigned flags;
    unsigned char aux[VT_MAX_BATCH];  // per-matrix launch parameter chosen by the family's host code
    VtMat mats[VT_MAX_BATCH];
};

// internal flag (not part of the C ABI's flag set): the general-matrix kernels ADD each thread's voxels along axis 0
// into a projection image (P.dst = n_mats images of o1 x o2, zeroed by the caller) instead of storing them
#define VT_INTERNAL_PROJECT 0x1000u

void vt_count_launch(int n = 1);

// 3-D float32 tensor map (dims/strides fastest axis first, strides in bytes for axes 1 and 2), no swizzle, zeros out
// of bounds.  `tmap` points to a CUtensorMap.  The driver entry point is resolved at run time so that the library
// has no link-time dependency on libcuda (it must load on machines without a driver).
int vt_encode_tmap_3d(void *tmap, const void *base, const unsigned long long dims[3], const unsigned long long strides[2],
                      const unsigned box[3]);
// the same for rank 3..5 (dims / box: rank entries, strides: rank - 1 entries in bytes)
int vt_encode_tmap_nd(void *tmap, const void *base, int rank, const unsigned long long *dims, const unsigned long long *strides,
                      const unsigned *box);

// ---------------------------------------------------------------------------------------------------
// per-kernel device timing (vt_profile_* in the C ABI): every launch site wraps its <<<>>> in a VtProf,
// which brackets it with CUDA events on the launch stream while profiling is enabled.
// ---------------------------------------------------------------------------------------------------
enum VtKernelId {
    VT_K_PREFILTER_X = 0,
    VT_K_PREFILTER_Y,
    VT_K_PREFILTER_Z,
    VT_K_PREFILTER_FUSED,
    VT_K_GATHER_LINEAR,
    VT_K_GATHER_CUBIC_TEX,
    VT_K_GATHER_CUBIC_SIMPLE,
    VT_K_BRICK_LINEAR,
    VT_K_BRICK_CUBIC_TEX,
    VT_K_BRICK_CUBIC_SIMPLE,
    VT_K_SLICE_LINEAR,
    VT_K_SLICE_CUBIC_TEX,
    VT_K_SLICE_CUBIC_SIMPLE,
    VT_K_TEX_LINEAR,
    VT_K_TEX_CUBIC,
    VT_K_PLANE_SUM,
    VT_K_PROJECT_2D,
    VT_K_Z4_LINEAR,
    VT_K_Z4_CUBIC_TEX,
    VT_K_Z4_CUBIC_SIMPLE,
    VT_K_PACK_Z4,
    VT_K_PAD_ROWS,
    VT_K_COUNT
};
// ---------------------------------------------------------------------------------------------------
// Programmatic dependent launch: the kernels of a one-shot call (xy prefilter -> z prefilter -> pack -> slice4) are each
// 30-70 us at the reference's benchmark sizes, so the launch gap between them is a visible share.  Launched with
// cudaLaunchAttributeProgrammaticStreamSerialization a kernel's CTAs become resident while the previous kernel of the
// stream drains; vt_pdl_wait() at the very top of the kernel then blocks until that kernel has completed and its
// writes are visible -- stream order as before, minus the launch latency.  VT_PDL=0 turns the attribute off.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vt_pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool vt_pdl_enabled();  // vt_api.cu
template <typename... KArgs, typename... Args>
inline cudaError_t vt_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = vt_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct VtProf {
    int id;
    cudaStream_t st;
    void *rec;
    VtProf(int id, cudaStream_t st);
    ~VtProf();
};

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
// texture-unit model (tex3D<float>, cudaFilterModeLinear, border, unnormalised) of the B200, measured with
// oracle/probe_tex.py and pinned by tests/golden/tex_probe_b200.npz:
//   per axis  X = floor(x*256 + 0.5) - 128;  base texel = X >> 8;  alpha = X & 255   (1/256ths)
//   the eight texel weights are integers in 1/256ths that sum to 256 (a, b, c = alphas along x, y, z):
//     for each z side S in {256-c, c}:  XF = (a*S+128)>>8, XN = S-XF,
//        W(xfar,yfar) = (b*XF+128)>>8, W(xfar,ynear) = XF - W(xfar,yfar),
//        W(xnear,ynear) = ((256-b)*XN+128)>>8, W(xnear,yfar) = XN - W(xnear,ynear)
// RULE 0 = this hardware model (parity), RULE 2 = exact float32 fractions (VT_WEIGHTS_EXACT).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vt_tex_fix_hw(float x, int &base, int &alpha)
{
    const int X = __float2int_rd(__fmaf_rn(x, 256.0f, 0.5f)) - 128;
    base = X >> 8;
    alpha = X & 255;
}

// small non-negative integer (< 2^23) -> float without the conversion pipe
__device__ __forceinline__ float vt_u2f(int v) { return __int_as_float(0x4B000000 | v) - 8388608.0f; }

// one z side of the weight rule: S = weight mass of the side; w[0..3] = (xn,yn), (xf,yn), (xn,yf), (xf,yf)
__device__ __forceinline__ void vt_tex_hw_side(int a, int b, int S, int w[4])
{
    const int XF = (a * S + 128) >> 8, XN = S - XF;
    const int ff = (b * XF + 128) >> 8;
    const int nn = ((256 - b) * XN + 128) >> 8;
    w[0] = nn;
    w[1] = XF - ff;
    w[2] = XN - nn;
    w[3] = ff;
}

// all eight weights of a fetch: w[zs*4 + ys*2 + xs], side 0 = near (lower index) texel, 1 = far
__device__ __forceinline__ void vt_tex_hw8(int a, int b, int c, int w[8])
{
    int lo[4], hi[4];
    vt_tex_hw_side(a, b, 256 - c, lo);
    vt_tex_hw_side(a, b, c, hi);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        w[k] = lo[k];
        w[4 + k] = hi[k];
    }
}
// The weight rule is not symmetric in the axes, so kernels that fix the fraction along one volume axis m ("march
// axis": its alpha is mu, its side ms) and vary the two others must put each alpha in its own slot: texture x is
// volume axis 2, y axis 1, z axis 0.  Returns the weight of (row side rs, column side cs) of the in-plane footprint,
// rows / columns being the two other axes in ascending order.
__device__ __forceinline__ int vt_tex_hw_inplane(int m, int a_col, int a_row, int mu, int ms, int rs, int cs)
{
    int w[8];
    if (m == 0) {
        vt_tex_hw8(a_col, a_row, mu, w);   // rows = axis 1 (y), columns = axis 2 (x), march = z
        return w[ms * 4 + rs * 2 + cs];
    }
    if (m == 1) {
        vt_tex_hw8(a_col, mu, a_row, w);   // rows = axis 0 (z), columns = axis 2 (x), march = y
        return w[rs * 4 + ms * 2 + cs];
    }
    vt_tex_hw8(mu, a_col, a_row, w);       // rows = axis 0 (z), columns = axis 1 (y), march = x
    return w[rs * 4 + cs * 2 + ms];
}

template <int RULE>
__device__ __forceinline__ void vt_tex_fix(float x, int &i, float &alpha)
{
    static_assert(RULE == 2, "only the exact rule uses float alphas");
    const float xb = __fadd_rn(x, -0.5f);
    const float fl = floorf(xb);
    i = (int)fl;
    alpha = __fsub_rn(xb, fl);
}

// ---------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) + mbarrier wrappers.  Waits are bounded: a lost completion traps instead of hanging.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned vt_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void vt_mbar_init(unsigned bar, unsigned arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void vt_mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void vt_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void vt_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool vt_mbar_try_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void vt_mbar_wait(unsigned bar, unsigned parity)
{
    unsigned spins = 0;
    while (!vt_mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();  // a TMA that never completes must not hang the GPU
}
// 3-D tiled box load: coordinates (c0 fastest) may lie outside the tensor, those elements arrive as zeros
__device__ __forceinline__ void vt_tma_load_3d(unsigned smem_dst, const void *tmap, unsigned bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::
            "r"(smem_dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy): evict_last keeps a volume that every matrix of a launch
// re-reads resident while the outputs stream through the cache
__device__ __forceinline__ unsigned long long vt_l2_policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void vt_tma_load_3d_hint(unsigned smem_dst, const void *tmap, unsigned bar, int c0, int c1, int c2,
                                                    unsigned long long policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::
            "r"(smem_dst),
        "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
// 4-D tiled box STORE shared -> global (bulk async group); elements outside the tensor are not written
__device__ __forceinline__ void vt_tma_store_4d(const void *tmap, unsigned smem_src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(tmap),
                 "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void vt_bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void vt_bulk_wait_read()  // at most N groups still READING shared memory
{
    asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void vt_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void vt_tma_prefetch_desc(const void *tmap)
{
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(tmap) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// cp.async (LDGSTS), 4-byte form.  `take == 0` writes a zero instead (ignore-src form: no global access is made).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void vt_cp_async4(unsigned smem_dst, const void *gmem_src, unsigned take)
{
    asm volatile(
        "{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n"
        " cp.async.ca.shared.global [%0], [%1], 4, p;\n}\n" ::"r"(smem_dst),
        "l"(gmem_src), "r"(take)
        : "memory");
}
__device__ __forceinline__ void vt_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void vt_cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// packed float32 pairs: sm_100 issues two IEEE fp32 FMAs per lane with one FFMA2 (fma.rn.f32x2); ptxas folds a
// pair built from the same scalar, vt_pk(t, t), into the instruction's broadcast operand form.  Each half is an
// ordinary fma.rn.f32, so results are bit-identical to the scalar code.
// ---------------------------------------------------------------------------------------------------
typedef unsigned long long vt_f2;
__device__ __forceinline__ vt_f2 vt_pk(float lo, float hi)
{
    vt_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void vt_unpk(vt_f2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ vt_f2 vt_fma2(vt_f2 a, vt_f2 b, vt_f2 c)
{
    vt_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ vt_f2 vt_fma2_rm(vt_f2 a, vt_f2 b, vt_f2 c)
{
    vt_f2 r;
    asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ vt_f2 vt_add2(vt_f2 a, vt_f2 b)
{
    vt_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ vt_f2 vt_add2_rm(vt_f2 a, vt_f2 b)
{
    vt_f2 r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ vt_f2 vt_mul2(vt_f2 a, vt_f2 b)
{
    vt_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ vt_f2 vt_bc(float x) { return vt_pk(x, x); }
__device__ __forceinline__ float vt_lo(vt_f2 v) { float a, b; vt_unpk(v, a, b); return a; }
__device__ __forceinline__ float vt_hi(vt_f2 v) { float a, b; vt_unpk(v, a, b); return b; }

// ---------------------------------------------------------------------------------------------------
// The texture-unit model again, entirely on the float pipes and two values per instruction.  Every quantity of the
// rule is a small integer (< 2^17), so float32 holds it exactly; the only inexact steps of the integer code are its
// floors, reproduced with round-toward-minus-infinity adds against VT_MAGIC = 1.5 * 2^23 (for |n| < 2^22 the sum
// MAGIC + n is exact at unit spacing, so RM(x + MAGIC) - MAGIC == floor(x)).  Same weights as vt_tex_fix_hw /
// vt_tex_hw_side bit for bit -- no F2I, no integer multiplies, no int->float conversions: the general-matrix kernels
// were bound by exactly those (154 / 573 instructions per voxel).
// ---------------------------------------------------------------------------------------------------
#define VT_MAGIC 12582912.0f
#define VT_MAGIC_BITS 0x4B400000
// two coordinates at once: alpha (float, 0..255) and base texel (int) of each
__device__ __forceinline__ void vt_tex_fix2(vt_f2 x, vt_f2 &alpha, int &base_lo, int &base_hi)
{
    const vt_f2 u = vt_fma2(x, vt_bc(256.0f), vt_bc(0.5f));          // the same fma as vt_tex_fix_hw
    const vt_f2 t = vt_add2_rm(u, vt_bc(VT_MAGIC));                   // MAGIC + floor(u)
    const vt_f2 X = vt_add2(t, vt_bc(-(VT_MAGIC + 128.0f)));          // floor(u) - 128, exact
    const vt_f2 bm = vt_fma2_rm(X, vt_bc(1.0f / 256.0f), vt_bc(VT_MAGIC));  // MAGIC + floor(X / 256)
    const vt_f2 base = vt_add2(bm, vt_bc(-VT_MAGIC));
    alpha = vt_fma2(base, vt_bc(-256.0f), X);                         // X - 256 * base, exact
    base_lo = __float_as_int(vt_lo(bm)) - VT_MAGIC_BITS;
    base_hi = __float_as_int(vt_hi(bm)) - VT_MAGIC_BITS;
}
// floor((v + 128) / 256) for integer-valued v, both halves
__device__ __forceinline__ vt_f2 vt_round8_2(vt_f2 v_plus_128)
{
    return vt_add2(vt_fma2_rm(v_plus_128, vt_bc(1.0f / 256.0f), vt_bc(VT_MAGIC)), vt_bc(-VT_MAGIC));
}
// the eight texel weights (1/256ths) of one fetch, as four {z-near, z-far} pairs; a, b, c = alphas along x, y, z
struct VtTexW {
    vt_f2 nn, fn, nf, ff;  // (x near, y near), (x far, y near), (x near, y far), (x far, y far)
};
__device__ __forceinline__ VtTexW vt_tex_weights2(float a, float b, float c)
{
    const vt_f2 S = vt_fma2(vt_bc(c), vt_pk(-1.0f, 1.0f), vt_pk(256.0f, 0.0f));  // {256 - c, c}
    const vt_f2 XF = vt_round8_2(vt_fma2(vt_bc(a), S, vt_bc(128.0f)));
    const vt_f2 XN = vt_fma2(XF, vt_bc(-1.0f), S);  // S - XF, exact
    VtTexW w;
    w.ff = vt_round8_2(vt_fma2(vt_bc(b), XF, vt_bc(128.0f)));
    w.nn = vt_round8_2(vt_fma2(vt_bc(256.0f - b), XN, vt_bc(128.0f)));
    w.fn = vt_fma2(w.ff, vt_bc(-1.0f), XF);  // XF - ff, exact
    w.nf = vt_fma2(w.nn, vt_bc(-1.0f), XN);  // XN - nn, exact
    return w;
}

// ---------------------------------------------------------------------------------------------------
// cubic B-spline pieces
// ---------------------------------------------------------------------------------------------------
// x / 6 correctly rounded (== __fdiv_rn(x, 6.0f)) in three instructions: q = RN(x * RN(1/6)), then one Newton step
// on the exact residual.  Checked exhaustively against IEEE division for every float x in [2^-125, 8] and x = 0
// (1.09e9 values, tools/check_div6.c); it differs only where the quotient is denormal, which a^3 with a a multiple
// of 2^-23 in [0, 1] never is.
__device__ __forceinline__ float vt_div6(float x)
{
    const float r6 = 1.0f / 6.0f;
    const float q = __fmul_rn(x, r6);
    return __fmaf_rn(__fmaf_rn(-6.0f, q, x), r6, q);
}

// bspline() of voltools/kernels/bspline.h:114-122, in its compiled operation order; branch-free (both pieces are
// evaluated and selected: a dozen instructions, no divergence, no division subroutine)
__device__ __forceinline__ float vt_bspline(float t)
{
    t = fabsf(t);
    const float a = __fsub_rn(2.0f, t);
    const float p = __fmaf_rn(a, __fmul_rn(__fmul_rn(t, -0.5f), t), 2.0f / 3.0f);
    const float d = vt_div6(__fmul_rn(__fmul_rn(a, a), a));
    return t < 1.0f ? p : (t < 2.0f ? d : 0.0f);
}
// the same for an argument known to lie in [0, 1] (taps 0 and 1 of a cubic footprint: |0 - f|, |1 - f|) ...
__device__ __forceinline__ float vt_bspline_inner(float t)
{
    const float a = __fsub_rn(2.0f, t);
    const float p = __fmaf_rn(a, __fmul_rn(__fmul_rn(t, -0.5f), t), 2.0f / 3.0f);
    return t < 1.0f ? p : 1.0f / 6.0f;  // t == 1: a = 1, a^3 / 6
}
// ... and in [1, 2] (taps -1 and 2: |-1 - f|, |2 - f|)
__device__ __forceinline__ float vt_bspline_outer(float t)
{
    const float a = __fsub_rn(2.0f, t);
    const float d = vt_div6(__fmul_rn(__fmul_rn(a, a), a));
    return t < 2.0f ? d : 0.0f;
}
// the four weights of a footprint for fraction f in [0, 1]: bspline(k - f), k = -1, 0, 1, 2
__device__ __forceinline__ void vt_bspline4(float f, float (&w)[4])
{
    w[0] = vt_bspline_outer(fabsf(__fsub_rn(-1.0f, f)));
    w[1] = vt_bspline_inner(fabsf(__fsub_rn(0.0f, f)));
    w[2] = vt_bspline_inner(fabsf(__fsub_rn(1.0f, f)));
    w[3] = vt_bspline_outer(fabsf(__fsub_rn(2.0f, f)));
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
