// vt_resample_z4.cu -- "slice4" kernel family: transforms that leave ONE axis alone, on a coefficient volume whose
// untouched axis is interleaved by four.
//
// When the matrix maps output axis m straight onto input axis m with an integer offset
//        M[m] = e_m + (0,0,0,t),  t integer,  M[r][m] = 0 for r != m
// (every rotation about axis m through the centre of the volume: BASELINE configs[0..2] and the README sweep for
// m = 0; `rotation=(25,0,0)` 'rzxz' of examples/transformation.py for m = 2), the in-plane position of an output
// column and therefore every in-plane interpolation weight is the same along the whole march over axis m, and the
// fraction along axis m is exactly 0.  vt_resample_slice.cu exploits that for m = 0 on the plain layout and ends up
// bound by shared-memory wavefronts: 16 scalar LDS per voxel and plane, ~2 wavefronts each at 45 degrees because a
// warp's 32 texels cannot be spread over 32 banks under a rotation.  This family changes the LAYOUT instead:
//
//   "Z4" layout of axis m:   L[g][y][x][j] = V[index 4g+j along axis m][y][x]      (j = 0..3, zero past the end)
//
// (y, x) being the other two axes in ascending order.  One TMA box then delivers, per texel of a tile's in-plane
// footprint, FOUR consecutive planes in 16 contiguous bytes, and a tap becomes one LDS.128 that serves four output
// planes:
//   * 4 instead of 16 LDS per voxel and plane, and the four planes of a load pair up naturally in FFMA2;
//   * bank conflicts are evaluated per QUARTER warp (8 lanes x 16 bytes = one 128-byte wavefront), so 8 texels
//     have to fall into 8 distinct 16-byte bank groups instead of 32 texels into 32 banks: the host picks, per
//     matrix, how 8 lanes tile the columns (1x8, 2x4, 4x2, 8x1) and the row pitch mod 8 (one of eight tensor maps
//     whose box widths differ by one texel) from a simulation -- 1.0-1.5 wavefronts per quarter (mean 1.19 over a
//     180-angle sweep) instead of 1.66-1.97;
//   * rows of 16-byte texels are always 16-byte aligned: no row padding, no rounding of the box start.
// The same kernel marches along axis 0, 1 or 2: only the output strides and the roles of the two in-plane indices in
// the reference's float32 coordinate recipe change (`prod_on_fast`).
//
// The layout is produced by vt_pack_z4 (any axis, from the plain layout) or directly by the prefilter's Z pass
// (axis 0).  Results are bit-identical to vt_resample_slice.cu (same weights, same summation order).
//
// Replaces the reference's `transform` kernel (voltools/transforms.py:253-282) + linearTex3D / cubicTex3D /
// cubicTex3DSimple (voltools/kernels/helper_interpolation.h:3-68) for this class of matrices.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "vt_common.cuh"

namespace {

constexpr int TS = 16;        // tile edge
constexpr int NT = TS * TS;   // threads per CTA, one column each
constexpr int BH_MAX = 32;    // max footprint rows (texels)
constexpr int BW_MAX = 40;    // max box width (texels): footprint <= 33 + 7 pitch classes
constexpr int STAGE_BYTES_MAX = BH_MAX * BW_MAX * 16;  // one ring stage: a footprint of four planes (<= 20 KB)
constexpr int SMEM_HEADER = 128;  // the ring stages' "full" mbarriers
constexpr int NPITCH = 8;     // tensor maps per launch: box widths bw0 .. bw0+7 (pitch mod 8 is what matters)
constexpr int N_SHAPES = 4;   // quarter-warp shapes 1x8, 2x4, 4x2, 8x1 (rows x columns of the tile)

struct Z4Mat {
    float ys, yf, yc;  // source y (slower in-plane axis) = f(slow index, fast index): coefficients, constant
    float xs, xf, xc;  // source x (contiguous in-plane axis)
    int tm;            // integer shift along the march axis
    unsigned char shape, pitch_idx, pad0, pad1;
};

struct Z4Params {
    float *dst;
    long long dst_batch_stride;
    long long os_slow, os_fast, os_m;  // output element strides of the tile's slow / fast axes and of the march axis
    int o_fast;                        // extent of the fast tile axis
    int o_slow, o_m;                   // full extents of the slow tile axis and of the march axis
    int slow_b, slow_e;                // range of the slow tile axis to produce
    int m_b, m_e;                      // range of the march axis to produce
    int s_y, s_x, s_m;                 // source extents: in-plane rows, in-plane columns, march axis
    int n_mats;
    int prod_on_fast;                  // which in-plane index the recipe multiplies first (see z4_coord)
    int bw0, bh;                       // box width of map 0 (map k: bw0 + k) and box height, texels
    int tiles_fast;
    int tsy;                           // rows of a tile: 16 (256-thread CTAs) or 8 (128-thread CTAs)
    int march;                         // the march axis (0..2): selects the slots of the texture-weight rule
    int mat_fastest;                   // block order: 1 = the matrix index varies fastest (blockIdx.x = tile * n_mats + mat)
    int vec_store;                     // march axis contiguous in the output and 16-byte alignable: STG.128 quads
    int tma_store;                     // ... and staged through shared memory + cp.async.bulk.tensor stores (UTMASTG)
    Z4Mat mats[VT_MAX_BATCH];
};

struct Z4Maps {
    CUtensorMap map[NPITCH];
    CUtensorMap omap;  // output tensor (x, a1, a0, matrix) with an 8 x 16 x 16 x 1 box: the TMA store path of march axis 2
};
constexpr int OTILE_BYTES = TS * TS * 32;  // one staged output tile (16-row tiles): two 16-byte quads = a 32-byte sector per column

// The reference's coordinate recipe (voltools/transforms.py:264-274 as compiled: t = a1*M1; t = fma(a0,M0,t);
// t = fma(a2,M2,t); t = M3 + t; p = t + 0.5) for a source row whose coefficient of the march index is zero: that term
// is an exact no-op, and what is left is one product and one fma over the two in-plane indices.  Which index is
// multiplied first depends on the axis numbers: march 0 -> a1*M1 then fma(a2); march 1 -> a0*M0 (fma(a0,M0,a1*0)
// rounds a0*M0 once, like the product) then fma(a2); march 2 -> a1*M1 then fma(a0).  With (slow, fast) = (a1,a2),
// (a0,a2), (a0,a1) that is "product on slow" for march 0 / 1 and "product on fast" for march 2.
__host__ __device__ __forceinline__ float z4_coord(float cs, float cf, float cc, float as, float af, int prod_on_fast)
{
#ifdef __CUDA_ARCH__
    const float t = prod_on_fast ? __fmaf_rn(as, cs, __fmul_rn(af, cf)) : __fmaf_rn(af, cf, __fmul_rn(as, cs));
    return __fadd_rn(__fadd_rn(cc, t), 0.5f);
#else
    const float t = prod_on_fast ? fmaf(as, cs, af * cf) : fmaf(af, cf, as * cs);
    return (cc + t) + 0.5f;
#endif
}

// thread -> column of the 16 x 16 tile: 8 consecutive lanes (a quarter warp = one LDS.128 wavefront) form a
// (1<<s) x (8>>s) patch; patches are laid out row-major, so that a warp's four patches cover 2x16, 2x16, 4x8, 8x4
__host__ __device__ __forceinline__ void z4_lane_pos(int tid, int s, int &ty, int &tx)
{
    const int q = tid >> 3, l = tid & 7;
    const int band = q >> (s + 1), qcol = q & ((2 << s) - 1);
    ty = (band << s) + (l >> (3 - s));
    tx = qcol * (8 >> s) + (l & ((8 >> s) - 1));
}

__device__ __forceinline__ void lds128(unsigned addr, vt_f2 &lo, vt_f2 &hi)
{
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr));
}

template <int INTERP>
struct Taps4;

// ---- linear: 4 taps with the texture unit's integer weights (fraction along the march axis 0 -> S = 256) ----
template <>
struct Taps4<VT_LINEAR> {
    static constexpr int LO = 0, HI = 2;  // footprint margins relative to floor(p - 0.5)
    static constexpr int BEFORE = 0, AFTER = 0;
    float w[4];
    unsigned r0, r1;  // byte offsets of the two tap rows inside a stage
    template <int RULE>
    __device__ __forceinline__ void init(float py, float px, int ylo, int xlo, int pitch, int m)
    {
        int by, bx;
        if (RULE == 0) {
            int a, b;
            vt_tex_fix_hw(px, bx, a);
            vt_tex_fix_hw(py, by, b);
#pragma unroll
            for (int k = 0; k < 4; k++) w[k] = vt_u2f(vt_tex_hw_inplane(m, a, b, 0, 0, k >> 1, k & 1)) * (1.0f / 256.0f);
        } else {
            float ax, ay;
            vt_tex_fix<2>(px, bx, ax);
            vt_tex_fix<2>(py, by, ay);
            w[0] = (1.0f - ax) * (1.0f - ay);
            w[1] = ax * (1.0f - ay);
            w[2] = (1.0f - ax) * ay;
            w[3] = ax * ay;
        }
        r0 = 16u * (unsigned)((by - ylo) * pitch + (bx - xlo));
        r1 = r0 + 16u * (unsigned)pitch;
    }
    __device__ __forceinline__ void planes(unsigned ring, float (&out)[4]) const
    {
        vt_f2 a01 = 0ull, a23 = 0ull;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            vt_f2 lo, hi;
            lds128(ring + (k < 2 ? r0 : r1) + 16u * (k & 1), lo, hi);
            const vt_f2 ww = vt_pk(w[k], w[k]);
            a01 = vt_fma2(lo, ww, a01);
            a23 = vt_fma2(hi, ww, a23);
        }
        vt_unpk(a01, out[0], out[1]);
        vt_unpk(a23, out[2], out[3]);
    }
};

// ---- cubic_simple: 16 in-plane taps, float32 B-spline weights; march-axis weights B(-1), B(0), B(1) ----------
template <>
struct Taps4<VT_CUBIC_SIMPLE> {
    static constexpr int LO = -1, HI = 2;
    static constexpr int BEFORE = 1, AFTER = 1;
    float w[16];
    unsigned row[4];
    float wz0, wz1, wz2;
    template <int RULE>
    __device__ __forceinline__ void init(float py, float px, int ylo, int xlo, int pitch, int)
    {
        const float cgx = __fadd_rn(px, -0.5f), cgy = __fadd_rn(py, -0.5f);
        const float fx0 = floorf(cgx), fy0 = floorf(cgy);
        const float fx = __fsub_rn(cgx, fx0), fy = __fsub_rn(cgy, fy0);
        float wx[4], wy[4];
        vt_bspline4(fx, wx);
        vt_bspline4(fy, wy);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) w[j * 4 + i] = __fmul_rn(wx[i], wy[j]);
        row[0] = 16u * (unsigned)(((int)fy0 - 1 - ylo) * pitch + ((int)fx0 - 1 - xlo));
#pragma unroll
        for (int j = 1; j < 4; j++) row[j] = row[j - 1] + 16u * (unsigned)pitch;
        wz0 = vt_bspline(-1.0f);  // the fraction along the march axis is exactly 0
        wz1 = vt_bspline(0.0f);
        wz2 = vt_bspline(1.0f);
    }
    // in-plane sums of the group's four planes, two FFMA2 per tap
    __device__ __forceinline__ void planes(unsigned ring, float (&out)[4]) const
    {
        vt_f2 a01 = 0ull, a23 = 0ull;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                vt_f2 lo, hi;
                lds128(ring + row[j] + 16u * i, lo, hi);
                const vt_f2 ww = vt_pk(w[j * 4 + i], w[j * 4 + i]);
                a01 = vt_fma2(lo, ww, a01);
                a23 = vt_fma2(hi, ww, a23);
            }
        vt_unpk(a01, out[0], out[1]);
        vt_unpk(a23, out[2], out[3]);
    }
};

// ---- cubic_tex: Ruijters' 8 trilinear fetches with the texture unit's integer weights -----------------------
// With fraction 0 along the march axis: h0 = idx + 0.3 -> planes idx-1 (S = 256-c0) and idx (S = c0), combined with
// g0; h1 = idx + 1.5 -> plane idx+1 (S = 256), combined with g1 (see vt_resample_slice.cu Taps<VT_CUBIC_TEX>).
// Each of the three planes gets its own set of 16 pre-multiplied tap weights.
template <>
struct Taps4<VT_CUBIC_TEX> {
    static constexpr int LO = -1, HI = 2;
    static constexpr int BEFORE = 1, AFTER = 1;
    float wa[16], wb[16], wc[16];
    unsigned adr[8];  // byte offsets of taps (row j, x pair k): adr[j*2+k], the pair is (adr, adr+16)
    template <int RULE>
    __device__ __forceinline__ void init(float py, float px, int ylo, int xlo, int pitch, int m)
    {
        float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
        vt_ruijters(px, g0x, g1x, h0x, h1x);
        vt_ruijters(py, g0y, g1y, h0y, h1y);
        vt_ruijters(8.5f, g0z, g1z, h0z, h1z);  // any texel centre: the fraction along the march axis is exactly 0
        const float gx[2] = {g0x, g1x}, gy[2] = {g0y, g1y};
        const float hx[2] = {h0x, h1x}, hy[2] = {h0y, h1y};
        const float gz[3] = {g0z, g0z, g1z};
        int bx[2], by[2];
        if (RULE == 0) {
            int bz0, c0, bz1, c1;
            vt_tex_fix_hw(h0z, bz0, c0);  // (7, 205)
            vt_tex_fix_hw(h1z, bz1, c1);  // (9, 0)
            // the three tap planes: (alpha, side) along the march axis = (c0, near), (c0, far), (c1 = 0, near)
            const int mu[3] = {c0, c0, c1}, ms[3] = {0, 1, 0};
            int ax[2], ay[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                vt_tex_fix_hw(hx[k], bx[k], ax[k]);
                vt_tex_fix_hw(hy[k], by[k], ay[k]);
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                float *w = p == 0 ? wa : (p == 1 ? wb : wc);
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const float g = __fmul_rn(__fmul_rn(gx[i], gy[j]), gz[p]) * (1.0f / 256.0f);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            w[(2 * j + (k >> 1)) * 4 + 2 * i + (k & 1)] =
                                g * vt_u2f(vt_tex_hw_inplane(m, ax[i], ay[j], mu[p], ms[p], k >> 1, k & 1));
                    }
            }
        } else {
            int bz0, bz1;
            float c0, c1;
            vt_tex_fix<2>(h0z, bz0, c0);
            vt_tex_fix<2>(h1z, bz1, c1);
            const float S[3] = {1.0f - c0, c0, 1.0f - c1};
            float ax[2], ay[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                vt_tex_fix<2>(hx[k], bx[k], ax[k]);
                vt_tex_fix<2>(hy[k], by[k], ay[k]);
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                float *w = p == 0 ? wa : (p == 1 ? wb : wc);
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const float g = gx[i] * gy[j] * gz[p] * S[p];
                        w[(2 * j) * 4 + 2 * i] = g * (1.0f - ax[i]) * (1.0f - ay[j]);
                        w[(2 * j) * 4 + 2 * i + 1] = g * ax[i] * (1.0f - ay[j]);
                        w[(2 * j + 1) * 4 + 2 * i] = g * (1.0f - ax[i]) * ay[j];
                        w[(2 * j + 1) * 4 + 2 * i + 1] = g * ax[i] * ay[j];
                    }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                adr[(2 * j) * 2 + k] = 16u * (unsigned)((by[j] - ylo) * pitch + (bx[k] - xlo));
                adr[(2 * j + 1) * 2 + k] = adr[(2 * j) * 2 + k] + 16u * (unsigned)pitch;
            }
    }
    // A-, B- and C-weighted in-plane sums of the group's four planes: per tap {wa, wb} x {t, t} for each plane (the
    // scalar texel is FFMA2's broadcast operand) and {t0, t1} x {wc, wc}, {t2, t3} x {wc, wc}: 6 FFMA2 per LDS.128
    __device__ __forceinline__ void planes3(unsigned ring, float (&qa)[4], float (&qb)[4], float (&qc)[4]) const
    {
        vt_f2 ab[4] = {0ull, 0ull, 0ull, 0ull}, c01 = 0ull, c23 = 0ull;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                vt_f2 lo, hi;
                lds128(ring + adr[j * 2 + (i >> 1)] + 16u * (i & 1), lo, hi);
                const vt_f2 wab = vt_pk(wa[j * 4 + i], wb[j * 4 + i]);
                const vt_f2 wcc = vt_pk(wc[j * 4 + i], wc[j * 4 + i]);
                float t0, t1, t2, t3;
                vt_unpk(lo, t0, t1);
                vt_unpk(hi, t2, t3);
                ab[0] = vt_fma2(wab, vt_pk(t0, t0), ab[0]);
                ab[1] = vt_fma2(wab, vt_pk(t1, t1), ab[1]);
                ab[2] = vt_fma2(wab, vt_pk(t2, t2), ab[2]);
                ab[3] = vt_fma2(wab, vt_pk(t3, t3), ab[3]);
                c01 = vt_fma2(lo, wcc, c01);
                c23 = vt_fma2(hi, wcc, c23);
            }
#pragma unroll
        for (int p = 0; p < 4; p++) vt_unpk(ab[p], qa[p], qb[p]);
        vt_unpk(c01, qc[0], qc[1]);
        vt_unpk(c23, qc[2], qc[3]);
    }
};

__host__ __device__ __forceinline__ int floordiv4(int v) { return v >> 2; }  // arithmetic shift: floor for negatives

// One block-wide barrier per step: everyone is done with the stage consumed in the previous step, thread 0 refills it.
// (A ring without the barrier -- every warp waits on the stage's "full" mbarrier, arrives on an "empty" one, and the
// last warp out of a stage refills it -- was built and measured SLOWER: cubic_tex 219 vs 264, cubic_simple 279 vs 300
// Gvox/s at 256^3; the elected-lane bookkeeping of every warp and step costs more than the barrier it removes.)
// TSY: rows of the tile (16 or 8).  8-row tiles = 128-thread CTAs, twice as many per SM: the same number of warps, but
// barriers that tie 4 warps instead of 8 and CTAs that drift apart more (knob VT_Z4_TSY, measured in DESIGN.md).
#ifndef VT_Z4_L2_HINTS
#define VT_Z4_L2_HINTS 0   // bit 0: TMA loads with L2 evict_last; bit 1: streaming output stores (A/B: tools/z4_l2_ab.sh)
#endif
#ifndef VT_Z4_CT8_RESIDENT
#define VT_Z4_CT8_RESIDENT 4  // resident 128-thread CTAs per SM the cubic_tex kernel is compiled for (register cap 65536 / 128 / N)
#endif
template <int INTERP>
__host__ __device__ constexpr int z4_resident(int tsy)
{
    return INTERP == VT_CUBIC_TEX ? (tsy == 8 ? VT_Z4_CT8_RESIDENT : 2) : 3 * (TS / tsy);
}
// VEC: the march axis is the contiguous output axis (march 2): quads + TMA stores (compiled out otherwise).
template <int INTERP, int RULE, bool OOB_ZERO, int NSTAGE, int TSY, bool VEC>
__global__ void __launch_bounds__(TS * TSY, z4_resident<INTERP>(TSY))
    vt_z4_kernel(const __grid_constant__ Z4Params P, const __grid_constant__ Z4Maps G, int m_chunk, unsigned stage_bytes)
{
    // [SMEM_HEADER: mbarriers, counters][NSTAGE stages of stage_bytes]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using T = Taps4<INTERP>;
    static_assert(NSTAGE == 3 || NSTAGE == 4, "the ring loop is unrolled for 3 or 4 stages");
    vt_pdl_wait();
    const int tid = threadIdx.x;
    // Block order.  Tile-fastest (the default: one matrix per wave of CTAs) or matrix-fastest (the same output tile of all
    // matrices side by side: their footprints overlap, DRAM reads of a 32-matrix launch at 256^3 drop from 1.91 to 1.09 GB).
    // Measured: no gain for the cubic kernels (shared-memory bound), and `linear` loses 8 % (neighbouring tiles of one
    // output volume no longer write the two halves of a cache line at the same time) -- profiles/r02w_z4_block_order.log.
    const int tile = P.mat_fastest ? (int)blockIdx.x / P.n_mats : (int)blockIdx.x;
    const int mat = P.mat_fastest ? (int)blockIdx.x - tile * P.n_mats : (int)blockIdx.z;
    const int tile_y = tile / P.tiles_fast, tile_x = tile - tile_y * P.tiles_fast;
    const Z4Mat &M = P.mats[mat];
    const int pidx = M.pitch_idx;
    const int pitch = P.bw0 + pidx;
    const int tm = M.tm;
    const int pof = P.prod_on_fast;
    const int zc0 = P.m_b + blockIdx.y * m_chunk;
    const int zc1 = min(zc0 + m_chunk, P.m_e);
    const int as0 = P.slow_b + tile_y * TSY, af0 = tile_x * TS;
    const int as1 = min(as0 + TSY, P.slow_e) - 1, af1 = min(af0 + TS, P.o_fast) - 1;

    // footprint of the tile: extremes are at the corners (the float recipe is monotone in either index)
    float y_min, y_max, x_min, x_max;
    {
        const float sa = (float)as0, sb = (float)as1, fa = (float)af0, fb = (float)af1;
        const float y00 = z4_coord(M.ys, M.yf, M.yc, sa, fa, pof), y01 = z4_coord(M.ys, M.yf, M.yc, sa, fb, pof);
        const float y10 = z4_coord(M.ys, M.yf, M.yc, sb, fa, pof), y11 = z4_coord(M.ys, M.yf, M.yc, sb, fb, pof);
        const float x00 = z4_coord(M.xs, M.xf, M.xc, sa, fa, pof), x01 = z4_coord(M.xs, M.xf, M.xc, sa, fb, pof);
        const float x10 = z4_coord(M.xs, M.xf, M.xc, sb, fa, pof), x11 = z4_coord(M.xs, M.xf, M.xc, sb, fb, pof);
        y_min = fminf(fminf(y00, y01), fminf(y10, y11));
        y_max = fmaxf(fmaxf(y00, y01), fmaxf(y10, y11));
        x_min = fminf(fminf(x00, x01), fminf(x10, x11));
        x_max = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11));
    }
    // clamp to a couple of texels around the source: everything further out is border (zero) anyway and columns that
    // far out are out of bounds; keeps the integer conversions safe for wild matrices
    y_min = fmaxf(y_min, -2.0f); x_min = fmaxf(x_min, -2.0f);
    y_max = fminf(y_max, (float)P.s_y + 2.0f); x_max = fminf(x_max, (float)P.s_x + 2.0f);
    const int ylo = (int)floorf(y_min - 0.5f) + T::LO, xlo = (int)floorf(x_min - 0.5f) + T::LO;

    const unsigned bars_s = vt_smem_u32(smem_raw);
    const unsigned ring_s = bars_s + (unsigned)SMEM_HEADER;
    if (tid == 0) {
        vt_tma_prefetch_desc(&G.map[pidx]);
#pragma unroll
        for (int i = 0; i < NSTAGE; i++) vt_mbar_init(bars_s + 8u * i, 1);
        vt_mbar_fence_init();
    }
    __syncthreads();

    // this thread's column
    int ty, tx;
    z4_lane_pos(tid, M.shape, ty, tx);
    const int a_s = as0 + ty, a_f = af0 + tx;
    const bool live = a_s < P.slow_e && a_f < P.o_fast;
    const float py = z4_coord(M.ys, M.yf, M.yc, (float)a_s, (float)a_f, pof);
    const float px = z4_coord(M.xs, M.xf, M.xc, (float)a_s, (float)a_f, pof);
    // transforms.py:276-278 for the two in-plane axes
    const bool inplane = live && !(px < 0 || py < 0 || px >= (float)P.s_x || py >= (float)P.s_y);

    // input planes needed: q = z + tm + d, d in [-BEFORE, AFTER]; group g holds planes 4g .. 4g+3
    const int q_first = zc0 + tm - T::BEFORE, q_last = zc1 - 1 + tm + T::AFTER;
    const int g_last = floordiv4(q_last);
    int g = floordiv4(q_first);
    const unsigned tx_bytes = 16u * (unsigned)(pitch * P.bh);

#if VT_Z4_L2_HINTS & 1
    const unsigned long long l2_keep = vt_l2_policy_evict_last();
#endif
    // one elected thread: stage group gg into ring stage st
    auto load_group = [&](int gg, unsigned st) {
        const unsigned bar = bars_s + 8u * st;
        vt_mbar_expect_tx(bar, tx_bytes);
        // groups / rows / columns outside the source arrive as zeros (= the texture's border mode)
#if VT_Z4_L2_HINTS & 1
        vt_tma_load_3d_hint(ring_s + st * stage_bytes, &G.map[pidx], bar, 4 * xlo, ylo, gg, l2_keep);
#else
        vt_tma_load_3d(ring_s + st * stage_bytes, &G.map[pidx], bar, 4 * xlo, ylo, gg);
#endif
    };
    constexpr int AHEAD = NSTAGE - 1;  // groups in flight after the prologue
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < AHEAD; i++)
            if (g + i <= g_last) load_group(g + i, (unsigned)i);
    }
    // the column's weights are computed while those first loads are in flight
    T taps;
    if (inplane) taps.template init<RULE>(py, px, ylo, xlo, pitch, P.march);
    float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // sliding window of per-plane sums
    unsigned phase = 0;
    const long long osm = P.os_m;
    // output pointer of this column at the output plane that input plane 4g is the LAST tap plane of
    float *dstp = P.dst + (size_t)mat * P.dst_batch_stride + (long long)a_s * P.os_slow + (long long)a_f * P.os_fast +
                  (long long)(4 * g - T::AFTER - tm) * P.os_m;

    // vector-store path (march axis = contiguous output axis)
    float prev[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    // a step's first output index is 4*gg - AFTER - tm: how many values of an aligned output quad come from the
    // previous step
    const int qshift = (4 - ((T::AFTER + tm) & 3)) & 3;
    float *const dcol = P.dst + (size_t)mat * P.dst_batch_stride + (long long)a_s * P.os_slow + (long long)a_f * P.os_fast;
    auto emit_quad = [&](int zq, const float (&v)[4]) {  // outputs zq .. zq+3 (zq a multiple of 4) of this column
        const bool all_in = zq >= zc0 && zq + 3 < zc1;                             // uniform
        const bool all_src = zq + tm >= 0 && zq + 3 + tm < P.s_m;                 // uniform
        if (all_in && (OOB_ZERO || all_src)) {
            if (OOB_ZERO ? live : inplane) {
                const float4 q = (inplane && all_src) ? make_float4(v[0], v[1], v[2], v[3])
                                                      : make_float4(inplane && (unsigned)(zq + tm) < (unsigned)P.s_m ? v[0] : 0.0f,
                                                                    inplane && (unsigned)(zq + 1 + tm) < (unsigned)P.s_m ? v[1] : 0.0f,
                                                                    inplane && (unsigned)(zq + 2 + tm) < (unsigned)P.s_m ? v[2] : 0.0f,
                                                                    inplane && (unsigned)(zq + 3 + tm) < (unsigned)P.s_m ? v[3] : 0.0f);
                *reinterpret_cast<float4 *>(dcol + zq) = q;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int zo = zq + j;
                if (zo >= zc0 && zo < zc1) {
                    const bool ok = inplane && (unsigned)(zo + tm) < (unsigned)P.s_m;
                    if (OOB_ZERO) {
                        if (live) dcol[zo] = ok ? v[j] : 0.0f;
                    } else if (ok) {
                        dcol[zo] = v[j];
                    }
                }
            }
        }
    };

    // TMA store path (march axis 2, barrier ring): every thread parks its aligned quads in a shared-memory tile laid out
    // as the box (x: 8, a1: 16, a0: 16) -- two consecutive steps fill a column's 32-byte sector --; after the next
    // block-wide barrier one thread hands the tile to the TMA engine (cp.async.bulk.tensor store).  256 whole sectors
    // leave as one bulk operation instead of sixteen STG.128 instructions that each touch 32 different cache lines with
    // half a sector.  Only for tiles every column of which may be written (OOB_SKIP: all columns inside the source
    // in-plane); two tiles alternate.
    const unsigned otile_s = ring_s + NSTAGE * stage_bytes;
    bool tile_ok = false;
    if (VEC && P.tma_store) {
        const bool mine = OOB_ZERO ? true : (inplane || !live);
        tile_ok = __syncthreads_and(mine ? 1 : 0) != 0 && as0 + TSY <= P.slow_e;
    }
    int pending_zq = 0, pending_buf = -1, obuf = 0;  // a tile completed in the previous step, not yet handed over
    bool half = false;                                // the even quad of the current tile is in place
    auto flush_tile = [&]() {                         // after a block-wide barrier
        if (pending_buf >= 0 && tid == 0) {
            vt_fence_proxy_async();
            vt_tma_store_4d(&G.omap, otile_s + (unsigned)pending_buf * OTILE_BYTES, pending_zq, af0, as0, mat);
            vt_bulk_commit();
        }
        pending_buf = -1;
    };

    auto stage_step = [&](auto cur_c, int gg) {
        constexpr unsigned CUR = decltype(cur_c)::value;
        constexpr unsigned FILL = (CUR + NSTAGE - 1) % NSTAGE;
        const unsigned ring = ring_s + CUR * stage_bytes;
        vt_mbar_wait(bars_s + 8u * CUR, phase);
        if (VEC && tile_ok && tid == 0) vt_bulk_wait_read<1>();  // the tile this step may overwrite has been read
        __syncthreads();  // everyone is done with the stage consumed in the previous step: refill it
        if (tid == 0 && gg + NSTAGE - 1 <= g_last) load_group(gg + NSTAGE - 1, FILL);
        if (VEC && tile_ok) flush_tile();
        float r[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (inplane) {
            if constexpr (INTERP == VT_LINEAR) {
                taps.planes(ring, r);
            } else if constexpr (INTERP == VT_CUBIC_SIMPLE) {
                float pq[4];
                taps.planes(ring, pq);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    // the reference accumulates kz = -1, 0, 1 in that order (helper_interpolation.h:51)
                    r[p] = fmaf(taps.wz2, pq[p], fmaf(taps.wz1, s1, __fmul_rn(taps.wz0, s2)));
                    s2 = s1;
                    s1 = pq[p];
                }
            } else {
                float qa[4], qb[4], qc[4];
                taps.planes3(ring, qa, qb, qc);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    r[p] = (s3 + s1) + qc[p];  // s3 = A-sum of plane q-2, s1 = B-sum of plane q-1
                    s3 = s2;
                    s2 = qa[p];
                    s1 = qb[p];
                }
            }
        }
        if constexpr (VEC) {
            // The march axis is the contiguous output axis (march 2): a thread's four values are neighbours in memory.
            // Re-cut them into 16-byte aligned quads -- `qa` values of a quad come from the previous step -- and store
            // each quad with one STG.128 (elements outside the chunk / the source fall back to scalar stores).
            float v[4];
            switch (qshift) {  // uniform
                case 0: v[0] = r[0]; v[1] = r[1]; v[2] = r[2]; v[3] = r[3]; break;
                case 1: v[0] = prev[3]; v[1] = r[0]; v[2] = r[1]; v[3] = r[2]; break;
                case 2: v[0] = prev[2]; v[1] = prev[3]; v[2] = r[0]; v[3] = r[1]; break;
                default: v[0] = prev[1]; v[1] = prev[2]; v[2] = prev[3]; v[3] = r[0]; break;
            }
            const int zq = 4 * gg - T::AFTER - tm - qshift;
            // an even quad (x a multiple of 8) opens a tile if its partner, the next step's quad, can be staged too
            const int odd = (zq >> 2) & 1;
            const int z8 = zq - 4 * odd;  // the sector: outputs z8 .. z8+7
            const bool sector_ok = tile_ok && z8 >= zc0 && z8 + 7 < zc1 &&
                                   (OOB_ZERO || (z8 + tm >= 0 && z8 + 7 + tm < P.s_m)) && 4 * g_last + 3 - T::AFTER - tm >= z8 + 7;
            if (sector_ok && (odd ? half : true)) {  // uniform
                float4 q = make_float4(v[0], v[1], v[2], v[3]);
                if (OOB_ZERO) {  // values past the ends of the source along the march axis are zeros
                    if ((unsigned)(zq + tm) >= (unsigned)P.s_m) q.x = 0.0f;
                    if ((unsigned)(zq + 1 + tm) >= (unsigned)P.s_m) q.y = 0.0f;
                    if ((unsigned)(zq + 2 + tm) >= (unsigned)P.s_m) q.z = 0.0f;
                    if ((unsigned)(zq + 3 + tm) >= (unsigned)P.s_m) q.w = 0.0f;
                }
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(otile_s + (unsigned)obuf * OTILE_BYTES +
                                                                              32u * (unsigned)(ty * TS + tx) + 16u * odd),
                             "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w)
                             : "memory");
                if (odd) {
                    pending_zq = z8;
                    pending_buf = obuf;
                    obuf ^= 1;
                    half = false;
                } else {
                    half = true;
                }
            } else {
                emit_quad(zq, v);
            }
#pragma unroll
            for (int p = 0; p < 4; p++) prev[p] = r[p];
        } else {
            const int zi0 = 4 * gg - T::AFTER;  // input plane at the centre of the step's first output voxel
            const int zo0 = zi0 - tm;          // its output index along the march axis
            // (r[] is zero for columns outside the source in-plane)
            if (zo0 >= zc0 && zo0 + 3 < zc1 && zi0 >= 0 && zi0 + 3 < P.s_m) {  // uniform: a step in the interior
                if (OOB_ZERO ? live : inplane) {
                    float *d = dstp;
#pragma unroll
                    for (int p = 0; p < 4; p++) {
#if VT_Z4_L2_HINTS & 2
                        __stcs(d, r[p]);  // streaming (evict-first) store: the outputs must not push the volume out of L2
#else
                        *d = r[p];
#endif
                        d += osm;
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const int zi = zi0 + p, zo = zo0 + p;
                    if (zo >= zc0 && zo < zc1) {  // uniform: past the warm-up planes, inside the chunk
                        const bool ok = inplane && (unsigned)zi < (unsigned)P.s_m;  // 0 <= p_m < s_m with p_m = zi + 0.5
                        if (OOB_ZERO) {
                            if (live) dstp[(long long)p * osm] = ok ? r[p] : 0.0f;
                        } else if (ok) {
                            dstp[(long long)p * osm] = r[p];
                        }
                    }
                }
            }
        }
        dstp += 4 * osm;
    };
    for (;;) {
        stage_step(std::integral_constant<unsigned, 0>{}, g);
        if (++g > g_last) break;
        stage_step(std::integral_constant<unsigned, 1>{}, g);
        if (++g > g_last) break;
        stage_step(std::integral_constant<unsigned, 2>{}, g);
        if (++g > g_last) break;
        if constexpr (NSTAGE == 4) {
            stage_step(std::integral_constant<unsigned, 3>{}, g);
            if (++g > g_last) break;
        }
        phase ^= 1u;
    }
    if (VEC && tile_ok) {  // the last staged tile
        __syncthreads();
        flush_tile();
        if (tid == 0) vt_bulk_wait_read<0>();  // shared memory must outlive the engine's reads
    }
    if (VEC && qshift) {  // the values of the last, incomplete quad
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        switch (qshift) {  // uniform
            case 1: v[0] = prev[3]; break;
            case 2: v[0] = prev[2]; v[1] = prev[3]; break;
            default: v[0] = prev[1]; v[1] = prev[2]; v[2] = prev[3]; break;
        }
        emit_quad(4 * (g_last + 1) - T::AFTER - tm - qshift, v);
    }
}

// ---------------------------------------------------------------------------------------------------
// layout conversion: plain (padded rows) -> Z4 of axis m
// ---------------------------------------------------------------------------------------------------
// one thread per Z4 texel (g, y, x): four loads along axis m (stride sm), one 16-byte store.  Lanes run along x,
// the contiguous axis of the destination; for m = 0, 1 that is also the contiguous axis of the source.  For m = 2
// the four loads of a thread are one aligned 16-byte load when the rows allow it.
__global__ void __launch_bounds__(256)
    vt_pack_z4_kernel(const float *__restrict__ src, float4 *__restrict__ dst, int dm, int dy, int dx, long long sm,
                      long long sy, long long sx)
{
    vt_pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y, g = blockIdx.z;
    if (x >= dx) return;
    const float *p = src + (long long)(4 * g) * sm + (long long)y * sy + (long long)x * sx;
    float4 v;
    v.x = __ldg(p);
    v.y = 4 * g + 1 < dm ? __ldg(p + sm) : 0.0f;
    v.z = 4 * g + 2 < dm ? __ldg(p + 2 * sm) : 0.0f;
    v.w = 4 * g + 3 < dm ? __ldg(p + 3 * sm) : 0.0f;
    dst[((size_t)g * dy + y) * dx + x] = v;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct AxisRoles {
    int m, ay, ax;         // march axis, source in-plane axes (rows, columns): ascending
    int slow, fast;        // output axes of the tile
    int prod_on_fast;
};
AxisRoles roles(int m)
{
    switch (m) {
        case 0: return {0, 1, 2, 1, 2, 0};
        case 1: return {1, 0, 2, 0, 2, 0};
        default: return {2, 0, 1, 0, 1, 1};
    }
}

bool z4_mat_ok(const VtMat &M, int m, const int sdim[3])
{
    const AxisRoles R = roles(m);
    if (sdim[m] >= 16384) return false;  // h0 = idx + 0.3 along the march axis must keep its 1/256 quantum
    for (int c = 0; c < 3; c++)
        if (M.r[m][c] != (c == m ? 1.0f : 0.0f)) return false;
    if (M.r[R.ay][m] != 0.0f || M.r[R.ax][m] != 0.0f) return false;
    const float t = M.r[m][3];
    if (!(fabsf(t) < 16384.0f) || t != floorf(t)) return false;
    const int rows[2] = {R.ay, R.ax};
    const int lim[2] = {BH_MAX - 6, BW_MAX - (NPITCH - 1) - 6};  // footprint edge <= floor(ext + 0.1) + 6 texels
    for (int i = 0; i < 2; i++) {
        const float ext = (fabsf(M.r[rows[i]][R.slow]) + fabsf(M.r[rows[i]][R.fast])) * (float)(TS - 1);
        if (!(ext <= (float)lim[i] - 0.1f)) return false;
        if (!(fabsf(M.r[rows[i]][3]) < 1e6f)) return false;
    }
    return true;
}

// Average wavefronts per quarter-warp LDS.128 of a matrix for every (pitch mod 8, quarter shape): lanes land on
// 16-byte bank group (y*pitch + x) mod 8 of their footprint origin (every other tap shifts all lanes alike).
// Four sample tiles; a tile's 256 origins are computed once and regrouped for each shape.  Memoised per matrix.
struct Z4Conflicts {
    float key[8];
    bool valid;
    float wf[8][N_SHAPES];
};
// (the statistics do not need the exact float32 recipe: double arithmetic, no libm fmaf on the host)
static inline int z4_origin(float cs, float cf, float cc, int as, int af)
{
    return (int)floor((double)cc + (double)as * cs + (double)af * cf);  // floor(p - 0.5), p = ... + 0.5
}
const Z4Conflicts &z4_conflicts(const Z4Mat &M, int pof, int tsy)
{
    constexpr int MEMO = 2048, WAYS = 4;
    static thread_local Z4Conflicts memo[MEMO];
    static thread_local unsigned victim = 0;
    const float key[8] = {M.ys, M.yf, M.yc, M.xs, M.xf, M.xc, (float)pof, (float)tsy};
    unsigned h = 2166136261u;
    for (int i = 0; i < 8; i++) {
        unsigned u;
        memcpy(&u, &key[i], 4);
        h = (h ^ u) * 16777619u;
    }
    const unsigned slot = (h ^ (h >> 13)) & (MEMO - 1);
    for (int w = 0; w < WAYS; w++) {
        const Z4Conflicts &c = memo[(slot + w) & (MEMO - 1)];
        if (c.valid && memcmp(c.key, key, sizeof key) == 0) return c;
    }
    int way = 0;
    for (int w = 0; w < WAYS; w++)
        if (!memo[(slot + w) & (MEMO - 1)].valid) { way = w; break; } else way = (int)(victim++ % WAYS);
    Z4Conflicts &e = memo[(slot + way) & (MEMO - 1)];
    memcpy(e.key, key, sizeof key);
    constexpr int NSMP = 3, NQ = 10;
    static const unsigned char T1[NSMP] = {1, 14, 5}, T2[NSMP] = {3, 8, 29};
    int cost[8][N_SHAPES] = {};
    for (int smp = 0; smp < NSMP; smp++) {
        const int as0 = 16 * T1[smp], af0 = 16 * T2[smp];
        const int y0 = z4_origin(M.ys, M.yf, M.yc, as0, af0), x0 = z4_origin(M.xs, M.xf, M.xc, as0, af0);
        int oy[TS][TS], ox[TS][TS];
        for (int ty = 0; ty < TS; ty++)
            for (int tx = 0; tx < TS; tx++) {
                oy[ty][tx] = z4_origin(M.ys, M.yf, M.yc, as0 + ty, af0 + tx) - y0;
                ox[ty][tx] = z4_origin(M.xs, M.xf, M.xc, as0 + ty, af0 + tx) - x0;
            }
        for (int s = 0; s < N_SHAPES; s++)
            for (int qi = 0; qi < NQ; qi++) {
                const int q = (5 * qi + 1 + smp) & (2 * tsy - 1);  // a sample of the tile's quarter warps
                int yy[8], xx[8], n = 0;
                for (int l = 0; l < 8; l++) {
                    int ty, tx;
                    z4_lane_pos(q * 8 + l, s, ty, tx);
                    const int y = oy[ty][tx], x = ox[ty][tx];
                    bool dup = false;  // same texel as an earlier lane: broadcast
                    for (int k = 0; k < n; k++) dup |= (yy[k] == y && xx[k] == x);
                    if (!dup) {
                        yy[n] = y;
                        xx[n] = x;
                        n++;
                    }
                }
                for (int pm = 0; pm < 8; pm++) {
                    unsigned cnt = 0;  // eight 4-bit counters, one per 16-byte bank group
                    for (int l = 0; l < n; l++) cnt += 1u << (4 * ((yy[l] * pm + xx[l]) & 7));
                    int worst = 0;
                    for (; cnt; cnt >>= 4) worst = std::max(worst, (int)(cnt & 15u));
                    cost[pm][s] += worst;
                }
            }
    }
    for (int pm = 0; pm < 8; pm++)
        for (int s = 0; s < N_SHAPES; s++) e.wf[pm][s] = (float)cost[pm][s] / (float)(NSMP * NQ);
    e.valid = true;
    return e;
}

struct Z4Plan {
    int chunks, m_chunk;
    int bw0, bh;
    float cost;  // modelled wavefronts per quarter-warp load, averaged over the matrices
};

template <int INTERP>
int plan_z4(Z4Params &P, int sms, const float ext_y[], const float ext_x[], Z4Plan &L)
{
    using T = Taps4<INTERP>;
    int need_h = 0, need_w = 0;
    for (int k = 0; k < P.n_mats; k++) {
        need_h = std::max(need_h, (int)floorf(ext_y[k] + 0.1f) + 6);
        need_w = std::max(need_w, (int)floorf(ext_x[k] + 0.1f) + 6);
    }
    need_h = std::min(need_h, BH_MAX);
    need_w = std::min(need_w, BW_MAX - (NPITCH - 1));
    // never wider / taller than the source plus its border: small volumes keep small boxes
    L.bw0 = need_w;
    L.bh = need_h;
    P.bw0 = L.bw0;
    P.bh = L.bh;
    // per matrix: quarter shape and pitch class with the fewest simulated wavefronts.  A wider box costs L2->shared
    // traffic (1/27 per texel of width), 4x2 / 8x1 patches make a warp's stores cover 4 / 8 rows.
    const char *force_s = getenv("VT_Z4_SHAPE"), *force_p = getenv("VT_Z4_PITCH");  // tuning knobs
    float total = 0.0f;
    for (int k = 0; k < P.n_mats; k++) {
        const Z4Conflicts &C = z4_conflicts(P.mats[k], P.prod_on_fast, P.tsy);
        float best = 1e30f;
        int bs = 0, bp = 0;
        for (int pi = 0; pi < NPITCH; pi++)
            for (int s = 0; s < N_SHAPES; s++) {
                // (8x1 patches make a warp store 8 rows x 16 bytes: half sectors, measured ~2x slower stores)
                const float pen = 0.012f * (float)pi + (s == 2 ? 0.02f : (s == 3 ? 0.30f : 0.0f));
                const float c = C.wf[(L.bw0 + pi) & 7][s] + pen;
                if (c < best) {
                    best = c;
                    bs = s;
                    bp = pi;
                }
            }
        if (force_s) bs = atoi(force_s) & 3;
        if (force_p) bp = atoi(force_p) & 7;
        P.mats[k].shape = (unsigned char)bs;
        P.mats[k].pitch_idx = (unsigned char)bp;
        total += C.wf[(L.bw0 + bp) & 7][bs];
    }
    L.cost = total / (float)P.n_mats;
    // chunks along the march axis: a CTA marches (m_chunk + WARM)/4 groups and pays a fixed start-up; CTAs run in
    // waves of (SMs x resident CTAs).  Long marches over large planes lose the L2 reuse between neighbouring tiles
    // (see vt_resample_slice.cu): cap a march at 2 GB of source planes.
    const int nm = P.m_e - P.m_b;
    const int tiles = ((P.slow_e - P.slow_b + P.tsy - 1) / P.tsy) * P.tiles_fast;
    constexpr int WARM = T::BEFORE + T::AFTER + 3;  // + up to 3 planes of group misalignment
    const int RESIDENT = z4_resident<INTERP>(P.tsy);
    constexpr int STARTUP = 12;
    const long long slots = (long long)sms * RESIDENT, per_chunk = (long long)tiles * P.n_mats;
    const long long plane_bytes = (long long)P.s_y * P.s_x * 4;
    const long long march_bytes = INTERP == VT_LINEAR ? (512LL << 20) : (2048LL << 20);
    const int march_cap = (int)std::max(32LL, march_bytes / std::max(plane_bytes, 1LL));
    int chunks = 1, m_chunk = nm;
    long long best_cost = -1;
    for (int c = 1; c <= 256 && (c == 1 || nm / c >= 16); c++) {
        int zc = ((nm + c - 1) / c + 3) / 4 * 4;  // whole groups
        if (zc < 4) break;
        if (zc > march_cap && nm / (c + 1) >= 16) continue;
        const int cc = (nm + zc - 1) / zc;
        const long long waves = (per_chunk * cc + slots - 1) / slots;
        const long long cost = waves * (zc + WARM + STARTUP);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            chunks = cc;
            m_chunk = zc;
        }
    }
    if (const char *e = getenv("VT_Z4_CHUNKS"))
        if (atoi(e) > 0) {
            m_chunk = ((nm + atoi(e) - 1) / atoi(e) + 3) / 4 * 4;
            chunks = (nm + m_chunk - 1) / m_chunk;
        }
    if (chunks > 65535) return VT_ERR_UNSUPPORTED;
    L.chunks = chunks;
    L.m_chunk = m_chunk;
    return VT_OK;
}

template <int INTERP, int RULE, int TSY>
int launch3(Z4Params &P, const float *d_src4, bool oob_zero, const float ext_y[], const float ext_x[], cudaStream_t st)
{
    int sms = 148, dev = 0;
    VT_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    Z4Plan L;
    int rc = plan_z4<INTERP>(P, sms, ext_y, ext_x, L);
    if (rc) return rc;
    Z4Maps G;
    memset(&G, 0, sizeof G);
    const unsigned long long groups = (unsigned long long)((P.s_m + 3) / 4);
    const unsigned long long gdim[3] = {4ull * (unsigned long long)P.s_x, (unsigned long long)P.s_y, groups};
    const unsigned long long gstr[2] = {16ull * (unsigned long long)P.s_x, 16ull * (unsigned long long)P.s_x * P.s_y};
    bool used[NPITCH] = {};
    for (int k = 0; k < P.n_mats; k++) used[P.mats[k].pitch_idx] = true;
    for (int pi = 0; pi < NPITCH; pi++) {
        if (!used[pi]) continue;
        const unsigned box[3] = {4u * (unsigned)(L.bw0 + pi), (unsigned)L.bh, 1u};
        rc = vt_encode_tmap_3d(&G.map[pi], d_src4, gdim, gstr, box);
        if (rc) return rc;
    }
    if (P.tma_store) {
        // output as a 4-D tensor (x, a1, a0, matrix); the box is what a CTA produces per step for march axis 2
        const unsigned long long odim[4] = {(unsigned long long)P.o_m, (unsigned long long)P.o_fast, (unsigned long long)P.o_slow,
                                            (unsigned long long)P.n_mats};
        const unsigned long long ostr[3] = {(unsigned long long)P.os_fast * 4, (unsigned long long)P.os_slow * 4,
                                            (unsigned long long)(P.n_mats > 1 ? P.dst_batch_stride : P.os_slow * P.o_slow) * 4};
        const unsigned obox[4] = {8u, (unsigned)TS, (unsigned)TSY, 1u};
        rc = vt_encode_tmap_nd(&G.omap, P.dst, 4, odim, ostr, obox);
        if (rc) P.tma_store = 0;  // (strides the engine cannot take: the STG.128 path writes the same values)
    }
    if (getenv("VT_Z4_DEBUG"))
        fprintf(stderr, "z4 interp %d: box %d x %d, mat0 shape %d pitch +%d, wavefronts %.3f, chunks %d x %d\n", INTERP, L.bw0,
                L.bh, P.mats[0].shape, P.mats[0].pitch_idx, L.cost, L.chunks, L.m_chunk);
    const int tiles = ((P.slow_e - P.slow_b + TSY - 1) / TSY) * P.tiles_fast;
    static const int order_env = getenv("VT_Z4_ORDER") ? atoi(getenv("VT_Z4_ORDER")) : -1;  // A/B knob: 0 tile-, 1 matrix-fastest
    P.mat_fastest = order_env > 0 && P.n_mats > 1;
    if ((long long)tiles * P.n_mats > 0x7fffffffLL) P.mat_fastest = 0;
    dim3 grid(P.mat_fastest ? tiles * P.n_mats : tiles, L.chunks, P.mat_fastest ? 1 : P.n_mats);
    // ring: stages sized for this launch's box; four stages when the resident CTAs still fit, else three
    const unsigned stage_bytes = ((unsigned)(16 * (L.bw0 + NPITCH - 1) * L.bh) + 127u) & ~127u;
    constexpr int RESIDENT = z4_resident<INTERP>(TSY);
    static const int force_stages = getenv("VT_Z4_STAGES") ? atoi(getenv("VT_Z4_STAGES")) : 0;  // tuning knob
    const size_t otile = P.tma_store ? 2 * OTILE_BYTES : 0;
    int nstage = ((size_t)SMEM_HEADER + 4u * (size_t)stage_bytes + otile) * RESIDENT <= (size_t)220 * 1024 ? 4 : 3;
    if (force_stages == 3 || force_stages == 4) nstage = force_stages;
    const size_t smem = SMEM_HEADER + (size_t)nstage * stage_bytes + otile;
    // the attribute is per device and per function: one flag per device (a process may drive several GPUs)
    static std::atomic<bool> attr_set_dev[64];
    std::atomic<bool> &attr_set = attr_set_dev[dev & 63];
    if (!attr_set.load(std::memory_order_acquire)) {
        const int mx = SMEM_HEADER + 4 * STAGE_BYTES_MAX + 2 * OTILE_BYTES;
#define VT_Z4_ATTR(Z, N) \
    VT_CUDA(cudaFuncSetAttribute(vt_z4_kernel<INTERP, RULE, Z, N, TSY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx)); \
    VT_CUDA(cudaFuncSetAttribute(vt_z4_kernel<INTERP, RULE, Z, N, TSY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))
        VT_Z4_ATTR(true, 3); VT_Z4_ATTR(false, 3); VT_Z4_ATTR(true, 4); VT_Z4_ATTR(false, 4);
#undef VT_Z4_ATTR
        attr_set.store(true, std::memory_order_release);
    }
    cudaError_t launch_err = cudaSuccess;
    {
        VtProf prof(VT_K_Z4_LINEAR + INTERP, st);
#define VT_Z4_GO(Z, N)                                                                                        \
    do {                                                                                                      \
        if (P.vec_store)                                                                                      \
            launch_err = vt_launch_pdl(vt_z4_kernel<INTERP, RULE, Z, N, TSY, true>, grid, dim3(TS * TSY), smem, st, P, G,         \
                                       L.m_chunk, stage_bytes);                                               \
        else                                                                                                  \
            launch_err = vt_launch_pdl(vt_z4_kernel<INTERP, RULE, Z, N, TSY, false>, grid, dim3(TS * TSY), smem, st, P, G,        \
                                       L.m_chunk, stage_bytes);                                               \
    } while (0)
        if (oob_zero) {
            if (nstage == 4) VT_Z4_GO(true, 4); else VT_Z4_GO(true, 3);
        } else {
            if (nstage == 4) VT_Z4_GO(false, 4); else VT_Z4_GO(false, 3);
        }
#undef VT_Z4_GO
    }
    vt_count_launch();
    VT_CUDA(launch_err);
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

template <int INTERP, int RULE>
int launch2(Z4Params &P, const float *d_src4, bool oob_zero, const float ext_y[], const float ext_x[], cudaStream_t st)
{
    if (P.tsy == 8) return launch3<INTERP, RULE, 8>(P, d_src4, oob_zero, ext_y, ext_x, st);
    return launch3<INTERP, RULE, 16>(P, d_src4, oob_zero, ext_y, ext_x, st);
}

// rows of a tile for this interpolator (knob VT_Z4_TSY = 8 / 16)
int z4_tile_rows(int interp)
{
    static const int force = getenv("VT_Z4_TSY") ? atoi(getenv("VT_Z4_TSY")) : 0;
    if (force == 8 || force == 16) return force;
    // measured (profiles/r02g_z4_tile_rows.log): 8-row tiles +4-5 % for the cubic kernels (sweep 343 vs 329 Gvox/s), the
    // linear kernel, bound by L2 -> shared traffic, loses 2-4 % to the larger footprint overlap
    return interp == VT_LINEAR ? 16 : 8;
}

// VtResampleParams (logical source dims, output, z-range, matrices) -> Z4Params for march axis m
void fill_z4(const VtResampleParams &P, int m, int interp, Z4Params &Q, float ext_y[], float ext_x[])
{
    const AxisRoles R = roles(m);
    const int sdim[3] = {P.s0, P.s1, P.s2}, odim[3] = {P.o0, P.o1, P.o2};
    const long long ostr[3] = {(long long)P.o1 * P.o2, (long long)P.o2, 1};
    memset(&Q, 0, sizeof Q);
    Q.dst = P.dst;
    Q.dst_batch_stride = P.dst_batch_stride;
    Q.os_slow = ostr[R.slow];
    Q.os_fast = ostr[R.fast];
    Q.os_m = ostr[m];
    Q.o_fast = odim[R.fast];
    Q.o_slow = odim[R.slow];
    Q.o_m = odim[m];
    // the C ABI's z-range restricts output axis 0: the march axis for m = 0, the slow tile axis otherwise
    Q.slow_b = m == 0 ? 0 : P.z_begin;
    Q.slow_e = m == 0 ? odim[R.slow] : P.z_end;
    Q.m_b = m == 0 ? P.z_begin : 0;
    Q.m_e = m == 0 ? P.z_end : odim[m];
    Q.s_y = sdim[R.ay];
    Q.s_x = sdim[R.ax];
    Q.s_m = sdim[m];
    Q.n_mats = P.n_mats;
    Q.prod_on_fast = R.prod_on_fast;
    Q.march = m;
    Q.vec_store = (m == 2 && (P.o2 % 4) == 0 && (P.dst_batch_stride % 4) == 0 && ((uintptr_t)P.dst % 16) == 0 &&
                   !getenv("VT_Z4_NO_VEC")) ? 1 : 0;
    Q.tma_store = (Q.vec_store && (P.o2 % 8) == 0 && !getenv("VT_Z4_NO_TMA_STORE")) ? 1 : 0;
    Q.tiles_fast = (Q.o_fast + TS - 1) / TS;
    Q.tsy = z4_tile_rows(interp);
    for (int k = 0; k < P.n_mats; k++) {
        const VtMat &M = P.mats[k];
        Z4Mat &Z = Q.mats[k];
        Z.ys = M.r[R.ay][R.slow]; Z.yf = M.r[R.ay][R.fast]; Z.yc = M.r[R.ay][3];
        Z.xs = M.r[R.ax][R.slow]; Z.xf = M.r[R.ax][R.fast]; Z.xc = M.r[R.ax][3];
        Z.tm = (int)M.r[m][3];
        ext_y[k] = fabsf(Z.ys) * (float)(Q.tsy - 1) + fabsf(Z.yf) * (float)(TS - 1);
        ext_x[k] = fabsf(Z.xs) * (float)(Q.tsy - 1) + fabsf(Z.xf) * (float)(TS - 1);
    }
}

template <int INTERP>
int launch1(const VtResampleParams &P, const float *d_src4, int m, cudaStream_t st)
{
    Z4Params Q;
    float ext_y[VT_MAX_BATCH], ext_x[VT_MAX_BATCH];
    fill_z4(P, m, INTERP, Q, ext_y, ext_x);
    const bool zero = (P.flags & VT_OOB_ZERO) != 0;
    if (INTERP != VT_CUBIC_SIMPLE && (P.flags & VT_WEIGHTS_EXACT)) return launch2<INTERP, 2>(Q, d_src4, zero, ext_y, ext_x, st);
    return launch2<INTERP, 0>(Q, d_src4, zero, ext_y, ext_x, st);
}

}  // namespace

// the axis (0, 1, 2) every matrix of the batch leaves alone in the way this family needs, or -1.  Axis 0 first: it is
// the one whose layout the prefilter can write directly.
int vt_z4_axis(const VtResampleParams &P, int interp)
{
    (void)interp;
    const int sdim[3] = {P.s0, P.s1, P.s2};
    const long long ostr_lim = 0x7fffffffLL;
    if ((long long)P.o1 * P.o2 > ostr_lim) return -1;
    for (int m = 0; m < 3; m++) {
        bool ok = true;
        for (int k = 0; k < P.n_mats && ok; k++) ok = z4_mat_ok(P.mats[k], m, sdim);
        if (ok) return m;
    }
    return -1;
}

bool vt_z4_axis_accepts(const VtResampleParams &P, int axis)
{
    const int sdim[3] = {P.s0, P.s1, P.s2};
    if ((long long)P.o1 * P.o2 > 0x7fffffffLL || axis < 0 || axis > 2) return false;
    for (int k = 0; k < P.n_mats; k++)
        if (!z4_mat_ok(P.mats[k], axis, sdim)) return false;
    return true;
}

size_t vt_z4_floats(int s0, int s1, int s2, int axis)
{
    const int d[3] = {s0, s1, s2};
    const AxisRoles R = roles(axis);
    return (size_t)((d[axis] + 3) / 4) * (size_t)d[R.ay] * (size_t)d[R.ax] * 4;
}

int vt_pack_z4_impl(const float *d_src, int s0, int s1, int s2, long long row, long long plane, float *d_dst4, int axis,
                    cudaStream_t st)
{
    const int d[3] = {s0, s1, s2};
    const long long str[3] = {plane, row, 1};
    const AxisRoles R = roles(axis);
    const int groups = (d[axis] + 3) / 4;
    if (d[R.ay] > 65535 || groups > 65535) return VT_ERR_UNSUPPORTED;
    dim3 grid((d[R.ax] + 255) / 256, d[R.ay], groups);
    {
        VtProf prof(VT_K_PACK_Z4, st);
        VT_CUDA(vt_launch_pdl(vt_pack_z4_kernel, grid, dim3(256), 0, st, d_src, (float4 *)d_dst4, d[axis], d[R.ay], d[R.ax],
                              str[axis], str[R.ay], str[R.ax]));
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

// host-only: the per-matrix decisions (quarter shape, pitch class, simulated wavefronts per quarter-warp load)
int vt_z4_plan_impl(const VtResampleParams &P, int axis, int interp, int sms, int *chunks, int *m_chunk, int *box_w,
                    int *box_h, int *shapes, int *pitches, float *wavefronts)
{
    Z4Params Q;
    float ext_y[VT_MAX_BATCH], ext_x[VT_MAX_BATCH];
    fill_z4(P, axis, interp, Q, ext_y, ext_x);
    Z4Plan L;
    int rc;
    switch (interp) {
        case VT_LINEAR: rc = plan_z4<VT_LINEAR>(Q, sms, ext_y, ext_x, L); break;
        case VT_CUBIC_TEX: rc = plan_z4<VT_CUBIC_TEX>(Q, sms, ext_y, ext_x, L); break;
        case VT_CUBIC_SIMPLE: rc = plan_z4<VT_CUBIC_SIMPLE>(Q, sms, ext_y, ext_x, L); break;
        default: return VT_ERR_INVALID_ARG;
    }
    if (rc) return rc;
    *chunks = L.chunks;
    *m_chunk = L.m_chunk;
    *box_w = L.bw0;
    *box_h = L.bh;
    for (int k = 0; k < P.n_mats; k++) {
        if (shapes) shapes[k] = Q.mats[k].shape;
        if (pitches) pitches[k] = L.bw0 + Q.mats[k].pitch_idx;
        if (wavefronts) wavefronts[k] = z4_conflicts(Q.mats[k], Q.prod_on_fast, Q.tsy).wf[(L.bw0 + Q.mats[k].pitch_idx) & 7][Q.mats[k].shape];
    }
    return VT_OK;
}

int vt_launch_z4(const VtResampleParams &P, const float *d_src4, int axis, int interp, cudaStream_t st)
{
    if (P.z_end <= P.z_begin || P.o0 <= 0 || P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    if (axis < 0 || axis > 2 || ((uintptr_t)d_src4 % 16) != 0) return VT_ERR_INVALID_ARG;
    switch (interp) {
        case VT_LINEAR: return launch1<VT_LINEAR>(P, d_src4, axis, st);
        case VT_CUBIC_TEX: return launch1<VT_CUBIC_TEX>(P, d_src4, axis, st);
        case VT_CUBIC_SIMPLE: return launch1<VT_CUBIC_SIMPLE>(P, d_src4, axis, st);
    }
    return VT_ERR_INVALID_ARG;
}
