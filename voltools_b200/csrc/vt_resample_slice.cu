// vt_resample_slice.cu -- "slice" kernel family: transforms that leave axis 0 alone.
//
// When the matrix maps output axis 0 straight onto input axis 0 with an integer offset
//        M[0] = (1, 0, 0, t0),  t0 integer,  M[1][0] = M[2][0] = 0
// (every rotation about axis 0 through the centre of the volume: BASELINE configs[0..2], the reference's
// README sweep `rotation=(0, i, 0)` with 'rzxz'), the sample point of output voxel (a0, a1, a2) is
// (a0 + t0, q1(a1,a2), q2(a1,a2)): the in-plane position and therefore every in-plane interpolation weight is
// the same for the whole column a0 = 0..o0-1, and the fraction along axis 0 is exactly 0.  The kernel exploits
// that (the general kernels cannot):
//   * a CTA owns a 16x16 tile of (a1, a2) columns and marches along axis 0;
//   * the in-plane footprint of the tile (its rotated bounding box + filter support) is the same rectangle in
//     every input plane; it is staged into a shared-memory ring of 3-4 stages of 2-4 planes each, by TMA box loads
//     (cp.async.bulk.tensor: texels outside the source arrive as zeros = the texture's border mode) when the source
//     rows are 16-byte aligned, else by per-element cp.async with per-thread offsets computed once;
//   * weights are computed once per column with the reference's exact float32 recipe and kept in registers;
//   * per plane a thread does 4 (linear) or 16 (cubic) shared-memory loads, two planes per FFMA2; the three axis-0
//     taps of the cubic modes come from a register sliding window over the per-plane sums;
//   * the ring loop is unrolled over its stages and the cubic modes keep planes at compile-time strides, so a tap
//     load is LDS [register + immediate];
//   * how a warp's lanes tile the CTA's columns (2x16 / 4x8 / 8x4) and the row pitch are chosen per matrix by the
//     host from a bank simulation (lane_pos, pitch_cost), the length of a march from the plane size (L2 reuse
//     between neighbouring tiles).
// The same per-column weight code serves the fused rotate-and-project path at the end of this file.
// Results: same arithmetic as the gather family up to float32 summation order (<= ~1e-7 of the range).
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

constexpr int TS = 16;            // tile edge in (a1, a2)
constexpr int NT = TS * TS;       // threads per CTA, one column each
constexpr int BMAX = 32;          // max footprint edge (texels)
constexpr int PITCH_MIN = 32, PITCH_MAX = 40;  // shared-memory row pitch, chosen per matrix by the host
constexpr int STAGE = BMAX * PITCH_MAX;        // floats per staged plane, cp.async variant (5 KB)
// Ring: NSTAGE stages of PPS planes each (per interpolator, see Taps<>).  One stage = one mbarrier round trip, one
// __syncthreads and one TMA box of depth PPS, so the per-stage overhead (~45 instructions per thread) is shared by
// PPS planes; the planes of a stage are also paired in FFMA2 instructions.  A plane step is far shorter than the
// HBM latency, so several planes must be in flight per CTA: (NSTAGE-1)*PPS planes x 2-4 resident CTAs.
constexpr int MAX_NSTAGE = 8;
constexpr int EPT = (BMAX * BMAX + NT - 1) / NT;  // footprint elements per thread (4)

// 4-byte async copy global -> shared; `take == 0` or `plane_ok == 0` writes a zero instead (ignore-src form: no
// global access is made), which is how the texture's border mode is reproduced while staging
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const void *gmem_src, unsigned take, unsigned plane_ok)
{
    asm volatile(
        "{\n .reg .pred p;\n setp.eq.u32 p, %2, 0;\n setp.eq.or.u32 p, %3, 0, p;\n"
        " cp.async.ca.shared.global [%0], [%1], 4, p;\n}\n" ::"r"(smem_dst),
        "l"(gmem_src), "r"(take), "r"(plane_ok)
        : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// How a warp's 32 lanes tile the CTA's 16 x 16 columns: 2 x 16, 4 x 8 or 8 x 4 (rows x columns).  Lanes land on bank
// (y*pitch + x) mod 32 of their footprint origin and the pitch is a multiple of 4 (TMA rows are multiples of 16 bytes),
// so which shape spreads a warp over the most banks depends on the matrix: at 0 and 90 degrees 2 x 16 puts its two rows
// on the same 16 banks (2 wavefronts per load) where 4 x 8 with pitch = 8 mod 32 / 4 mod 32 is conflict free.  The host
// picks the pitch (per launch when it is the TMA box width) and the shape (per matrix, P.aux bits 6-7) from a bank
// simulation (pitch_cost); over a 180-angle sweep that is 1.66 wavefronts
// per load instead of 1.96 (45 degrees stays at 1.97 under every shape and pitch).
constexpr int N_LAYOUTS = 3;
__host__ __device__ __forceinline__ void lane_pos(int tid, int layout, int &ty, int &tx)
{
    const int lane = tid & 31, warp = tid >> 5;
    if (layout == 0) {
        ty = tid >> 4;
        tx = tid & 15;
    } else if (layout == 1) {
        ty = 4 * (warp >> 1) + (lane >> 3);
        tx = 8 * (warp & 1) + (lane & 7);
    } else {
        ty = 8 * (warp >> 2) + (lane >> 2);
        tx = 4 * (warp & 3) + (lane & 3);
    }
}

// in-plane coordinate of a column, reference recipe (transforms.py:264-274) with the a0 term dropped: it is an
// exact no-op because M[r][0] == 0 (fma(a0, 0, t) == t).
__host__ __device__ __forceinline__ float inplane_coord(const float *row, float a1, float a2)
{
#ifdef __CUDA_ARCH__
    float t = __fmul_rn(a1, row[1]);
    t = __fmaf_rn(a2, row[2], t);
    t = __fadd_rn(row[3], t);
    return __fadd_rn(t, 0.5f);
#else
    float t = a1 * row[1];
    t = fmaf(a2, row[2], t);
    t = row[3] + t;
    return t + 0.5f;
#endif
}

// planes per ring stage of the cubic kernels (even; the ring then has 4 stages of 2 planes or 3 stages of 4).
// Measured: 4 planes per stage (half the barriers, more loads in flight, but 76 / 80+spill registers) is 6-10 % SLOWER
// than 2 (cubic_tex 250^3: 0.087 vs 0.082 ms; 256^3 12-angle sums 1.113 vs 1.001 ms).
#ifndef VT_CUBIC_PPS
#define VT_CUBIC_PPS 2
#endif

template <int INTERP>
struct Taps;

// ---- linear: one plane, 4 taps with the texture unit's integer weights (c = 0 -> S = 256) ----------------
template <>
struct Taps<VT_LINEAR> {
    static constexpr int LO = 0, HI = 2;  // footprint margins relative to floor(p - 0.5)
    static constexpr int PLANES_BEFORE = 0, PLANES_AFTER = 0;
    static constexpr int PPS = 4, NSTAGE = 3;
    float w[4];
    int r0, r1;  // element offsets of the two tap rows within a stage
    template <int RULE>
    __device__ __forceinline__ void init(float p1, float p2, int ylo, int xlo, int pitch)
    {
        int by, bx;
        if (RULE == 0) {
            int a, b;
            vt_tex_fix_hw(p2, bx, a);
            vt_tex_fix_hw(p1, by, b);
            int wi[4];
            vt_tex_hw_side(a, b, 256, wi);
#pragma unroll
            for (int k = 0; k < 4; k++) w[k] = vt_u2f(wi[k]) * (1.0f / 256.0f);
        } else {
            float ax, ay;
            vt_tex_fix<2>(p2, bx, ax);
            vt_tex_fix<2>(p1, by, ay);
            w[0] = (1.0f - ax) * (1.0f - ay);
            w[1] = ax * (1.0f - ay);
            w[2] = (1.0f - ax) * ay;
            w[3] = ax * ay;
        }
        r0 = (by - ylo) * pitch + (bx - xlo);
        r1 = r0 + pitch;
    }
    // per-plane sums of the stage's four planes (plane p starts at s + p*pe); planes are paired in FFMA2s
    __device__ __forceinline__ void planes(const float *s, int pe, float (&out)[PPS]) const
    {
        const float *s1 = s + pe, *s2 = s1 + pe, *s3 = s2 + pe;
        vt_f2 a01 = 0ull, a23 = 0ull;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = (k < 2 ? r0 : r1) + (k & 1);
            const vt_f2 ww = vt_pk(w[k], w[k]);
            a01 = vt_fma2(vt_pk(s[o], s1[o]), ww, a01);
            a23 = vt_fma2(vt_pk(s2[o], s3[o]), ww, a23);
        }
        vt_unpk(a01, out[0], out[1]);
        vt_unpk(a23, out[2], out[3]);
    }
};

// ---- cubic_simple: 16 in-plane taps, float32 B-spline weights; axis-0 weights B(-1), B(0), B(1) ------------
template <>
struct Taps<VT_CUBIC_SIMPLE> {
    static constexpr int LO = -1, HI = 2;
    static constexpr int PLANES_BEFORE = 1, PLANES_AFTER = 1;
    static constexpr int PPS = VT_CUBIC_PPS, NSTAGE = VT_CUBIC_PPS == 2 ? 4 : 3;
    float w[16];
    int row[4];
    float wz0, wz1, wz2;
    template <int RULE>
    __device__ __forceinline__ void init(float p1, float p2, int ylo, int xlo, int pitch)
    {
        const float cgx = __fadd_rn(p2, -0.5f), cgy = __fadd_rn(p1, -0.5f);
        const float fx0 = floorf(cgx), fy0 = floorf(cgy);
        const float fx = __fsub_rn(cgx, fx0), fy = __fsub_rn(cgy, fy0);
        float wx[4], wy[4];
        vt_bspline4(fx, wx);
        vt_bspline4(fy, wy);
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) w[j * 4 + i] = __fmul_rn(wx[i], wy[j]);
        row[0] = ((int)fy0 - 1 - ylo) * pitch + ((int)fx0 - 1 - xlo);
#pragma unroll
        for (int j = 1; j < 4; j++) row[j] = row[j - 1] + pitch;
        wz0 = vt_bspline(-1.0f);  // fraction along axis 0 is exactly 0
        wz1 = vt_bspline(0.0f);
        wz2 = vt_bspline(1.0f);
    }
    // in-plane sums of the stage's two planes, one FFMA2 per tap: {plane q, plane q+1} x {w, w}
    __device__ __forceinline__ void planes(const float *s, int pe, float (&out)[PPS]) const
    {
        vt_f2 acc[PPS / 2];
#pragma unroll
        for (int h = 0; h < PPS / 2; h++) acc[h] = 0ull;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const vt_f2 ww = vt_pk(w[j * 4 + i], w[j * 4 + i]);
#pragma unroll
                for (int h = 0; h < PPS / 2; h++)
                    acc[h] = vt_fma2(vt_pk(s[(2 * h) * pe + row[j] + i], s[(2 * h + 1) * pe + row[j] + i]), ww, acc[h]);
            }
#pragma unroll
        for (int h = 0; h < PPS / 2; h++) vt_unpk(acc[h], out[2 * h], out[2 * h + 1]);
    }
};

// ---- cubic_tex: Ruijters' 8 trilinear fetches with the texture unit's integer weights ---------------------
// With fraction 0 along axis 0: h0z = idx + 0.3 -> texels idx-1 (S = 256-c0) and idx (S = c0), combined with g0z;
// h1z = idx + 1.5 -> texel idx+1 (S = 256-c1 = 256), combined with g1z (texel idx+2 has S = c1 = 0).
// Each of the three planes gets its own set of 16 pre-multiplied tap weights.
template <>
struct Taps<VT_CUBIC_TEX> {
    static constexpr int LO = -1, HI = 2;
    static constexpr int PLANES_BEFORE = 1, PLANES_AFTER = 1;
    static constexpr int PPS = VT_CUBIC_PPS, NSTAGE = VT_CUBIC_PPS == 2 ? 4 : 3;
    float wa[16], wb[16], wc[16];
    int adr[8];  // element offsets of taps (row j, x pair k): adr[j*2+k], the pair is (adr, adr+1)
    template <int RULE>
    __device__ __forceinline__ void init(float p1, float p2, int ylo, int xlo, int pitch)
    {
        float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
        vt_ruijters(p2, g0x, g1x, h0x, h1x);
        vt_ruijters(p1, g0y, g1y, h0y, h1y);
        vt_ruijters(8.5f, g0z, g1z, h0z, h1z);  // any texel centre: the fraction along axis 0 is exactly 0
        const float gx[2] = {g0x, g1x}, gy[2] = {g0y, g1y};
        const float hx[2] = {h0x, h1x}, hy[2] = {h0y, h1y};
        const float gz[3] = {g0z, g0z, g1z};
        int bx[2], by[2];
        if (RULE == 0) {
            int bz0, c0, bz1, c1;
            vt_tex_fix_hw(h0z, bz0, c0);  // (7, 205)
            vt_tex_fix_hw(h1z, bz1, c1);  // (9, 0)
            const int S[3] = {256 - c0, c0, 256 - c1};
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
                        int wi[4];
                        vt_tex_hw_side(ax[i], ay[j], S[p], wi);
                        const float g = __fmul_rn(__fmul_rn(gx[i], gy[j]), gz[p]) * (1.0f / 256.0f);
                        w[(2 * j) * 4 + 2 * i] = g * vt_u2f(wi[0]);
                        w[(2 * j) * 4 + 2 * i + 1] = g * vt_u2f(wi[1]);
                        w[(2 * j + 1) * 4 + 2 * i] = g * vt_u2f(wi[2]);
                        w[(2 * j + 1) * 4 + 2 * i + 1] = g * vt_u2f(wi[3]);
                    }
            }
        } else {
            // exact float32 weights: alphas are the exact fractions of h - 0.5
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
                adr[(2 * j) * 2 + k] = (by[j] - ylo) * pitch + (bx[k] - xlo);
                adr[(2 * j + 1) * 2 + k] = adr[(2 * j) * 2 + k] + pitch;
            }
    }
    // A-, B- and C-weighted sums of the stage's two planes: per tap {wa, wb} x {t, t} for each plane (the scalar
    // texel is FFMA2's broadcast operand) and {t0, t1} x {wc, wc}: 3 FFMA2 instead of 6 FFMA
    __device__ __forceinline__ void planes3(const float *s, int pe, float (&qa)[PPS], float (&qb)[PPS], float (&qc)[PPS]) const
    {
        vt_f2 ab[PPS], c[PPS / 2];
#pragma unroll
        for (int p = 0; p < PPS; p++) ab[p] = 0ull;
#pragma unroll
        for (int h = 0; h < PPS / 2; h++) c[h] = 0ull;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int o = adr[j * 2 + (i >> 1)] + (i & 1);
                const vt_f2 wab = vt_pk(wa[j * 4 + i], wb[j * 4 + i]);
                const vt_f2 wcc = vt_pk(wc[j * 4 + i], wc[j * 4 + i]);
#pragma unroll
                for (int h = 0; h < PPS / 2; h++) {
                    const float t0 = s[(2 * h) * pe + o], t1 = s[(2 * h + 1) * pe + o];
                    ab[2 * h] = vt_fma2(wab, vt_pk(t0, t0), ab[2 * h]);
                    ab[2 * h + 1] = vt_fma2(wab, vt_pk(t1, t1), ab[2 * h + 1]);
                    c[h] = vt_fma2(vt_pk(t0, t1), wcc, c[h]);
                }
            }
#pragma unroll
        for (int p = 0; p < PPS; p++) vt_unpk(ab[p], qa[p], qb[p]);
#pragma unroll
        for (int h = 0; h < PPS / 2; h++) vt_unpk(c[h], qc[2 * h], qc[2 * h + 1]);
    }
};

// Launch-wide staging parameters.  TMA variant: one tensor map over the source volume (x fastest), every plane
// of a tile's footprint is ONE cp.async.bulk.tensor box load of box_w x box_h texels (coordinates may be negative
// or past the end: those texels arrive as zeros = the texture's border mode, for whole planes too).
// Measured on B200: the box must START on a 16-byte boundary of the row (x coordinate a multiple of 4 floats),
// otherwise the load faults ("illegal instruction"); so the box start is rounded down and the box is 3 wider.
struct VtSliceStaging {
    CUtensorMap tmap;
    int box_w, box_h;      // TMA box (box_w is also the shared-memory row pitch of the TMA variant); box depth = PPS
    int plane_elems;       // floats between consecutive planes of a stage
    unsigned stage_bytes;  // ring stage size in bytes (multiple of 128)
};

template <int INTERP, int RULE, bool OOB_ZERO, bool TMA>
__global__ void __launch_bounds__(NT, INTERP == VT_LINEAR ? 4 : 3)
    vt_slice_kernel(const __grid_constant__ VtResampleParams P, const __grid_constant__ VtSliceStaging G, int z_chunk)
{
    // [128 B of mbarriers: "full" 0..7 (TMA bytes landed), "empty" 8..15 (every warp is done reading)][NSTAGE stages]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned long long *bars = (unsigned long long *)smem_raw;
    unsigned char *ring = smem_raw + 128;
    using T = Taps<INTERP>;
    constexpr int PPS = T::PPS, NSTAGE = T::NSTAGE;
    static_assert(NSTAGE <= MAX_NSTAGE, "mbarrier block holds MAX_NSTAGE barriers");
    static_assert(NSTAGE == 3 || NSTAGE == 4, "the ring loop is unrolled for 3 or 4 stages");
    const int tid = threadIdx.x;
    const int ntx = (P.o2 + TS - 1) / TS;
    const int tile_y = blockIdx.x / ntx, tile_x = blockIdx.x - tile_y * ntx;
    const int mat = blockIdx.z;
    const VtMat &M = P.mats[mat];
    const int pitch = TMA ? G.box_w : (int)(P.aux[mat] & 63);
    const int layout = (int)(P.aux[mat] >> 6);
    // Cubic modes: planes and ring stages sit at COMPILE-TIME strides (STAGE floats per plane) and the ring loop is
    // unrolled over the stages, so every tap load is LDS [per-thread register + immediate]: no address arithmetic per
    // stage (with run-time strides each stage cost ~24 LEA / MOV instructions next to its 48 FFMA2 + 32 LDS, and
    // cubic_tex sits on its issue rate).  linear keeps the densely packed box (4 planes per stage: a fixed stride would
    // cost a fourth resident CTA) and only gets the unrolled ring.
    constexpr bool FIXED = INTERP != VT_LINEAR;
    const unsigned stage_bytes = FIXED ? (unsigned)(PPS * STAGE * 4) : G.stage_bytes;
    const int pe = FIXED ? STAGE : G.plane_elems;
    const int t0 = (int)M.r[0][3];
    const int zc0 = P.z_begin + blockIdx.y * z_chunk;
    const int zc1 = min(zc0 + z_chunk, P.z_end);
    const int a1_0 = tile_y * TS, a2_0 = tile_x * TS;
    const int a1_1 = min(a1_0 + TS, P.o1) - 1, a2_1 = min(a2_0 + TS, P.o2) - 1;

    // footprint of the tile: extremes are at the corners (the float recipe is monotone in a1 and in a2)
    float y_min, y_max, x_min, x_max;
    {
        const float c1a = (float)a1_0, c1b = (float)a1_1, c2a = (float)a2_0, c2b = (float)a2_1;
        const float y00 = inplane_coord(M.r[1], c1a, c2a), y01 = inplane_coord(M.r[1], c1a, c2b);
        const float y10 = inplane_coord(M.r[1], c1b, c2a), y11 = inplane_coord(M.r[1], c1b, c2b);
        const float x00 = inplane_coord(M.r[2], c1a, c2a), x01 = inplane_coord(M.r[2], c1a, c2b);
        const float x10 = inplane_coord(M.r[2], c1b, c2a), x11 = inplane_coord(M.r[2], c1b, c2b);
        y_min = fminf(fminf(y00, y01), fminf(y10, y11));
        y_max = fmaxf(fmaxf(y00, y01), fmaxf(y10, y11));
        x_min = fminf(fminf(x00, x01), fminf(x10, x11));
        x_max = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11));
    }
    // clamp the footprint to a couple of texels around the source: everything further out is border (zero)
    // anyway and columns that far out are out of bounds; keeps the integer conversions safe for wild matrices
    y_min = fmaxf(y_min, -2.0f); x_min = fmaxf(x_min, -2.0f);
    y_max = fminf(y_max, (float)P.s1 + 2.0f); x_max = fminf(x_max, (float)P.s2 + 2.0f);
    const int ylo = (int)floorf(y_min - 0.5f) + T::LO, yhi = (int)floorf(y_max - 0.5f) + T::HI;
    int xlo = (int)floorf(x_min - 0.5f) + T::LO;
    const int xhi = (int)floorf(x_max - 0.5f) + T::HI;
    // TMA: the box must start on a 16-byte boundary of the row (the host widened the box by 3 texels for this)
    if (TMA) xlo &= ~3;
    int bh = min(yhi - ylo + 1, BMAX), bw = min(xhi - xlo + 1, BMAX);  // host guarantees <= BMAX (<= the TMA box)
    if (bh <= 0 || bw <= 0) bh = bw = 0;  // tile entirely outside the source: nothing to stage

    // cp.async variant: per-thread share of the footprint, constant along the march
    const unsigned ring_s = vt_smem_u32(ring);
    const unsigned bars_s = vt_smem_u32(bars);
    unsigned sdst[EPT], goff[EPT], gsz[EPT];  // smem byte address in stage 0, byte offset in a plane, 4 or 0
    const int nel = bh * bw;
    const int kmax = (nel + NT - 1) / NT;     // uniform
    if constexpr (!TMA) {
#pragma unroll
        for (int k = 0; k < EPT; k++) {
            const int e = tid + k * NT;
            const int r = bw > 0 ? e / bw : 0, c = e - r * bw;
            const int y = ylo + r, x = xlo + c;
            const bool in = e < nel;
            const bool val = in && (unsigned)y < (unsigned)P.s1 && (unsigned)x < (unsigned)P.s2;
            // elements past the footprint are parked on the last float of the plane (never read)
            sdst[k] = ring_s + (in ? 4u * (unsigned)(r * pitch + c) : 4u * (unsigned)pe - 4u);
            gsz[k] = val ? 4u : 0u;
            goff[k] = val ? 4u * (unsigned)(y * (int)P.src_row + x) : 0u;
        }
    } else {
        if (tid == 0) {
            vt_tma_prefetch_desc(&G.tmap);
#pragma unroll
            for (int i = 0; i < NSTAGE; i++) {
                vt_mbar_init(bars_s + 8u * i, 1);
                vt_mbar_init(bars_s + 8u * (MAX_NSTAGE + i), NT / 32);  // one arrival per warp
            }
            vt_mbar_fence_init();
        }
        __syncthreads();
    }
    const size_t plane_bytes = (size_t)P.src_plane * sizeof(float);

    // this thread's column
    int ty, tx;
    lane_pos(tid, layout, ty, tx);
    const int a1 = a1_0 + ty, a2 = a2_0 + tx;
    const bool live = a1 < P.o1 && a2 < P.o2;
    const float p1 = inplane_coord(M.r[1], (float)a1, (float)a2);
    const float p2 = inplane_coord(M.r[2], (float)a1, (float)a2);
    // transforms.py:276-278 for the two in-plane axes
    const bool inplane = live && !(p2 < 0 || p1 < 0 || p2 >= (float)P.s2 || p1 >= (float)P.s1);
    const size_t oplane = (size_t)P.o1 * P.o2;

    // input planes needed: q = z + t0 + d, d in [-PLANES_BEFORE, PLANES_AFTER]
    const int q_first = zc0 + t0 - T::PLANES_BEFORE, q_last = zc1 - 1 + t0 + T::PLANES_AFTER;
    // running pointers: input plane q (may point outside the source for out-of-range q: never dereferenced then)
    const char *srcq = (const char *)P.src + (long long)q_first * (long long)plane_bytes;
    float *dstz = P.dst + (size_t)mat * P.dst_batch_stride + ((size_t)a1 * P.o2 + a2) +
                  (long long)(q_first - t0 - T::PLANES_AFTER) * (long long)oplane;

    // stage input planes q .. q+PPS-1 (gq = address of plane q) into ring stage `st`
    auto issue = [&](int q, const char *gq, unsigned st) {
        if constexpr (TMA) {
            if (tid == 0 && q <= q_last) {
                const unsigned bar = bars_s + 8u * st;
                vt_mbar_expect_tx(bar, (unsigned)(G.box_w * G.box_h * PPS) * 4u);
                if constexpr (FIXED) {  // one box of depth 1 per plane, at the fixed plane stride
#pragma unroll
                    for (int p = 0; p < PPS; p++)
                        vt_tma_load_3d(ring_s + st * stage_bytes + 4u * (unsigned)(p * STAGE), &G.tmap, bar, xlo, ylo, q + p);
                } else {
                    vt_tma_load_3d(ring_s + st * stage_bytes, &G.tmap, bar, xlo, ylo, q);  // planes past the source: zeros
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < PPS; p++) {
                if (q + p <= q_last) {
                    const unsigned zin = (unsigned)(q + p) < (unsigned)P.s0 ? 1u : 0u;            // uniform
                    const char *g = zin ? gq + (size_t)p * plane_bytes : (const char *)P.src;      // keep the address in bounds
                    const unsigned sb = st * stage_bytes + 4u * (unsigned)(p * pe);
#pragma unroll
                    for (int k = 0; k < EPT; k++)
                        if (k < kmax) cp_async4(sdst[k] + sb, g + goff[k], gsz[k], zin);
                }
            }
            cp_async_commit();
        }
    };

    // prologue: NSTAGE-1 stages in flight
#pragma unroll
    for (int i = 0; i < NSTAGE - 1; i++) issue(q_first + i * PPS, srcq + (size_t)(i * PPS) * plane_bytes, (unsigned)i);
    // the column's weights are computed while those first loads are in flight
    T taps;
    if (inplane) taps.template init<RULE>(p1, p2, ylo, xlo, pitch);
    float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // sliding window of per-plane sums
    unsigned phase = 0;                      // mbarrier parity of the ring's current round
    const char *srcf = srcq + (size_t)((NSTAGE - 1) * PPS) * plane_bytes;  // address of plane q + (NSTAGE-1)*PPS

    constexpr bool LOOSE = TMA && INTERP == VT_LINEAR;  // full/empty mbarrier ring without a block-wide barrier
    bool first_round = true;
    // one ring stage: planes q .. q+PPS-1 live in stage CUR (compile time); refills the stage consumed before it
    auto stage_step = [&](auto cur_c, int q) {
        constexpr unsigned CUR = decltype(cur_c)::value;
        constexpr unsigned FILL = (CUR + NSTAGE - 1) % NSTAGE;
        if constexpr (LOOSE) {
            // No block-wide barrier in the linear kernel's TMA ring: every thread waits for its stage's bytes on the
            // "full" barrier, and the one thread that refills a stage first waits on that stage's "empty" barrier, on
            // which each warp arrives once it has read the stage.  The refilled stage is the one consumed in the previous
            // iteration, so only that thread's warp ever waits for the slowest warp; the others run ahead through the
            // stages in flight.  Measured at 512^3, 45 degrees: linear 0.218 -> 0.213 ms; the cubic kernels got SLOWER
            // (cubic_tex 0.535 -> 0.589 ms: the refilling warp becomes the straggler of a compute-heavy stage), so they
            // keep the barrier.
            if (tid == 0 && q + (NSTAGE - 1) * PPS <= q_last && !(CUR == 0 && first_round))
                vt_mbar_wait(bars_s + 8u * (MAX_NSTAGE + FILL), CUR == 0 ? (phase ^ 1u) : phase);
            issue(q + (NSTAGE - 1) * PPS, srcf, FILL);
            vt_mbar_wait(bars_s + 8u * CUR, phase);
        } else {
            if constexpr (TMA) vt_mbar_wait(bars_s + 8u * CUR, phase);
            else cp_async_wait<NSTAGE - 2>();
            __syncthreads();  // stage CUR has landed for every thread; everyone is done with the previous stage
            issue(q + (NSTAGE - 1) * PPS, srcf, FILL);
        }
        const float *s = (const float *)(ring + CUR * stage_bytes);
        float r[PPS];
#pragma unroll
        for (int p = 0; p < PPS; p++) r[p] = 0.0f;
        if (inplane) {
            if constexpr (INTERP == VT_LINEAR) {
                taps.planes(s, pe, r);
            } else if constexpr (INTERP == VT_CUBIC_SIMPLE) {
                float pq[PPS];
                taps.planes(s, pe, pq);
#pragma unroll
                for (int p = 0; p < PPS; p++) {
                    // the reference accumulates kz = -1, 0, 1 in that order (helper_interpolation.h:51)
                    r[p] = fmaf(taps.wz2, pq[p], fmaf(taps.wz1, s1, __fmul_rn(taps.wz0, s2)));
                    s2 = s1;
                    s1 = pq[p];
                }
            } else {
                float qa[PPS], qb[PPS], qc[PPS];
                taps.planes3(s, pe, qa, qb, qc);
#pragma unroll
                for (int p = 0; p < PPS; p++) {
                    r[p] = (s3 + s1) + qc[p];  // s3 = A-sum of plane q-2, s1 = B-sum of plane q-1
                    s3 = s2;                   // s2 = A-sum of plane q-1
                    s2 = qa[p];
                    s1 = qb[p];
                }
            }
        }
#pragma unroll
        for (int p = 0; p < PPS; p++) {
            const int zi = q + p - T::PLANES_AFTER;            // input plane at the centre of the output voxel
            const int zo = zi - t0;                            // output plane
            if (zo >= zc0 && zo < zc1) {                       // uniform: past the warm-up planes, inside the chunk
                const bool ok = inplane && (unsigned)zi < (unsigned)P.s0;  // 0 <= p0 < s0 with p0 = zi + 0.5
                if (OOB_ZERO) {
                    if (live) *dstz = ok ? r[p] : 0.0f;
                } else if (ok) {
                    *dstz = r[p];
                }
            }
            dstz += oplane;
        }
        if constexpr (LOOSE) {
            __syncwarp();
            if ((tid & 31) == 0) vt_mbar_arrive(bars_s + 8u * (MAX_NSTAGE + CUR));  // this warp is done with stage CUR
        }
        srcf += (size_t)PPS * plane_bytes;
    };
    for (int q = q_first;;) {
        stage_step(std::integral_constant<unsigned, 0>{}, q);
        if ((q += PPS) > q_last) break;
        stage_step(std::integral_constant<unsigned, 1>{}, q);
        if ((q += PPS) > q_last) break;
        stage_step(std::integral_constant<unsigned, 2>{}, q);
        if ((q += PPS) > q_last) break;
        if constexpr (NSTAGE == 4) {
            stage_step(std::integral_constant<unsigned, 3>{}, q);
            if ((q += PPS) > q_last) break;
        }
        phase ^= 1u;
        first_round = false;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// can this batch run on the slice family?
bool slice_ok(const VtResampleParams &P, int interp)
{
    if (P.s0 >= 16384) return false;  // h0z = idx + 0.3 must keep its 1/256 quantum (see Taps<VT_CUBIC_TEX>)
    if (P.src_plane >= (1ll << 29)) return false;  // 32-bit byte offsets inside a plane
    for (int k = 0; k < P.n_mats; k++) {
        const VtMat &M = P.mats[k];
        if (M.r[0][0] != 1.0f || M.r[0][1] != 0.0f || M.r[0][2] != 0.0f) return false;
        if (M.r[1][0] != 0.0f || M.r[2][0] != 0.0f) return false;
        const float t0 = M.r[0][3];
        if (!(fabsf(t0) < 16384.0f) || t0 != floorf(t0)) return false;
        for (int r = 1; r <= 2; r++) {
            const float ext = (fabsf(M.r[r][1]) + fabsf(M.r[r][2])) * (float)(TS - 1);
            if (!(ext <= (float)(BMAX - 5) - 0.1f)) return false;  // footprint edge <= ext + 5 texels
            if (!(fabsf(M.r[r][3]) < 1e6f)) return false;
        }
    }
    return true;
}

// Shared-memory row pitch and warp shape.  A warp's 32 columns read the same tap of their own footprint position; lanes
// land on bank (y*pitch + x) mod 32.  The host simulates the first tap of a few warps (every other tap shifts all lanes
// by the same amount) and feeds the average number of wavefronts per load into a cost model fitted to measurements at
// 512^3 over (angle, pitch, shape) (tools/slice_grid_probe.py, profiles/r01k_slice_grid.log), in SM cycles per voxel:
//      cubic_simple   0.30 + 0.45 * wavefronts          (shared-memory bound: 16 loads per voxel and plane)
//      cubic_tex      max(1.13, that)                   (1.13 = its instruction-issue floor)
//      linear         max(0.40, 0.19 + 0.0045 * pitch + 0.06 * wavefronts)   (L2 -> SM traffic grows with the box width)
//   + 0.02 (4 x 8) / 0.12 (8 x 4) measured: a warp's stores cover 4 or 8 rows instead of 2.
constexpr int PITCH_SAMPLES = 16;
constexpr int PITCH_LO = 12, PITCH_HI = PITCH_MAX;  // pitches the tables cover (bit p - PITCH_LO of the masks): the
                                                    // narrowest TMA box (strong magnification) is 12 texels wide
static_assert(PITCH_HI - PITCH_LO < 32, "one mask bit per pitch");

// Average wavefronts per load of a matrix for every (pitch, warp shape) asked for.  The footprint origins of the sample
// warps do not depend on the pitch, so they are computed once (the float coordinate recipe is the expensive part) and
// each pitch is then integer bank counting; results are memoised per matrix (a sweep comes back with the same matrices
// launch after launch, and a launch of 32 new matrices must stay well below the ~2 ms its kernels take).
struct ConflictTable {
    float key[6];
    unsigned valid;  // pitches already filled
    float wf[PITCH_HI - PITCH_LO + 1][N_LAYOUTS];
};
const ConflictTable &conflict_table(const VtMat &M, unsigned need)
{
    static thread_local ConflictTable memo[512];
    const float key[6] = {M.r[1][1], M.r[1][2], M.r[1][3], M.r[2][1], M.r[2][2], M.r[2][3]};
    unsigned h = 2166136261u;
    for (int i = 0; i < 6; i++) {
        unsigned u;
        memcpy(&u, &key[i], 4);
        h = (h ^ u) * 16777619u;
    }
    ConflictTable &e = memo[(h >> 7) & 511];
    if (memcmp(e.key, key, sizeof key) != 0) {
        memcpy(e.key, key, sizeof key);
        e.valid = 0;
    }
    const unsigned todo = need & ~e.valid;
    if (!todo) return e;
    short oy[N_LAYOUTS][PITCH_SAMPLES][32], ox[N_LAYOUTS][PITCH_SAMPLES][32];
    for (int layout = 0; layout < N_LAYOUTS; layout++)
        for (int sample = 0; sample < PITCH_SAMPLES; sample++) {
            // tiles spread over the image (different sub-texel phases), a different warp of the CTA in each
            // (irregular positions: an arithmetic progression of tiles resonates with the lattice at some angles)
            static const unsigned char T1[16] = {1, 14, 5, 27, 9, 2, 19, 30, 11, 23, 4, 16, 7, 21, 13, 26};
            static const unsigned char T2[16] = {3, 8, 29, 12, 1, 22, 17, 6, 25, 10, 31, 15, 20, 2, 28, 18};
            const int a1_0 = 16 * T1[sample & 15], a2_0 = 16 * T2[sample & 15];
            const float y0 = floorf(inplane_coord(M.r[1], (float)a1_0, (float)a2_0) - 0.5f);
            const float x0 = floorf(inplane_coord(M.r[2], (float)a1_0, (float)a2_0) - 0.5f);
            for (int lane = 0; lane < 32; lane++) {
                int ty, tx;
                lane_pos(32 * ((5 * sample + (sample >> 3)) & 7) + lane, layout, ty, tx);
                const int a1 = a1_0 + ty, a2 = a2_0 + tx;
                // relative to the tile's first column: small numbers whatever the translation is
                oy[layout][sample][lane] = (short)(floorf(inplane_coord(M.r[1], (float)a1, (float)a2) - 0.5f) - y0);
                ox[layout][sample][lane] = (short)(floorf(inplane_coord(M.r[2], (float)a1, (float)a2) - 0.5f) - x0);
            }
        }
    for (int pitch = PITCH_LO; pitch <= PITCH_HI; pitch++) {
        if (!(todo >> (pitch - PITCH_LO) & 1u)) continue;
        for (int layout = 0; layout < N_LAYOUTS; layout++) {
            int cost = 0;
            for (int sample = 0; sample < PITCH_SAMPLES; sample++) {
                unsigned char count[32] = {0};
                int first_addr[32];
                int worst = 0;
                for (int lane = 0; lane < 32; lane++) {
                    const int addr = oy[layout][sample][lane] * pitch + ox[layout][sample][lane] + (1 << 20);
                    const int bank = addr & 31;
                    if (count[bank] && first_addr[bank] == addr) continue;  // same word: broadcast
                    if (!count[bank]) first_addr[bank] = addr;
                    if (++count[bank] > worst) worst = count[bank];
                }
                cost += worst;
            }
            e.wf[pitch - PITCH_LO][layout] = (float)cost / (float)PITCH_SAMPLES;
        }
    }
    e.valid |= todo;
    return e;
}
float pitch_conflicts(const VtMat &M, int pitch, int layout)
{
    return conflict_table(M, 1u << (pitch - PITCH_LO)).wf[pitch - PITCH_LO][layout];
}
template <int INTERP>
float pitch_cost(const VtMat &M, int pitch, int layout)
{
    const float wf = pitch_conflicts(M, pitch, layout);
    // (a little more than measured -- 0.02 / 0.12 -- so that sampling noise in `wf` does not flip a tie away from 2 x 16)
    const float pen = layout == 0 ? 0.0f : (layout == 1 ? 0.05f : 0.15f);
    if (INTERP == VT_LINEAR) return fmaxf(0.40f, 0.19f + 0.0045f * (float)pitch + 0.06f * wf) + pen;
    const float lds = 0.30f + 0.45f * wf;
    return (INTERP == VT_CUBIC_TEX ? fmaxf(1.13f, lds) : lds) + pen;
}

// TMA staging is possible when the source rows/planes are 16-byte aligned
bool tma_ok(const VtResampleParams &P)
{
    return (P.src_row % 4) == 0 && (P.src_plane % 4) == 0 && ((uintptr_t)P.src % 16) == 0;
}

// Everything the host decides about a launch, without touching the device (also exported for inspection and for the
// CPU-side tests: vt_slice_plan).
struct SlicePlan {
    int chunks, z_chunk;  // z-chunks per column tile and their length in output planes
    bool tma;             // TMA box staging (16-byte aligned source rows) or per-element cp.async
    int box_w, box_h;     // TMA box = shared-memory pitch x rows (TMA variant)
    float cost;           // modelled SM cycles per voxel and plane, summed over the matrices (TMA variant)
    unsigned char aux[VT_MAX_BATCH];  // per matrix: warp shape << 6 | pitch (pitch only in the cp.async variant)
};

template <int INTERP>
int plan_slice(const VtResampleParams &P, int sms, SlicePlan &L)
{
    constexpr int PPS = Taps<INTERP>::PPS;
    const int nz = P.z_end - P.z_begin;
    const int tiles = ((P.o1 + TS - 1) / TS) * ((P.o2 + TS - 1) / TS);
    // z-chunks: a CTA marches z_chunk + WARM planes and pays a fixed start-up (weights, descriptor fetch, first
    // loads in flight) worth a few more; CTAs run in waves of (SMs x resident CTAs).  Pick the chunk count with
    // the smallest waves x march length -- for a single 250^3 volume (256 tiles) that is the difference between
    // 5 ragged waves and 3 full ones.
    constexpr int WARM = Taps<INTERP>::PLANES_BEFORE + Taps<INTERP>::PLANES_AFTER;
    constexpr int RESIDENT = INTERP == VT_LINEAR ? 4 : 3;  // __launch_bounds__ of the kernel
    constexpr int STARTUP = 8;
    const long long slots = (long long)sms * RESIDENT, per_chunk = (long long)tiles * P.n_mats;
    int chunks = 1, z_chunk = nz;
    long long best_chunk_cost = -1;
    // Long marches over large planes lose the L2 reuse between neighbouring tiles (their footprints overlap ~3.4x): CTAs
    // drift apart along z, and once the drift times the plane size exceeds the L2 the overlap is fetched from HBM again.
    // Measured at 1024^3 linear, 45 degrees (4 MB planes): 1 chunk of 1024 planes 3.49 ms, 2: 3.08, 4: 2.25, 8: 1.73,
    // 16: 1.75 ms.  So a march of the linear kernel covers at most 512 MB of source planes; the cubic kernels, bound on
    // chip, were indifferent at 1 chunk and lost 4-5 % to the extra start-ups with 8, so they get 2 GB.
    const long long plane_bytes = (long long)P.src_plane * 4;
    const long long march_bytes = INTERP == VT_LINEAR ? (512LL << 20) : (2048LL << 20);
    const int march_cap = (int)std::max(32LL, march_bytes / std::max(plane_bytes, 1LL));
    for (int c = 1; c <= 256 && (c == 1 || nz / c >= 16); c++) {
        int zc = (nz + c - 1) / c;
        if (c > 1) zc = (zc + WARM + PPS - 1) / PPS * PPS - WARM;  // a whole number of stages per chunk
        if (zc < 1) break;
        if (zc > march_cap && nz / (c + 1) >= 16) continue;
        const int cc = (nz + zc - 1) / zc;
        const long long waves = (per_chunk * cc + slots - 1) / slots;
        const long long cost = waves * (zc + WARM + STARTUP);
        if (best_chunk_cost < 0 || cost < best_chunk_cost) {
            best_chunk_cost = cost;
            chunks = cc;
            z_chunk = zc;
        }
    }
    if (const char *e = getenv("VT_SLICE_CHUNKS")) {  // tuning knob
        if (atoi(e) > 0) {
            z_chunk = (nz + atoi(e) - 1) / atoi(e);
            chunks = (nz + z_chunk - 1) / z_chunk;
        }
    }
    if (chunks > 65535 || P.n_mats > 65535) return VT_ERR_UNSUPPORTED;
    L.chunks = chunks;
    L.z_chunk = z_chunk;
    L.tma = tma_ok(P) && !(P.flags & VT_STAGE_CP_ASYNC);
    L.box_w = L.box_h = 0;
    L.cost = 0.0f;
    if (L.tma) {
        // box: the largest footprint of the batch (+5 texels of filter support / rounding), rows padded to 16 B
        float ext_y = 0.0f, ext_x = 0.0f;
        for (int k = 0; k < P.n_mats; k++) {
            ext_y = fmaxf(ext_y, (fabsf(P.mats[k].r[1][1]) + fabsf(P.mats[k].r[1][2])) * (float)(TS - 1));
            ext_x = fmaxf(ext_x, (fabsf(P.mats[k].r[2][1]) + fabsf(P.mats[k].r[2][2])) * (float)(TS - 1));
        }
        const int need_h = min(BMAX, (int)floorf(ext_y + 0.1f) + 6), need_w = min(BMAX, (int)floorf(ext_x + 0.1f) + 6);
        // + 3: the box start is rounded down to a multiple of 4 texels (16-byte aligned start address)
        // the box width is the shared-memory pitch of every matrix of the launch; each matrix then takes its best shape
        int best_w = (need_w + 3 + 3) / 4 * 4;
        float best_cost = 1e30f;
        unsigned char shape[VT_MAX_BATCH];
        unsigned mask = 0;
        for (int w = best_w; w <= PITCH_MAX; w += 4) mask |= 1u << (w - PITCH_LO);
        for (int k = 0; k < P.n_mats; k++) conflict_table(P.mats[k], mask);  // all widths of a matrix in one pass
        for (int w = (need_w + 3 + 3) / 4 * 4; w <= PITCH_MAX; w += 4) {
            float total = 0.0f;
            unsigned char sh[VT_MAX_BATCH];
            for (int k = 0; k < P.n_mats; k++) {
                float bc = 1e30f;
                for (int layout = 0; layout < N_LAYOUTS; layout++) {
                    const float c = pitch_cost<INTERP>(P.mats[k], w, layout);
                    if (c < bc) {
                        bc = c;
                        sh[k] = (unsigned char)layout;
                    }
                }
                total += bc;
            }
            if (total < best_cost) {
                best_cost = total;
                best_w = w;
                memcpy(shape, sh, sizeof sh);
            }
        }
        const char *force = getenv("VT_SLICE_LAYOUT");  // tuning knobs
        for (int k = 0; k < P.n_mats; k++)
            L.aux[k] = (unsigned char)((force ? atoi(force) % N_LAYOUTS : shape[k]) << 6);
        if (const char *e = getenv("VT_SLICE_W"))
            if (atoi(e) >= (need_w + 3 + 3) / 4 * 4 && atoi(e) <= PITCH_MAX && atoi(e) % 4 == 0) best_w = atoi(e);
        if (getenv("VT_SLICE_DEBUG"))
            fprintf(stderr, "slice interp %d: need %d x %d, box_w %d, shape[0] %d, cost %.3f\n", INTERP, need_w, need_h, best_w,
                    (int)(L.aux[0] >> 6), best_cost);
        L.box_w = best_w;
        L.box_h = need_h;
        L.cost = best_cost;
    } else {
        unsigned mask = 0;
        for (int pitch = PITCH_MIN; pitch <= PITCH_MAX; pitch++) mask |= 1u << (pitch - PITCH_LO);
        for (int k = 0; k < P.n_mats; k++) {
            conflict_table(P.mats[k], mask);
            int best = 33, best_layout = 0;
            float best_cost = 1e30f;
            for (int pitch = PITCH_MIN; pitch <= PITCH_MAX; pitch++)
                for (int layout = 0; layout < N_LAYOUTS; layout++) {
                    // (per-element cp.async staging: the pitch does not change the staged traffic; the fit is for TMA
                    // but the ordering by wavefronts is what matters here)
                    const float c = pitch_cost<INTERP>(P.mats[k], pitch, layout);
                    if (c < best_cost) {
                        best_cost = c;
                        best = pitch;
                        best_layout = layout;
                    }
                }
            L.aux[k] = (unsigned char)(best | (best_layout << 6));
        }
    }
    return VT_OK;
}

template <int INTERP, int RULE>
int launch2(VtResampleParams &P, cudaStream_t st)
{
    constexpr int PPS = Taps<INTERP>::PPS, NSTAGE = Taps<INTERP>::NSTAGE;
    const int tiles = ((P.o1 + TS - 1) / TS) * ((P.o2 + TS - 1) / TS);
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SlicePlan L;
    int rc = plan_slice<INTERP>(P, sms, L);
    if (rc) return rc;
    const int z_chunk = L.z_chunk;
    const bool tma = L.tma;
    memcpy(P.aux, L.aux, sizeof L.aux);
    dim3 grid(tiles, L.chunks, P.n_mats);
    VtSliceStaging G;
    memset(&G, 0, sizeof G);
    if (tma) {
        G.box_w = L.box_w;
        G.box_h = L.box_h;
        constexpr bool FIXED = INTERP != VT_LINEAR;  // see the kernel: cubic planes sit at the fixed stride STAGE
        G.plane_elems = FIXED ? STAGE : G.box_w * G.box_h;  // linear: a box of depth PPS lands as PPS densely packed planes
        G.stage_bytes = ((unsigned)(G.plane_elems * PPS * 4) + 127u) & ~127u;
        const unsigned long long gdim[3] = {(unsigned long long)P.s2, (unsigned long long)P.s1, (unsigned long long)P.s0};
        const unsigned long long gstr[2] = {(unsigned long long)P.src_row * 4, (unsigned long long)P.src_plane * 4};
        const unsigned box[3] = {(unsigned)G.box_w, (unsigned)G.box_h, FIXED ? 1u : (unsigned)PPS};
        rc = vt_encode_tmap_3d(&G.tmap, P.src, gdim, gstr, box);
        if (rc) return rc;
    } else {
        G.plane_elems = STAGE;
        G.stage_bytes = PPS * STAGE * 4;
    }
    const size_t smem = 128 + (size_t)NSTAGE * G.stage_bytes;
    // the attribute is per device and per function: one flag per device (a process may drive several GPUs)
    static std::atomic<bool> attr_set_dev[64];
    int attr_dev = 0;
    VT_CUDA(cudaGetDevice(&attr_dev));
    std::atomic<bool> &attr_set = attr_set_dev[attr_dev & 63];
    if (!attr_set.load(std::memory_order_acquire)) {
        const int mx = 128 + NSTAGE * PPS * STAGE * 4;
        VT_CUDA(cudaFuncSetAttribute(vt_slice_kernel<INTERP, RULE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        VT_CUDA(cudaFuncSetAttribute(vt_slice_kernel<INTERP, RULE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        VT_CUDA(cudaFuncSetAttribute(vt_slice_kernel<INTERP, RULE, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        VT_CUDA(cudaFuncSetAttribute(vt_slice_kernel<INTERP, RULE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        attr_set.store(true, std::memory_order_release);
    }
    {
        VtProf prof(VT_K_SLICE_LINEAR + INTERP, st);
        const bool zero = (P.flags & VT_OOB_ZERO) != 0;
        if (tma) {
            if (zero) vt_slice_kernel<INTERP, RULE, true, true><<<grid, NT, smem, st>>>(P, G, z_chunk);
            else vt_slice_kernel<INTERP, RULE, false, true><<<grid, NT, smem, st>>>(P, G, z_chunk);
        } else {
            if (zero) vt_slice_kernel<INTERP, RULE, true, false><<<grid, NT, smem, st>>>(P, G, z_chunk);
            else vt_slice_kernel<INTERP, RULE, false, false><<<grid, NT, smem, st>>>(P, G, z_chunk);
        }
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

template <int INTERP>
int launch1(VtResampleParams &P, cudaStream_t st)
{
    if (INTERP != VT_CUBIC_SIMPLE && (P.flags & VT_WEIGHTS_EXACT)) return launch2<INTERP, 2>(P, st);
    return launch2<INTERP, 0>(P, st);
}


// ---------------------------------------------------------------------------------------------------
// rotate-and-project (the reference's examples/projections.py:20-26 transforms the volume and
// then sums it over axis 0).  For a slice-family matrix the resampling is a fixed in-plane operator applied to every
// input plane plus three fixed axis-0 weights, so by linearity the sum over axis 0 of the transformed volume is that
// in-plane operator applied ONCE to sums of input planes:
//      proj = A(sum of planes V-1) + B(sum of planes V) + C(sum of planes V+1),
// V = the input planes at the centre of an in-bounds output plane, planes outside the source counting as zero (border
// mode), and (A, B, C) the interpolator's per-plane weight sets (linear: B only; cubic_simple: the same 16 taps times
// B(-1), B(0), B(1); cubic_tex: the three integer-weighted sets of Taps<VT_CUBIC_TEX>).  Cost: one read of the source
// (4 B/voxel, no output volume) + a 2-D resample, instead of o0 resampled planes written and read back.
// Float32 summation order differs from transform-then-sum (~1e-7 of the projection's range).
// ---------------------------------------------------------------------------------------------------
constexpr int PS_NT = 128;   // threads per CTA of the plane-sum kernel, along x
constexpr int PS_PAD = 2;    // zero border (texels) around the summed planes: cubic footprints reach -2 .. s+1
constexpr int PS_UNROLL = 8;

struct VtPlaneRanges {
    int b0, b1;  // V = [b0, b1): centre planes; A sums [b0-1, b1-1) and C sums [b0+1, b1+1), both clipped to the source
};

// partial[chunk][3][s1][s2] = sums of the source planes of z-chunk `chunk` that fall in the A / B / C ranges
__global__ void __launch_bounds__(PS_NT)
    vt_plane_sum_kernel(const float *__restrict__ src, int s0, int s1, int s2, long long src_row, long long src_plane,
                        VtPlaneRanges R, int z_chunk, float *__restrict__ partial)
{
    const int x = blockIdx.x * PS_NT + threadIdx.x, y = blockIdx.y, c = blockIdx.z;
    if (x >= s2) return;
    // planes any of the three ranges needs: [b0-1, b1+1) clipped
    const int lo = max(R.b0 - 1, 0), hi = min(R.b1 + 1, s0);
    const int q0 = lo + c * z_chunk, q1 = min(q0 + z_chunk, hi);
    const float *p = src + (size_t)q0 * src_plane + (size_t)y * src_row + x;
    float sa = 0.0f, sb = 0.0f, sc = 0.0f;
    float common = 0.0f;  // planes inside all three ranges (everything but two planes at either end): one add each
    int q = q0;
    for (; q + PS_UNROLL <= q1; q += PS_UNROLL) {
        float v[PS_UNROLL];
#pragma unroll
        for (int u = 0; u < PS_UNROLL; u++) v[u] = __ldg(p + (size_t)u * src_plane);
        if (q >= R.b0 + 1 && q + PS_UNROLL <= R.b1 - 1) {  // uniform
#pragma unroll
            for (int u = 0; u < PS_UNROLL; u++) common += v[u];
        } else {
#pragma unroll
            for (int u = 0; u < PS_UNROLL; u++) {
                const int qq = q + u;
                if (qq >= R.b0 - 1 && qq < R.b1 - 1) sa += v[u];
                if (qq >= R.b0 && qq < R.b1) sb += v[u];
                if (qq >= R.b0 + 1 && qq < R.b1 + 1) sc += v[u];
            }
        }
        p += (size_t)PS_UNROLL * src_plane;
    }
    sa += common;
    sb += common;
    sc += common;
    for (; q < q1; q++) {
        const float v = __ldg(p);
        if (q >= R.b0 - 1 && q < R.b1 - 1) sa += v;
        if (q >= R.b0 && q < R.b1) sb += v;
        if (q >= R.b0 + 1 && q < R.b1 + 1) sc += v;
        p += src_plane;
    }
    const size_t n = (size_t)s1 * s2, o = (size_t)y * s2 + x;
    float *out = partial + (size_t)c * 3 * n;
    out[o] = sa;
    out[n + o] = sb;
    out[2 * n + o] = sc;
}

// sums[k][PS_PAD + y][PS_PAD + x] = sum over chunks of partial[chunk][k][y][x] (fixed order: deterministic);
// linear keeps its single sum (B) in plane 0, the cubic modes A, B, C in planes 0, 1, 2
__global__ void __launch_bounds__(PS_NT)
    vt_plane_sum_finish_kernel(const float *__restrict__ partial, int chunks, int s1, int s2, int linear, int pitch, int pe,
                               float *__restrict__ sums)
{
    const int x = blockIdx.x * PS_NT + threadIdx.x, y = blockIdx.y, k = blockIdx.z;
    if (x >= s2) return;
    const size_t n = (size_t)s1 * s2;
    const float *p = partial + (size_t)(linear ? 1 : k) * n + (size_t)y * s2 + x;
    float s = 0.0f;
    for (int c = 0; c < chunks; c++) s += __ldg(p + (size_t)c * 3 * n);
    sums[(size_t)k * pe + (size_t)(PS_PAD + y) * pitch + (PS_PAD + x)] = s;
}

// one thread per projection pixel (a1, a2) and matrix: the interpolators' in-plane part on the summed planes
template <int INTERP, int RULE>
__global__ void __launch_bounds__(NT)
    vt_project2d_kernel(const __grid_constant__ VtResampleParams P, const float *__restrict__ sums, int pitch, int pe)
{
    using T = Taps<INTERP>;
    const int a2 = blockIdx.x * TS + (threadIdx.x & (TS - 1)), a1 = blockIdx.y * TS + (threadIdx.x / TS);
    const int mat = blockIdx.z;
    if (a1 >= P.o1 || a2 >= P.o2) return;
    const VtMat &M = P.mats[mat];
    const float p1 = inplane_coord(M.r[1], (float)a1, (float)a2);
    const float p2 = inplane_coord(M.r[2], (float)a1, (float)a2);
    float r = 0.0f;
    // transforms.py:276-278 for the two in-plane axes (axis 0 is in the plane ranges)
    if (!(p2 < 0 || p1 < 0 || p2 >= (float)P.s2 || p1 >= (float)P.s1)) {
        T taps;
        taps.template init<RULE>(p1, p2, -PS_PAD, -PS_PAD, pitch);
        if constexpr (INTERP == VT_LINEAR) {
            float o[T::PPS];
            taps.planes(sums, pe, o);
            r = o[0];
        } else if constexpr (INTERP == VT_CUBIC_SIMPLE) {
            float ab[T::PPS], bc[T::PPS];
            taps.planes(sums, pe, ab);  // in-plane sums of the A, B (and, with 4 planes per stage, C) plane sums
            float sc;
            if constexpr (T::PPS >= 3) {
                sc = ab[2];
            } else {
                taps.planes(sums + pe, pe, bc);
                sc = bc[1];
            }
            r = fmaf(taps.wz2, sc, fmaf(taps.wz1, ab[1], __fmul_rn(taps.wz0, ab[0])));
        } else {
            float qa[T::PPS], qb[T::PPS], qc[T::PPS], ra[T::PPS], rb[T::PPS], rc[T::PPS];
            taps.planes3(sums, pe, qa, qb, qc);
            float cc;
            if constexpr (T::PPS >= 3) {
                cc = qc[2];
            } else {
                taps.planes3(sums + pe, pe, ra, rb, rc);
                cc = rc[1];
            }
            r = (qa[0] + qb[1]) + cc;
        }
    }
    P.dst[(size_t)mat * P.dst_batch_stride + (size_t)a1 * P.o2 + a2] = r;
}

struct ProjectLayout {
    int pitch, pe, chunks, z_chunk;
    size_t partial_floats, sums_floats;
};
ProjectLayout project_layout(int s0, int s1, int s2)
{
    ProjectLayout L;
    L.pitch = (s2 + 2 * PS_PAD + 3) / 4 * 4;
    L.pe = L.pitch * (s1 + 2 * PS_PAD);
    // enough threads in flight to cover the HBM latency: ~600 K, each with PS_UNROLL loads outstanding
    const long long cols = (long long)s1 * s2;
    long long c = (600000 + cols - 1) / cols;
    const int span = s0 + 2;
    if (c > span / PS_UNROLL) c = span / PS_UNROLL;
    if (c > 64) c = 64;
    if (c < 1) c = 1;
    L.z_chunk = (span + (int)c - 1) / (int)c;
    L.chunks = (span + L.z_chunk - 1) / L.z_chunk;
    L.partial_floats = (size_t)L.chunks * 3 * (size_t)cols;
    L.sums_floats = (size_t)4 * L.pe;  // 4 planes: Taps<VT_LINEAR>::planes reads a stage of four
    return L;
}

template <int INTERP, int RULE>
int project2(VtResampleParams &P, float *ws, cudaStream_t st)
{
    const ProjectLayout L = project_layout(P.s0, P.s1, P.s2);
    float *partial = ws, *sums = ws + L.partial_floats;
    // matrices with the same integer shift along axis 0 share their plane sums
    for (int first = 0; first < P.n_mats;) {
        const float t0f = P.mats[first].r[0][3];
        int count = 1;
        while (first + count < P.n_mats && P.mats[first + count].r[0][3] == t0f) count++;
        const int t0 = (int)t0f;
        VtPlaneRanges R;
        R.b0 = max(0, P.z_begin + t0);
        R.b1 = min(P.s0, P.z_end + t0);
        if (R.b1 < R.b0) R.b1 = R.b0;
        VT_CUDA(cudaMemsetAsync(sums, 0, L.sums_floats * sizeof(float), st));
        const int lo = max(R.b0 - 1, 0), hi = min(R.b1 + 1, P.s0);
        if (R.b1 > R.b0 && hi > lo) {
            const int zc = (hi - lo + L.chunks - 1) / L.chunks;
            const int chunks = (hi - lo + zc - 1) / zc;
            {
                VtProf prof(VT_K_PLANE_SUM, st);
                dim3 grid((P.s2 + PS_NT - 1) / PS_NT, P.s1, chunks);
                vt_plane_sum_kernel<<<grid, PS_NT, 0, st>>>(P.src, P.s0, P.s1, P.s2, P.src_row, P.src_plane, R, zc, partial);
            }
            dim3 grid2((P.s2 + PS_NT - 1) / PS_NT, P.s1, INTERP == VT_LINEAR ? 1 : 3);
            vt_plane_sum_finish_kernel<<<grid2, PS_NT, 0, st>>>(partial, chunks, P.s1, P.s2, INTERP == VT_LINEAR ? 1 : 0,
                                                                 L.pitch, L.pe, sums);
            vt_count_launch(2);
        }
        VtResampleParams Q = P;
        Q.n_mats = count;
        for (int k = 0; k < count; k++) Q.mats[k] = P.mats[first + k];
        Q.dst = P.dst + (size_t)first * P.dst_batch_stride;
        {
            VtProf prof(VT_K_PROJECT_2D, st);
            dim3 grid((P.o2 + TS - 1) / TS, (P.o1 + TS - 1) / TS, count);
            vt_project2d_kernel<INTERP, RULE><<<grid, NT, 0, st>>>(Q, sums, L.pitch, L.pe);
        }
        vt_count_launch();
        VT_CUDA(cudaGetLastError());
        first += count;
    }
    return VT_OK;
}

template <int INTERP>
int project1(VtResampleParams &P, float *ws, cudaStream_t st)
{
    if (INTERP != VT_CUBIC_SIMPLE && (P.flags & VT_WEIGHTS_EXACT)) return project2<INTERP, 2>(P, ws, st);
    return project2<INTERP, 0>(P, ws, st);
}

}  // namespace

int vt_slice_supported(const VtResampleParams &P, int interp) { return slice_ok(P, interp) ? 1 : 0; }

// host-only: what vt_launch_slice would decide for these parameters on a GPU with `sms` multiprocessors
int vt_slice_plan_impl(const VtResampleParams &P, int interp, int sms, int *chunks, int *z_chunk, int *tma, int *box_w,
                       int *box_h, int *shapes, int *pitches)
{
    if (!slice_ok(P, interp)) return VT_ERR_UNSUPPORTED;
    SlicePlan L;
    int rc;
    switch (interp) {
        case VT_LINEAR: rc = plan_slice<VT_LINEAR>(P, sms, L); break;
        case VT_CUBIC_TEX: rc = plan_slice<VT_CUBIC_TEX>(P, sms, L); break;
        case VT_CUBIC_SIMPLE: rc = plan_slice<VT_CUBIC_SIMPLE>(P, sms, L); break;
        default: return VT_ERR_INVALID_ARG;
    }
    if (rc) return rc;
    *chunks = L.chunks;
    *z_chunk = L.z_chunk;
    *tma = L.tma ? 1 : 0;
    *box_w = L.box_w;
    *box_h = L.box_h;
    for (int k = 0; k < P.n_mats; k++) {
        if (shapes) shapes[k] = L.aux[k] >> 6;
        if (pitches) pitches[k] = L.tma ? L.box_w : (L.aux[k] & 63);
    }
    return VT_OK;
}

int vt_launch_slice(VtResampleParams &P, int interp, cudaStream_t st)
{
    if (P.z_end <= P.z_begin || P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    switch (interp) {
        case VT_LINEAR: return launch1<VT_LINEAR>(P, st);
        case VT_CUBIC_TEX: return launch1<VT_CUBIC_TEX>(P, st);
        case VT_CUBIC_SIMPLE: return launch1<VT_CUBIC_SIMPLE>(P, st);
    }
    return VT_ERR_INVALID_ARG;
}

size_t vt_slice_project_workspace_bytes(int s0, int s1, int s2)
{
    const ProjectLayout L = project_layout(s0, s1, s2);
    return (L.partial_floats + L.sums_floats) * sizeof(float);
}

// P.dst = n_mats projections of o1 x o2 (P.dst_batch_stride apart); sums output planes [z_begin, z_end)
int vt_launch_slice_project(VtResampleParams &P, int interp, float *d_workspace, cudaStream_t st)
{
    if (P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    switch (interp) {
        case VT_LINEAR: return project1<VT_LINEAR>(P, d_workspace, st);
        case VT_CUBIC_TEX: return project1<VT_CUBIC_TEX>(P, d_workspace, st);
        case VT_CUBIC_SIMPLE: return project1<VT_CUBIC_SIMPLE>(P, d_workspace, st);
    }
    return VT_ERR_INVALID_ARG;
}
