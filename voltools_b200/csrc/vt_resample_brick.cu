// vt_resample_brick.cu -- "brick" kernel family: general affine matrices through a TMA-staged shared-memory brick.
//
// A CTA owns an 8 x 8 x 16 tile of output voxels.  The image of the tile under the matrix is a parallelepiped; its
// axis-aligned bounding box plus the filter support (the "brick", typically ~18 x 22 x 26 texels = 40 KB for a
// rotation) is fetched with ONE 3-D cp.async.bulk.tensor box load: coordinates outside the source arrive as zeros,
// which is exactly the texture's border addressing mode, so the interpolators need no bounds checks at all.  Every
// thread then produces 4 voxels from shared memory: 8 (linear) or 64 (cubic) taps each, 4-byte bank-granular
// accesses instead of the 32-byte sectors the direct global gathers of vt_resample_gather.cu pay for.
// Several CTAs per SM overlap one CTA's box load with the others' arithmetic.
//
// Replaces the reference's `transform` kernel (voltools/transforms.py:253-282) + linearTex3D / cubicTex3D /
// cubicTex3DSimple (voltools/kernels/helper_interpolation.h:3-68) for matrices the slice family cannot take.
// Needs 16-byte aligned source rows (TMA); otherwise the gather family runs.
//
// Arithmetic: linear and cubic_tex are operation-for-operation the gather family's (bit-identical results);
// cubic_simple sums separably (x, then y, then z) instead of in the reference's single running sum: same weights,
// float32 rounding differs by ~1e-7 of the range.
#include <cuda.h>

#include <atomic>

#include "vt_common.cuh"

namespace {

constexpr int TY = 8, TX = 16;                // output tile in (a1, a2); along a0 it is 2 * VPT planes
constexpr int NT = 256;                       // threads: 16 (x) x 8 (y) x 2 (z halves), VPT voxels each along z
// Two tile depths.  8 planes (VPT = 4, the default): bricks <= 72 KB, 3 CTAs per SM.  16 planes (VPT = 8, knob
// VT_BRICK_VPT=8): bricks <= 110 KB, 2 CTAs per SM, ~25 % less L2 -> shared-memory traffic per voxel and half the per-tile
// set-up -- measured SLOWER at 512^3 (linear 173 vs 209, cubic_tex 56 vs 67 Gvox/s under the full affine): a CTA waits
// for its whole brick before it computes, so the third resident CTA hides more than the deeper tile saves.
constexpr int MAX_BRICK_BYTES_4 = 72 * 1024, MAX_BRICK_BYTES_8 = 110 * 1024;
__host__ __device__ constexpr int max_brick_bytes(int vpt) { return vpt == 4 ? MAX_BRICK_BYTES_4 : MAX_BRICK_BYTES_8; }

struct VtBrickStaging {
    CUtensorMap tmap;
    int bw, bh, bd;  // box = brick dimensions (bw multiple of 4; also the row pitch)
    int layout;      // which sub-box of the CTA's 16 x 8 x 2 thread block a warp covers (thread_pos; chosen against bank conflicts)
};

// thread -> position inside the 8 x 8 x 16 tile (tz selects a group of VPT consecutive planes).  The CTA's threads
// form a 16 (x) x 8 (y) x 2 (z groups) block; `layout` is the sub-box of it a warp covers: which one spreads a warp's
// 32 brick addresses over the most banks depends on the matrix (tune_brick).
constexpr int N_BRICK_LAYOUTS = 6;
__host__ __device__ __forceinline__ void thread_pos(int tid, int layout, int &tx, int &ty, int &tz)  // tz: 0 / 1
{
    // log2 of the warp's extent along x and y (the rest of its 32 lanes goes along z)
    const int lwx = layout == 0 ? 4 : (layout == 1 ? 3 : (layout == 2 ? 2 : (layout == 3 ? 3 : (layout == 4 ? 2 : 4))));
    const int lwy = layout == 0 ? 1 : (layout == 1 ? 2 : (layout == 2 ? 3 : (layout == 3 ? 1 : (layout == 4 ? 2 : 0))));
    const int lwz = 5 - lwx - lwy;  // 0 or 1
    const int lane = tid & 31, warp = tid >> 5;
    const int lx = lane & ((1 << lwx) - 1), ly = (lane >> lwx) & ((1 << lwy) - 1), lz = lane >> (lwx + lwy);
    const int nwx = 4 - lwx, nwy = 3 - lwy;  // log2 of the number of warps along x and y
    const int wx = warp & ((1 << nwx) - 1), wy = (warp >> nwx) & ((1 << nwy) - 1), wz = warp >> (nwx + nwy);
    tx = (wx << lwx) + lx;
    ty = (wy << lwy) + ly;
    tz = (wz << lwz) + lz;
}

struct Brick {
    const float *s;
    int py, pz;          // row and plane pitch (floats)
    int zlo, ylo, xlo;   // source index of brick element (0,0,0)
    __device__ __forceinline__ const float *at(int i0, int i1, int i2) const
    {
        return s + (i0 - zlo) * pz + (i1 - ylo) * py + (i2 - xlo);
    }
};

// one trilinear fetch with the float-pipeline weights (vt_tex_weights2): q = the (x near, y near, z near) texel.  The
// four FFMA2 carry the z-near and z-far partial sums side by side; they are added once at the end, so the float32
// summation order differs from the gather family's single chain (<= 1 ulp of the result; the unit itself rounds once).
__device__ __forceinline__ float tex_fetch_f(const float *q, int py, int pz, const VtTexW &w)
{
    const float *qz = q + pz;
    vt_f2 r = vt_mul2(w.nn, vt_pk(q[0], qz[0]));
    r = vt_fma2(w.fn, vt_pk(q[1], qz[1]), r);
    r = vt_fma2(w.nf, vt_pk(q[py], qz[py]), r);
    r = vt_fma2(w.ff, vt_pk(q[py + 1], qz[py + 1]), r);
    return __fmul_rn(__fadd_rn(vt_lo(r), vt_hi(r)), 1.0f / 256.0f);
}

// cubicTex3D (helper_interpolation.h:8-40) with the float-pipeline weights: the two fetch positions of an axis go
// through vt_tex_fix2 as a pair, and the x-side products of the weight rule, which depend on (x alpha, z side) only,
// are shared by the four fetches that use them: ~330 instructions per voxel instead of ~730.
__device__ __forceinline__ float cubic_tex_brick_f(const float *s, int py, int pz, int xlo, int ylo, int zlo, float x, float y,
                                                   float z)
{
    float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
    vt_ruijters(x, g0x, g1x, h0x, h1x);
    vt_ruijters(y, g0y, g1y, h0y, h1y);
    vt_ruijters(z, g0z, g1z, h0z, h1z);
    vt_f2 ax, ay, az;
    int bx[2], by[2], bz[2];
    vt_tex_fix2(vt_pk(h0x, h1x), ax, bx[0], bx[1]);
    vt_tex_fix2(vt_pk(h0y, h1y), ay, by[0], by[1]);
    vt_tex_fix2(vt_pk(h0z, h1z), az, bz[0], bz[1]);
    const float a[2] = {vt_lo(ax), vt_hi(ax)}, b[2] = {vt_lo(ay), vt_hi(ay)}, c[2] = {vt_lo(az), vt_hi(az)};
    vt_f2 XF[2][2], XN[2][2];  // [x fetch i][z fetch k], halves = z near / far side
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const vt_f2 S = vt_fma2(vt_bc(c[k]), vt_pk(-1.0f, 1.0f), vt_pk(256.0f, 0.0f));  // {256 - c, c}
#pragma unroll
        for (int i = 0; i < 2; i++) {
            XF[i][k] = vt_round8_2(vt_fma2(vt_bc(a[i]), S, vt_bc(128.0f)));
            XN[i][k] = vt_fma2(XF[i][k], vt_bc(-1.0f), S);
        }
    }
    const int ox[2] = {bx[0] - xlo, bx[1] - xlo};
    float tz[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        float tyv[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float *row = s + (bz[k] - zlo) * pz + (by[j] - ylo) * py;
            const vt_f2 bj = vt_bc(b[j]), nbj = vt_bc(256.0f - b[j]);
            float tx[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                VtTexW w;
                w.ff = vt_round8_2(vt_fma2(bj, XF[i][k], vt_bc(128.0f)));
                w.nn = vt_round8_2(vt_fma2(nbj, XN[i][k], vt_bc(128.0f)));
                w.fn = vt_fma2(w.ff, vt_bc(-1.0f), XF[i][k]);
                w.nf = vt_fma2(w.nn, vt_bc(-1.0f), XN[i][k]);
                tx[i] = tex_fetch_f(row + ox[i], py, pz, w);
            }
            tyv[j] = __fmaf_rn(g0x, tx[0], __fmul_rn(g1x, tx[1]));
        }
        tz[k] = __fmaf_rn(g0y, tyv[0], __fmul_rn(g1y, tyv[1]));
    }
    return __fmaf_rn(g0z, tz[0], __fmul_rn(g1z, tz[1]));
}

// tex3D<float> on the brick (x along axis 2, y along axis 1, z along axis 0); see vt_common.cuh for the weight rule
template <int RULE>
__device__ __forceinline__ float tex3d_brick(const Brick &b, float x, float y, float z)
{
    int i2, i1, i0;
    if (RULE == 0) {
        vt_f2 axy, az;
        int unused;
        vt_tex_fix2(vt_pk(x, y), axy, i2, i1);
        vt_tex_fix2(vt_pk(z, z), az, i0, unused);
        return tex_fetch_f(b.at(i0, i1, i2), b.py, b.pz, vt_tex_weights2(vt_lo(axy), vt_hi(axy), vt_lo(az)));
    } else {
        float ax, ay, az;
        vt_tex_fix<2>(x, i2, ax);
        vt_tex_fix<2>(y, i1, ay);
        vt_tex_fix<2>(z, i0, az);
        const float *q = b.at(i0, i1, i2);
        const float bx = 1.0f - ax, by = 1.0f - ay, bz = 1.0f - az;
        const float r00 = bx * q[0] + ax * q[1], r01 = bx * q[b.py] + ax * q[b.py + 1];
        const float r10 = bx * q[b.pz] + ax * q[b.pz + 1], r11 = bx * q[b.pz + b.py] + ax * q[b.pz + b.py + 1];
        const float s0 = by * r00 + ay * r01, s1 = by * r10 + ay * r11;
        return bz * s0 + az * s1;
    }
}

// cubicTex3D, helper_interpolation.h:8-40 (same fetch / combine order as the gather family)
template <int RULE>
__device__ __forceinline__ float cubic_tex_brick(const Brick &b, float x, float y, float z)
{
    float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
    vt_ruijters(x, g0x, g1x, h0x, h1x);
    vt_ruijters(y, g0y, g1y, h0y, h1y);
    vt_ruijters(z, g0z, g1z, h0z, h1z);
    float t000 = tex3d_brick<RULE>(b, h0x, h0y, h0z), t100 = tex3d_brick<RULE>(b, h1x, h0y, h0z);
    t000 = __fmaf_rn(g0x, t000, __fmul_rn(g1x, t100));
    float t010 = tex3d_brick<RULE>(b, h0x, h1y, h0z), t110 = tex3d_brick<RULE>(b, h1x, h1y, h0z);
    t010 = __fmaf_rn(g0x, t010, __fmul_rn(g1x, t110));
    t000 = __fmaf_rn(g0y, t000, __fmul_rn(g1y, t010));
    float t001 = tex3d_brick<RULE>(b, h0x, h0y, h1z), t101 = tex3d_brick<RULE>(b, h1x, h0y, h1z);
    t001 = __fmaf_rn(g0x, t001, __fmul_rn(g1x, t101));
    float t011 = tex3d_brick<RULE>(b, h0x, h1y, h1z), t111 = tex3d_brick<RULE>(b, h1x, h1y, h1z);
    t011 = __fmaf_rn(g0x, t011, __fmul_rn(g1x, t111));
    t001 = __fmaf_rn(g0y, t001, __fmul_rn(g1y, t011));
    return __fmaf_rn(g0z, t000, __fmul_rn(g1z, t001));
}

// cubicTex3DSimple, helper_interpolation.h:42-68: the reference's weights, summed separably
__device__ __forceinline__ float cubic_simple_brick(const Brick &b, float x, float y, float z)
{
    const float cgx = __fadd_rn(x, -0.5f), cgy = __fadd_rn(y, -0.5f), cgz = __fadd_rn(z, -0.5f);
    const float fx0 = floorf(cgx), fy0 = floorf(cgy), fz0 = floorf(cgz);
    const float fx = __fsub_rn(cgx, fx0), fy = __fsub_rn(cgy, fy0), fz = __fsub_rn(cgz, fz0);
    float wx[4], wy[4], wz[4];
    vt_bspline4(fx, wx);
    vt_bspline4(fy, wy);
    vt_bspline4(fz, wz);
    const float *q = b.at((int)fz0 - 1, (int)fy0 - 1, (int)fx0 - 1);
    float r = 0.0f;
#pragma unroll
    for (int kz = 0; kz < 4; kz++) {
        float ry = 0.0f;
#pragma unroll
        for (int ky = 0; ky < 4; ky++) {
            const float *row = q + kz * b.pz + ky * b.py;
            float rx = wx[0] * row[0];
            rx = fmaf(wx[1], row[1], rx);
            rx = fmaf(wx[2], row[2], rx);
            rx = fmaf(wx[3], row[3], rx);
            ry = fmaf(wy[ky], rx, ry);
        }
        r = fmaf(wz[kz], ry, r);
    }
    return r;
}

// MODE: 0 = out-of-bounds voxels are skipped, 1 = written as zero, 2 = rotate-and-project (voxels are summed along axis 0)
template <int INTERP, int RULE, int MODE, int VPT>
__global__ void __launch_bounds__(NT, VPT == 4 ? 3 : 2)
    vt_brick_kernel(const __grid_constant__ VtResampleParams P, const __grid_constant__ VtBrickStaging G)
{
    constexpr int TZ = 2 * VPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];  // [128 B: mbarrier][brick]
    const unsigned bar = vt_smem_u32(smem_raw);
    const float *brick = (const float *)(smem_raw + 128);
    const int tid = threadIdx.x;
    const int nz = P.z_end - P.z_begin;
    const int nzt = (nz + TZ - 1) / TZ;
    const int mat = blockIdx.z / nzt;
    const int zt = blockIdx.z - mat * nzt;
    const VtMat &M = P.mats[mat];
    const int a0_0 = P.z_begin + zt * TZ, a1_0 = blockIdx.y * TY, a2_0 = blockIdx.x * TX;
    const int a0_1 = min(a0_0 + TZ, P.z_end) - 1, a1_1 = min(a1_0 + TY, P.o1) - 1, a2_1 = min(a2_0 + TX, P.o2) - 1;

    // Brick origin: per input axis the extremes are at tile corners (the float recipe is monotone in each index).
    // Eight lanes of warp 0 take one corner each and min / max are reduced with shuffles: the per-tile set-up costs a
    // few dozen instructions on one warp instead of ~300 on every thread (a thread only produces VPT = 4 voxels, so
    // that redundant prologue was more than half of the trilinear kernel's instruction count).
    constexpr int LO = INTERP == VT_LINEAR ? 0 : -1;
    int *setup = (int *)(smem_raw + 16);  // [lo0, lo1, lo2, interior], after the mbarrier
    if (tid < 32) {
        const int c = tid & 7;
        const float fa0 = (float)((c & 4) ? a0_1 : a0_0), fa1 = (float)((c & 2) ? a1_1 : a1_0);
        const float fa2 = (float)((c & 1) ? a2_1 : a2_0);
        int lo_[3];
        bool inside = true;  // every voxel of the tile samples inside the source: no per-voxel bounds tests
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const float p = vt_row_finish(M.r[r], vt_row_base(M.r[r], fa0, fa1), fa2);
            float mn = p, mx = p;
#pragma unroll
            for (int off = 4; off >= 1; off >>= 1) {
                mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            }
            // sample points further out than 2 texels are out of bounds anyway: keep the conversion safe
            const float dim = (float)(r == 0 ? P.s0 : (r == 1 ? P.s1 : P.s2));
            inside = inside && mn >= 0.0f && mx < dim;  // transforms.py:276-278 holds for the extremes, hence for all
            mn = fminf(fmaxf(mn, -2.0f), dim + 2.0f);
            lo_[r] = (int)floorf(mn - 0.5f) + LO;
        }
        lo_[2] &= ~3;  // the TMA box must start on a 16-byte boundary of the row
        if (tid == 0) {
            vt_tma_prefetch_desc(&G.tmap);
            vt_mbar_init(bar, 1);
            vt_mbar_fence_init();
            vt_mbar_expect_tx(bar, (unsigned)(G.bw * G.bh * G.bd) * 4u);
            vt_tma_load_3d(vt_smem_u32(brick), &G.tmap, bar, lo_[2], lo_[1], lo_[0]);
            setup[0] = lo_[0];
            setup[1] = lo_[1];
            setup[2] = lo_[2];
            setup[3] = inside ? 1 : 0;
        }
    }
    __syncthreads();  // the barrier is initialised and the set-up published before anyone goes on
    const int4 su = *reinterpret_cast<const int4 *>(setup);
    const bool interior = su.w != 0;
    const Brick b{brick, G.bw, G.bw * G.bh, su.x, su.y, su.z};
    // this thread's voxels: (a0 = a0_0 + tz*VPT + v, a1, a2)
    int tx, ty, tz;
    thread_pos(tid, G.layout, tx, ty, tz);
    const int a1 = a1_0 + ty, a2 = a2_0 + tx;
    const bool live = a1 < P.o1 && a2 < P.o2;
    const float f0 = (float)P.s0, f1 = (float)P.s1, f2 = (float)P.s2;
    const float fa1 = (float)a1, fa2 = (float)a2;
    float *__restrict__ dst = P.dst + (size_t)mat * P.dst_batch_stride + ((size_t)a1 * P.o2 + a2);
    const size_t oplane = (size_t)P.o1 * P.o2;
    vt_mbar_wait(bar, 0);
    if (!live) return;
    constexpr bool project = MODE == 2;  // sum along axis 0 instead of storing
    constexpr bool OOB_ZERO = MODE == 1;
    float acc = 0.0f;
    const int av0 = a0_0 + tz * VPT;
    if (INTERP == VT_LINEAR && RULE == 0 && interior && av0 + VPT - 1 <= a0_1) {
        // Fast path of the trilinear kernel: whole tile in bounds, all VPT voxels present.  Two voxels per instruction
        // through the coordinate recipe and vt_tex_fix2; weights on the float pipes (vt_tex_weights2).
        const float b0 = __fmul_rn(fa1, M.r[0][1]), b1 = __fmul_rn(fa1, M.r[1][1]), b2 = __fmul_rn(fa1, M.r[2][1]);
#pragma unroll
        for (int v = 0; v < VPT; v += 2) {
            const vt_f2 a0p = vt_pk((float)(av0 + v), (float)(av0 + v + 1));
            vt_f2 al[3];
            int bl[3], bh[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const float br = r == 0 ? b0 : (r == 1 ? b1 : b2);
                // vt_row_base / vt_row_finish, both voxels at once
                vt_f2 t = vt_fma2(a0p, vt_bc(M.r[r][0]), vt_bc(br));
                t = vt_fma2(vt_bc(fa2), vt_bc(M.r[r][2]), t);
                t = vt_add2(vt_add2(vt_bc(M.r[r][3]), t), vt_bc(0.5f));
                vt_tex_fix2(t, al[r], bl[r], bh[r]);
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float a = h ? vt_hi(al[2]) : vt_lo(al[2]), bb = h ? vt_hi(al[1]) : vt_lo(al[1]);
                const float c = h ? vt_hi(al[0]) : vt_lo(al[0]);
                const VtTexW w = vt_tex_weights2(a, bb, c);
                const float r = tex_fetch_f(b.at(h ? bh[0] : bl[0], h ? bh[1] : bl[1], h ? bh[2] : bl[2]), b.py, b.pz, w);
                if (project) acc += r;
                else dst[(size_t)(av0 + v + h) * oplane] = r;
            }
        }
    } else {
#pragma unroll
        for (int v = 0; v < VPT; v++) {
            const int a0 = av0 + v;
            if (a0 > a0_1) break;
            const float fa0 = (float)a0;
            const float p0 = vt_row_finish(M.r[0], vt_row_base(M.r[0], fa0, fa1), fa2);
            const float p1 = vt_row_finish(M.r[1], vt_row_base(M.r[1], fa0, fa1), fa2);
            const float p2 = vt_row_finish(M.r[2], vt_row_base(M.r[2], fa0, fa1), fa2);
            // transforms.py:276-278
            if (!interior && (p2 < 0 || p1 < 0 || p0 < 0 || p2 >= f2 || p1 >= f1 || p0 >= f0)) {
                if (OOB_ZERO && !project) dst[(size_t)a0 * oplane] = 0.0f;
                continue;
            }
            float r;
            if (INTERP == VT_LINEAR) r = tex3d_brick<RULE>(b, p2, p1, p0);
            else if (INTERP == VT_CUBIC_TEX) {
                if (RULE == 0) r = cubic_tex_brick_f(b.s, b.py, b.pz, b.xlo, b.ylo, b.zlo, p2, p1, p0);
                else r = cubic_tex_brick<RULE>(b, p2, p1, p0);
            } else r = cubic_simple_brick(b, p2, p1, p0);
            if (project) acc += r;
            else dst[(size_t)a0 * oplane] = r;
        }
    }
    if (project && acc != 0.0f) atomicAdd(dst, acc);
}

// brick dimensions needed by the batch (max over matrices), or false if some matrix needs more than fits
bool brick_dims(const VtResampleParams &P, int interp, int vpt, int &bw, int &bh, int &bd)
{
    const int TZ = 2 * vpt;
    if ((P.src_row % 4) != 0 || (P.src_plane % 4) != 0 || ((uintptr_t)P.src % 16) != 0) return false;
    if (P.n_mats < 1) return false;
    const int t[3] = {TZ - 1, TY - 1, TX - 1};
    const int support = interp == VT_LINEAR ? 5 : 6;  // taps + floor slack (see the coverage argument in DESIGN.md)
    int need[3] = {0, 0, 0};
    for (int k = 0; k < P.n_mats; k++) {
        for (int r = 0; r < 3; r++) {
            const float *m = P.mats[k].r[r];
            const float ext = fabsf(m[0]) * t[0] + fabsf(m[1]) * t[1] + fabsf(m[2]) * t[2];
            if (!(ext < 250.0f) || !(fabsf(m[3]) < 1e6f)) return false;
            const int n = (int)floorf(ext + 0.1f) + support;
            if (n > need[r]) need[r] = n;
        }
    }
    bd = need[0];
    bh = need[1];
    bw = (need[2] + 3 + 3) / 4 * 4;  // + 3: the box start is rounded down to a multiple of 4 texels
    if (bd > 256 || bh > 256 || bw > 256) return false;
    return (size_t)bw * bh * bd * 4 <= (size_t)max_brick_bytes(vpt);
}

// host copy of the coordinate recipe (same operations; used only to predict bank conflicts)
float host_coord(const float *row, float a0, float a1, float a2)
{
    float t = a1 * row[1];
    t = fmaf(a0, row[0], t);
    t = fmaf(a2, row[2], t);
    return (row[3] + t) + 0.5f;
}

// Shared-memory bank conflicts are what bounds the cubic brick kernels (64 four-byte loads per voxel): a warp's
// lanes hit bank (z*bw*bh + y*bw + x) mod 32.  The brick may be padded (bw by multiples of 4 texels, bh by a few
// rows) and the warp may tile the output 2 x 16 or 4 x 8; pick the combination with the fewest predicted conflicts
// for the first matrix of the batch over a couple of sample tiles.
void tune_brick(const VtResampleParams &P, VtBrickStaging &G, int VPT)
{
    const int TZ = 2 * VPT;
    const int MAX_BRICK_BYTES = max_brick_bytes(VPT);
    const VtMat &M = P.mats[0];
    constexpr int NS = 2;
    // The search below costs ~100 us of host time.  Small launches cannot win that back (tools/latency_probe.py: a 16^3
    // filt_bspline transform under a general matrix took 198 us per call, 88 us under a slice-family one), and a
    // resident volume keeps getting the same matrices: skip it below 8 M output voxels, memoise it above.
    if ((long long)(P.z_end - P.z_begin) * P.o1 * P.o2 * P.n_mats < (8LL << 20)) return;
    struct Memo {
        VtMat m;
        int key[7];
        int bw, bh, layout;
        bool valid;
    };
    static thread_local Memo memo[8];
    static thread_local unsigned memo_next = 0;
    const int key[7] = {P.o1, P.o2, P.z_begin + 4096 * VPT, P.z_end, G.bw, G.bh, G.bd};
    for (const Memo &e : memo)
        if (e.valid && memcmp(&e.m, &M, sizeof M) == 0 && memcmp(e.key, key, sizeof key) == 0) {
            G.bw = e.bw;
            G.bh = e.bh;
            G.layout = e.layout;
            return;
        }
    Memo &slot = memo[memo_next++ & 7];
    slot.m = M;
    memcpy(slot.key, key, sizeof key);
    slot.valid = false;
    static thread_local int iz[N_BRICK_LAYOUTS][NS][NT], iy[N_BRICK_LAYOUTS][NS][NT], ix[N_BRICK_LAYOUTS][NS][NT];
    for (int layout = 0; layout < N_BRICK_LAYOUTS; layout++)
        for (int smp = 0; smp < NS; smp++) {
            const int a0_0 = P.z_begin + ((P.z_end - P.z_begin) / 3 * (smp + 1)) / TZ * TZ;
            const int a1_0 = (P.o1 / 3 * (smp + 1)) / TY * TY, a2_0 = (P.o2 / 3 * (2 - smp)) / TX * TX;
            for (int tid = 0; tid < NT; tid++) {
                int tx, ty, tz;
                thread_pos(tid, layout, tx, ty, tz);
                const float a0 = (float)(a0_0 + tz * VPT), a1 = (float)(a1_0 + ty), a2 = (float)(a2_0 + tx);
                iz[layout][smp][tid] = (int)floorf(host_coord(M.r[0], a0, a1, a2) - 0.5f);
                iy[layout][smp][tid] = (int)floorf(host_coord(M.r[1], a0, a1, a2) - 0.5f);
                ix[layout][smp][tid] = (int)floorf(host_coord(M.r[2], a0, a1, a2) - 0.5f);
            }
        }
    const int bw0 = G.bw, bh0 = G.bh;
    long best = -1;
    for (int bw = bw0; bw <= bw0 + 8; bw += 4)
        for (int bh = bh0; bh <= bh0 + 7; bh++) {
            if ((size_t)bw * bh * G.bd * 4 > (size_t)MAX_BRICK_BYTES) continue;
            for (int layout = 0; layout < N_BRICK_LAYOUTS; layout++) {
                long cost = 0;
                for (int smp = 0; smp < NS; smp++)
                    for (int w = 0; w < NT / 32; w++) {
                        unsigned char cnt[32] = {0};
                        int first[32];
                        int worst = 0;
                        for (int l = 0; l < 32; l++) {
                            const int t = w * 32 + l;
                            const int addr = (iz[layout][smp][t] * bh + iy[layout][smp][t]) * bw + ix[layout][smp][t];
                            const int bank = addr & 31;
                            if (cnt[bank] && first[bank] == addr) continue;
                            if (!cnt[bank]) first[bank] = addr;
                            if (++cnt[bank] > worst) worst = cnt[bank];
                        }
                        cost += worst;
                    }
                // prefer smaller bricks on ties (less shared memory, less L2 traffic); a warp whose lanes cover only 4
                // consecutive x writes 16-byte pieces of a row: a small handicap
                cost = cost * 4096 + (long)bw * bh / 8 + ((layout == 2 || layout == 4) ? cost * 4096 / 16 : 0);
                if (best < 0 || cost < best) {
                    best = cost;
                    G.bw = bw;
                    G.bh = bh;
                    G.layout = layout;
                }
            }
        }
    slot.bw = G.bw;
    slot.bh = G.bh;
    slot.layout = G.layout;
    slot.valid = true;
}

template <int INTERP, int RULE, int VPT>
int launch3(const VtResampleParams &P, cudaStream_t st)
{
    constexpr int TZ = 2 * VPT;
    VtBrickStaging G;
    memset(&G, 0, sizeof G);
    if (!brick_dims(P, INTERP, VPT, G.bw, G.bh, G.bd)) return VT_ERR_UNSUPPORTED;
    tune_brick(P, G, VPT);  // (linear too: a rotation that sends a warp's x run down the brick's z axis is 8-way conflicted untuned)
    const unsigned long long gdim[3] = {(unsigned long long)P.s2, (unsigned long long)P.s1, (unsigned long long)P.s0};
    const unsigned long long gstr[2] = {(unsigned long long)P.src_row * 4, (unsigned long long)P.src_plane * 4};
    const unsigned box[3] = {(unsigned)G.bw, (unsigned)G.bh, (unsigned)G.bd};
    const int rc = vt_encode_tmap_3d(&G.tmap, P.src, gdim, gstr, box);
    if (rc) return rc;
    const int nz = P.z_end - P.z_begin;
    const int nzt = (nz + TZ - 1) / TZ;
    dim3 grid((P.o2 + TX - 1) / TX, (P.o1 + TY - 1) / TY, nzt * P.n_mats);
    if (grid.y > 65535u || grid.z > 65535u) return VT_ERR_UNSUPPORTED;
    const size_t smem = 128 + (size_t)G.bw * G.bh * G.bd * 4;
    // the attribute is per device and per function: one flag per device (a process may drive several GPUs)
    static std::atomic<bool> attr_set_dev[64];
    int attr_dev = 0;
    VT_CUDA(cudaGetDevice(&attr_dev));
    std::atomic<bool> &attr_set = attr_set_dev[attr_dev & 63];
    if (!attr_set.load(std::memory_order_acquire)) {
        const int mx = 128 + max_brick_bytes(VPT);
        VT_CUDA(cudaFuncSetAttribute(vt_brick_kernel<INTERP, RULE, 0, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        VT_CUDA(cudaFuncSetAttribute(vt_brick_kernel<INTERP, RULE, 1, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        VT_CUDA(cudaFuncSetAttribute(vt_brick_kernel<INTERP, RULE, 2, VPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
        attr_set.store(true, std::memory_order_release);
    }
    {
        VtProf prof(VT_K_BRICK_LINEAR + INTERP, st);
        if (P.flags & VT_INTERNAL_PROJECT) vt_brick_kernel<INTERP, RULE, 2, VPT><<<grid, NT, smem, st>>>(P, G);
        else if (P.flags & VT_OOB_ZERO) vt_brick_kernel<INTERP, RULE, 1, VPT><<<grid, NT, smem, st>>>(P, G);
        else vt_brick_kernel<INTERP, RULE, 0, VPT><<<grid, NT, smem, st>>>(P, G);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

// tile depth (see the constants at the top: the deep tile is a measurement knob)
template <int INTERP, int RULE>
int launch2(const VtResampleParams &P, cudaStream_t st)
{
    static const int force = getenv("VT_BRICK_VPT") ? atoi(getenv("VT_BRICK_VPT")) : 0;  // tuning knob: 4 / 8
    int bw, bh, bd;
    const bool deep_fits = brick_dims(P, INTERP, 8, bw, bh, bd) && (P.z_end - P.z_begin) >= 16;
    const bool deep = force == 8 && deep_fits;
    if (deep) return launch3<INTERP, RULE, 8>(P, st);
    return launch3<INTERP, RULE, 4>(P, st);
}

template <int INTERP>
int launch1(const VtResampleParams &P, cudaStream_t st)
{
    if (INTERP != VT_CUBIC_SIMPLE && (P.flags & VT_WEIGHTS_EXACT)) return launch2<INTERP, 2>(P, st);
    return launch2<INTERP, 0>(P, st);
}

}  // namespace

int vt_brick_supported(const VtResampleParams &P, int interp)
{
    int bw, bh, bd;
    return brick_dims(P, interp, 4, bw, bh, bd) ? 1 : 0;
}

int vt_launch_brick(const VtResampleParams &P, int interp, cudaStream_t st)
{
    if (P.z_end <= P.z_begin || P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    switch (interp) {
        case VT_LINEAR: return launch1<VT_LINEAR>(P, st);
        case VT_CUBIC_TEX: return launch1<VT_CUBIC_TEX>(P, st);
        case VT_CUBIC_SIMPLE: return launch1<VT_CUBIC_SIMPLE>(P, st);
    }
    return VT_ERR_INVALID_ARG;
}
