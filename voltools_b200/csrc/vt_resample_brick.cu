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

constexpr int TZ = 8, TY = 8, TX = 16;        // output tile
constexpr int NT = 256;                       // threads: 16 (x) x 8 (y) x 2 (z halves), 4 voxels each along z
constexpr int VPT = TZ * TY * TX / NT;        // 4
constexpr int MAX_BRICK_BYTES = 72 * 1024;    // 3 CTAs per SM

struct VtBrickStaging {
    CUtensorMap tmap;
    int bw, bh, bd;  // box = brick dimensions (bw multiple of 4; also the row pitch)
    int layout;      // how a warp's 32 lanes tile (a1, a2): 0 = 2 x 16, 1 = 4 x 8 (chosen against bank conflicts)
};

// thread -> position inside the 8 x 8 x 16 tile (tz selects a group of VPT consecutive planes)
__host__ __device__ __forceinline__ void thread_pos(int tid, int layout, int &tx, int &ty, int &tz)
{
    if (layout == 0) {
        tx = tid & 15;
        ty = (tid >> 4) & 7;
        tz = tid >> 7;
    } else {
        const int lane = tid & 31, warp = tid >> 5;
        tx = (warp & 1) * 8 + (lane & 7);
        ty = ((warp >> 1) & 1) * 4 + (lane >> 3);
        tz = warp >> 2;
    }
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

// tex3D<float> on the brick (x along axis 2, y along axis 1, z along axis 0); see vt_common.cuh for the weight rule
template <int RULE>
__device__ __forceinline__ float tex3d_brick(const Brick &b, float x, float y, float z)
{
    int i2, i1, i0;
    if (RULE == 0) {
        int a, bb, c;
        vt_tex_fix_hw(x, i2, a);
        vt_tex_fix_hw(y, i1, bb);
        vt_tex_fix_hw(z, i0, c);
        const float *q = b.at(i0, i1, i2);
        int wn[4], wf[4];
        vt_tex_hw_side(a, bb, 256 - c, wn);
        vt_tex_hw_side(a, bb, c, wf);
        float r = __fmul_rn(vt_u2f(wn[0]), q[0]);
        r = __fmaf_rn(vt_u2f(wn[1]), q[1], r);
        r = __fmaf_rn(vt_u2f(wn[2]), q[b.py], r);
        r = __fmaf_rn(vt_u2f(wn[3]), q[b.py + 1], r);
        r = __fmaf_rn(vt_u2f(wf[0]), q[b.pz], r);
        r = __fmaf_rn(vt_u2f(wf[1]), q[b.pz + 1], r);
        r = __fmaf_rn(vt_u2f(wf[2]), q[b.pz + b.py], r);
        r = __fmaf_rn(vt_u2f(wf[3]), q[b.pz + b.py + 1], r);
        return __fmul_rn(r, 1.0f / 256.0f);
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
template <int INTERP, int RULE, int MODE>
__global__ void __launch_bounds__(NT, 3)
    vt_brick_kernel(const __grid_constant__ VtResampleParams P, const __grid_constant__ VtBrickStaging G)
{
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

    // brick origin: per input axis the extremes are at tile corners (the float recipe is monotone in each index)
    constexpr int LO = INTERP == VT_LINEAR ? 0 : -1;
    int lo[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        float mn = 3.0e38f;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const float fa0 = (float)((c & 4) ? a0_1 : a0_0), fa1 = (float)((c & 2) ? a1_1 : a1_0);
            const float fa2 = (float)((c & 1) ? a2_1 : a2_0);
            mn = fminf(mn, vt_row_finish(M.r[r], vt_row_base(M.r[r], fa0, fa1), fa2));
        }
        // sample points further out than 2 texels are out of bounds anyway: keep the conversion safe
        const float dim = (float)(r == 0 ? P.s0 : (r == 1 ? P.s1 : P.s2));
        mn = fminf(fmaxf(mn, -2.0f), dim + 2.0f);
        lo[r] = (int)floorf(mn - 0.5f) + LO;
    }
    lo[2] &= ~3;  // the TMA box must start on a 16-byte boundary of the row
    if (tid == 0) {
        vt_tma_prefetch_desc(&G.tmap);
        vt_mbar_init(bar, 1);
        vt_mbar_fence_init();
        vt_mbar_expect_tx(bar, (unsigned)(G.bw * G.bh * G.bd) * 4u);
        vt_tma_load_3d(vt_smem_u32(brick), &G.tmap, bar, lo[2], lo[1], lo[0]);
    }
    const Brick b{brick, G.bw, G.bw * G.bh, lo[0], lo[1], lo[2]};
    // this thread's voxels: (a0 = a0_0 + tz*VPT + v, a1, a2)
    int tx, ty, tz;
    thread_pos(tid, G.layout, tx, ty, tz);
    const int a1 = a1_0 + ty, a2 = a2_0 + tx;
    const bool live = a1 < P.o1 && a2 < P.o2;
    const float f0 = (float)P.s0, f1 = (float)P.s1, f2 = (float)P.s2;
    const float fa1 = (float)a1, fa2 = (float)a2;
    float *__restrict__ dst = P.dst + (size_t)mat * P.dst_batch_stride + ((size_t)a1 * P.o2 + a2);
    const size_t oplane = (size_t)P.o1 * P.o2;
    __syncthreads();  // the barrier is initialised before anyone polls it
    vt_mbar_wait(bar, 0);
    if (!live) return;
    constexpr bool project = MODE == 2;  // sum along axis 0 instead of storing
    constexpr bool OOB_ZERO = MODE == 1;
    float acc = 0.0f;
#pragma unroll
    for (int v = 0; v < VPT; v++) {
        const int a0 = a0_0 + tz * VPT + v;
        if (a0 > a0_1) break;
        const float fa0 = (float)a0;
        const float p0 = vt_row_finish(M.r[0], vt_row_base(M.r[0], fa0, fa1), fa2);
        const float p1 = vt_row_finish(M.r[1], vt_row_base(M.r[1], fa0, fa1), fa2);
        const float p2 = vt_row_finish(M.r[2], vt_row_base(M.r[2], fa0, fa1), fa2);
        // transforms.py:276-278
        if (p2 < 0 || p1 < 0 || p0 < 0 || p2 >= f2 || p1 >= f1 || p0 >= f0) {
            if (OOB_ZERO && !project) dst[(size_t)a0 * oplane] = 0.0f;
            continue;
        }
        float r;
        if (INTERP == VT_LINEAR) r = tex3d_brick<RULE>(b, p2, p1, p0);
        else if (INTERP == VT_CUBIC_TEX) r = cubic_tex_brick<RULE>(b, p2, p1, p0);
        else r = cubic_simple_brick(b, p2, p1, p0);
        if (project) acc += r;
        else dst[(size_t)a0 * oplane] = r;
    }
    if (project && acc != 0.0f) atomicAdd(dst, acc);
}

// brick dimensions needed by the batch (max over matrices), or false if some matrix needs more than fits
bool brick_dims(const VtResampleParams &P, int interp, int &bw, int &bh, int &bd)
{
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
    return (size_t)bw * bh * bd * 4 <= (size_t)MAX_BRICK_BYTES;
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
void tune_brick(const VtResampleParams &P, VtBrickStaging &G)
{
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
    const int key[7] = {P.o1, P.o2, P.z_begin, P.z_end, G.bw, G.bh, G.bd};
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
    static thread_local int iz[2][NS][NT], iy[2][NS][NT], ix[2][NS][NT];
    for (int layout = 0; layout < 2; layout++)
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
            for (int layout = 0; layout < 2; layout++) {
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
                // prefer smaller bricks on ties (less shared memory, less L2 traffic)
                cost = cost * 4096 + (long)bw * bh / 8;
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

template <int INTERP, int RULE>
int launch2(const VtResampleParams &P, cudaStream_t st)
{
    VtBrickStaging G;
    memset(&G, 0, sizeof G);
    if (!brick_dims(P, INTERP, G.bw, G.bh, G.bd)) return VT_ERR_UNSUPPORTED;
    if (INTERP != VT_LINEAR) tune_brick(P, G);
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
        VT_CUDA(cudaFuncSetAttribute(vt_brick_kernel<INTERP, RULE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     128 + MAX_BRICK_BYTES));
        VT_CUDA(cudaFuncSetAttribute(vt_brick_kernel<INTERP, RULE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     128 + MAX_BRICK_BYTES));
        VT_CUDA(cudaFuncSetAttribute(vt_brick_kernel<INTERP, RULE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     128 + MAX_BRICK_BYTES));
        attr_set.store(true, std::memory_order_release);
    }
    {
        VtProf prof(VT_K_BRICK_LINEAR + INTERP, st);
        if (P.flags & VT_INTERNAL_PROJECT) vt_brick_kernel<INTERP, RULE, 2><<<grid, NT, smem, st>>>(P, G);
        else if (P.flags & VT_OOB_ZERO) vt_brick_kernel<INTERP, RULE, 1><<<grid, NT, smem, st>>>(P, G);
        else vt_brick_kernel<INTERP, RULE, 0><<<grid, NT, smem, st>>>(P, G);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
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
    return brick_dims(P, interp, bw, bh, bd) ? 1 : 0;
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
