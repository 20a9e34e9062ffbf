// vt_prefilter_win.cu -- fast cubic B-spline prefilter: two kernels, 16 B/voxel of HBM traffic in total.
//
// Replaces _bspline_prefilter (voltools/transforms.py:290-309) and SamplesToCoefficients3DX/Y/Z
// (voltools/kernels/bspline.h:58-99), which make three in-place passes with one thread per line and two
// sweeps per pass through global memory (48 B/voxel, the X pass uncoalesced).
//
// The recursion c+[n] = lambda*s[n] + z*c+[n-1], c[n] = z*(c[n+1] - c+[n]) has the pole z = sqrt(3)-2, so the
// influence of a start value decays as |z|^k: |z|^12 = 1.4e-7 is below float32 resolution (the reference itself
// truncates its causal start sum at 12 terms, bspline.h:7).  Starting a recursion K = 12 samples early
// ("warm-up") with an approximate start value therefore reproduces the full-line recursion to float32 rounding,
// which is what lets a line be cut into independently processed windows:
//
//   kernel 1 (xy): a CTA walks down one z-plane 16 rows at a time.  The 16 rows are staged in shared memory
//       (coalesced, register-prefetched one step ahead); the X recursion runs on them in 16-sample segments, one
//       (row, segment) task per thread: a local recursion with zero carry-in, then the carry of the previous
//       segment is added to the first 12 samples (|z|^k decay: the rest is below float32 resolution), for both
//       directions; the exact start formulas of the reference are used at the true line ends.  Then one thread
//       per column continues the Y recursion down the plane exactly like kernel 2 does along z (causal value in
//       a register, anticausal restart from 12 rows ahead over a register window) and stores finished rows
//       (coalesced).  Every sample is read once and written once.  src -> dst, out of place.
//   kernel 2 (z): one thread per (y, x) column sweeps along z ONCE: the causal value is carried in a register;
//       the anticausal recursion is restarted every B = 16 planes from K planes ahead with the reference's own
//       start formula c = z/(z-1)*c+ (exact at the last plane, a |z|^12-accurate stand-in elsewhere) over a
//       register window.  Reads run ahead of writes, so it works in place.
//
// Accuracy vs the sequential variant (vt_prefilter.cu, bit-identical to the reference): <= ~3e-7 of the
// coefficient range (tests/test_gpu_parity.py::test_prefilter).
#include <cstdlib>

#include "vt_common.cuh"

namespace {

__device__ constexpr float kPole = -0.26794922351837158203f;
__device__ constexpr float kNegPole = 0.26794922351837158203f;
__device__ constexpr float kLambda = 5.9999995231628417969f;
__device__ constexpr float kAnti = 0.21132488548755645752f;   // z / (z - 1)
// steady-state start for a causal warm-up: c+ ~ lambda * s / (1 - z) if the signal were constant
__device__ constexpr float kWarm = 5.9999995231628417969f / 1.26794922351837158203f;

constexpr int K = 12;  // warm-up / look-ahead length

// One step of each recursion with the multiply by the new sample OFF the dependent chain: the chain is one FMA (4
// cycles) per sample instead of the reference's FMUL -> FFMA / FSUB -> FMUL pairs (8).  Same value up to one
// rounding per step (these kernels are the windowed variant: 1e-6 of the range, not bit parity; variant 1 keeps the
// reference's exact operation order).
__device__ __forceinline__ float causal_step(float s, float prev)
{
    return __fmaf_rn(kPole, prev, __fmul_rn(s, kLambda));
}
__device__ __forceinline__ float anticausal_step(float next, float c)
{
    return __fmaf_rn(kPole, next, __fmul_rn(kNegPole, c));
}

// exact causal start of a line (InitialCausalCoefficient, bspline.h:2-19) over elements e[0], e[step], ...
__device__ __forceinline__ float causal_init(const float *e, int n, int step)
{
    const int horizon = n < 12 ? n : 12;
    float zn = kPole, sum = e[0];
    for (int k = 0; k < horizon; k++) {
        sum = __fmaf_rn(zn, e[k * step], sum);
        zn = __fmul_rn(zn, kPole);
    }
    return __fmul_rn(kLambda, sum);
}

// ---------------------------------------------------------------------------------------------------
// kernel 1: X and Y passes, marching down a plane
// ---------------------------------------------------------------------------------------------------
constexpr int RB = 16;   // rows per step
constexpr int SEG = 16;  // X chunk: samples per thread in the X pass
constexpr int NBUF = 4;  // staging tiles of the XY kernel: Y pass | X pass | in flight | being refilled
constexpr int HX = 16;   // X halo of interior strips (>= K; a multiple of SEG keeps the Y pass's warps chunk-aligned)
// powers of the pole: kPow[j] = z^j
__device__ constexpr float kPow[17] = {1.0000000000e+00f,  -2.6794922352e-01f, 7.1796786384e-02f,  -1.9237893163e-02f,
                                       5.1547785351e-03f,  -1.3812189059e-03f, 3.7009653334e-04f,  -9.9167078735e-05f,
                                       2.6571741746e-05f,  -7.1198775683e-06f, 1.9077656660e-06f,  -5.1118432885e-07f,
                                       1.3697144399e-07f,  -3.6701392062e-08f, 9.8341095049e-09f,  -2.6350420058e-09f,
                                       7.0605745940e-10f};

// Position of strip-local column lx inside a tile row.  A row is a sequence of 16-sample chunks; the four 16-byte
// units of chunk c are rotated by (c >> 1), which makes the X pass's LDS.128 / STS.128 (lane = chunk, i.e. a
// 64-byte lane stride) bank-conflict free while a warp of the Y pass (32 consecutive columns = 2 chunks with the
// same rotation) still reads a permutation of 32 consecutive words.
__device__ __forceinline__ int tile_col(int lx)
{
    const int c = lx >> 4, g = (lx >> 2) & 3;
    return (c << 4) + (((g + (c >> 1)) & 3) << 2) + (lx & 3);
}

// The X pass of one (row, chunk) task, entirely in registers.  Both recursions run locally with a zero carry-in;
// the carry of the neighbouring chunk (one warp shuffle) is then added to the 12 samples it can still reach
// (|z|^k decay: beyond 12 samples it is below float32 resolution, and a chunk's own end value is its true end
// value to |z|^16).  Line ends: the reference's exact causal start (bspline.h:2-19) for the chunk at x = 0, and the
// anticausal start c[W-1] = z/(z-1) c+[W-1] through the per-sample factor cf[] (-z inside the line, z/(z-1) on its
// last sample, 0 beyond it).
template <int L>
__device__ __forceinline__ void x_pass(float *tb, int xc, const float (&cf)[SEG], bool exact_start)
{
    const int rot = (xc >> 1) & 3;
    float t[SEG];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const float4 q = *reinterpret_cast<const float4 *>(tb + (((g + rot) & 3) << 2));
        t[4 * g] = q.x; t[4 * g + 1] = q.y; t[4 * g + 2] = q.z; t[4 * g + 3] = q.w;
    }
    // ---- causal ----
    float sum = t[0];
#pragma unroll
    for (int n = 0; n < 12; n++) sum = __fmaf_rn(kPow[n + 1], t[n], sum);  // samples past the line end are staged as 0
    float v = __fmul_rn(kLambda, xc == 0 ? (exact_start ? sum : __fmul_rn(kWarm / kLambda, t[0])) : t[0]);
    t[0] = v;
#pragma unroll
    for (int k = 1; k < SEG; k++) {
        v = __fmaf_rn(kPole, v, __fmul_rn(kLambda, t[k]));
        t[k] = v;
    }
    float carry = __shfl_up_sync(0xffffffffu, v, 1, L);
    if (xc == 0) carry = 0.0f;
#pragma unroll
    for (int k = 0; k < K; k++) t[k] = __fmaf_rn(kPow[k + 1], carry, t[k]);
    // ---- anticausal ----
    float u = 0.0f;
#pragma unroll
    for (int k = SEG - 1; k >= 0; k--) {
        u = __fmaf_rn(kPole, u, __fmul_rn(cf[k], t[k]));
        t[k] = u;
    }
    float nxt = __shfl_down_sync(0xffffffffu, u, 1, L);  // first sample of the next chunk (true value to |z|^16)
    if (xc == L - 1) nxt = 0.0f;
#pragma unroll
    for (int k = SEG - K; k < SEG; k++) t[k] = __fmaf_rn(kPow[SEG - k], nxt, t[k]);
#pragma unroll
    for (int g = 0; g < 4; g++)
        *reinterpret_cast<float4 *>(tb + (((g + rot) & 3) << 2)) = make_float4(t[4 * g], t[4 * g + 1], t[4 * g + 2], t[4 * g + 3]);
}

// One CTA walks down (a y-chunk of) one z-plane RB rows at a time.  Per step: the RB rows are staged into a
// shared-memory tile with cp.async (NBUF tiles: the next steps' rows are in flight during this step's math),
// the X recursion runs on them with one (row, 16-sample chunk) task per thread (x_pass), then one thread per
// column continues the Y recursion down the plane (causal value in a register, anticausal restart from K rows
// ahead over a register window) and stores finished rows (coalesced).  NT = RB * L threads.
template <int NT>
__global__ void __launch_bounds__(NT) prefilter_xy_kernel(const float *__restrict__ src, float *__restrict__ dst, int H,
                                                          int W, long long dst_row, long long dst_plane, int y_chunk,
                                                          int x_strip, int z_first)
{
    constexpr int L = NT / RB;   // chunks (X-pass lanes) per row
    constexpr int P = L * SEG;   // tile row pitch in floats (= NT)
    extern __shared__ __align__(16) float smem[];  // NBUF tiles [RB][P]
    vt_pdl_wait();
    const int z = z_first + blockIdx.z;
    const int x0 = blockIdx.x * x_strip, x1 = min(x0 + x_strip, W);  // columns written by this CTA
    const int xa = max(x0 - HX, 0), xb = min(x1 + HX, W);            // columns staged (X warm-up on both sides)
    const int sw = xb - xa;
    const int yc0 = blockIdx.y * y_chunk, yc1 = min(yc0 + y_chunk, H);  // rows written
    const int ra = max(yc0 - K, 0), rb = min(yc1 + K, H);               // rows processed (Y warm-up / look-ahead)
    const int tid = threadIdx.x;
    const bool stager = tid < sw;         // this thread stages column xa + tid of every row
    const bool has_col = x0 + tid < x1;   // this thread sweeps column x0 + tid along y
    const int npad = (x1 == W) ? (int)(dst_row - W) : 0;  // pad columns (written as zeros) belong to the last strip
    // X-pass task of this thread
    const int xrow = tid / L, xc = tid % L;
    float cf[SEG];
#pragma unroll
    for (int k = 0; k < SEG; k++) {
        const int x = xa + xc * SEG + k;
        cf[k] = x < W - 1 ? kNegPole : (x == W - 1 ? kAnti : 0.0f);
    }
    // running pointers
    const float *sp = src + (size_t)z * H * W + (size_t)ra * W + xa + tid;  // this thread's column, first row to stage
    float *op = dst + (size_t)z * dst_plane + (long long)(ra - K) * dst_row + x0 + tid;  // row r0 - K of the output
    const unsigned tile_s = vt_smem_u32(smem);
    const unsigned stage_dst = tile_s + 4u * (unsigned)tile_col(tid);
    const float *ycol = smem + tile_col(x0 + tid - xa);

    // rows r0 .. r0+RB-1 -> tile[buf]; rows past rb and columns past sw: zeros.  Always one commit group per call
    // (empty once the rows are exhausted) so that "all but the newest NBUF-2 groups" is the step being consumed.
    auto stage = [&](int r0, int buf) {
        if (r0 < rb) {
#pragma unroll
            for (int i = 0; i < RB; i++)
                vt_cp_async4(stage_dst + 4u * (unsigned)((buf * RB + i) * P), sp + (size_t)i * W,
                             (stager && r0 + i < rb) ? 1u : 0u);
        }
        vt_cp_async_commit();
        sp += (size_t)RB * W;
    };

    // The X pass runs ONE STEP AHEAD of the Y pass (software pipeline): in an iteration every thread filters its
    // (row, chunk) of tile k+1 along x and then sweeps its column of tile k along y, so a step needs one block-wide
    // barrier instead of two and the two passes of different warps overlap (the kernel was waiting on its barriers, not
    // on issue slots: packing the Y pass in FFMA2 changed nothing).  Tiles: k (Y) | k+1 (X) | k+2 (in flight) | k+3
    // (being refilled): two steps of rows are in flight ahead of the X pass -- with 1-2 CTAs per SM a single step
    // (16 KB) does not cover the HBM latency.
#pragma unroll
    for (int i = 0; i < NBUF - 1; i++) stage(ra + i * RB, i);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NBUF - 2) : "memory");
    __syncthreads();  // tile 0 has landed
    if (rb - ra > 0) x_pass<L>(smem + xrow * P + xc * SEG, xc, cf, xa == 0);
    float cp[K + RB];  // causal Y values of rows [r0 - K, r0 + RB)
#pragma unroll
    for (int k = 0; k < K + RB; k++) cp[k] = 0.0f;
    float prev = 0.0f;
    int buf = 0;

    for (int r0 = ra;; r0 += RB, buf = buf + 1 == NBUF ? 0 : buf + 1) {
        const int nrows = min(RB, rb - r0);  // <= 0 once the rows are exhausted (flush steps)
        const int nb = buf + 1 == NBUF ? 0 : buf + 1;  // tile of the next step
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NBUF - 3) : "memory");
        __syncthreads();  // tile[nb] has landed; tile[buf] is X-filtered; the previous step's column sweep is done
        stage(r0 + (NBUF - 1) * RB, buf == 0 ? NBUF - 1 : buf - 1);  // refills the previous step's tile
        if (rb - (r0 + RB) > 0) x_pass<L>(smem + (nb * RB + xrow) * P + xc * SEG, xc, cf, xa == 0);
        // ---- Y: one thread per column, rows r0 .. r0+nrows-1 enter the window ----
        const int w0 = r0 - K;  // row of cp[0]
        if (has_col) {
            const float *c = ycol + buf * RB * P;
            if (r0 > ra && nrows == RB && w0 >= ra) {
                // steady state (uniform): a full step strictly inside the processed rows
#pragma unroll
                for (int k = 0; k < RB; k++) {
                    prev = causal_step(c[k * P], prev);
                    cp[K + k] = prev;
                }
                float a = __fmul_rn(kAnti, cp[K + RB - 1]);
                if (w0 >= yc0 && w0 + RB <= yc1) {
#pragma unroll
                    for (int k = K + RB - 2; k >= 0; k--) {
                        a = anticausal_step(a, cp[k]);
                        if (k < RB) op[(size_t)k * dst_row] = a;
                    }
                } else {
#pragma unroll
                    for (int k = K + RB - 2; k >= 0; k--) {
                        a = anticausal_step(a, cp[k]);
                        if (k < RB && w0 + k >= yc0 && w0 + k < yc1) op[(size_t)k * dst_row] = a;
                    }
                }
            } else {
                int kstart = 0;
                if (r0 == ra) {  // first row of the line (true start: exact formula) or of the warm-up
                    prev = ra == 0 ? causal_init(c, min(rb, RB), P) : __fmul_rn(kWarm, c[0]);
                    cp[K] = prev;
                    kstart = 1;
                }
#pragma unroll
                for (int k = 0; k < RB; k++) {
                    if (k >= kstart) {
                        if (r0 + k < rb) prev = causal_step(c[k * P], prev);
                        cp[K + k] = prev;
                    }
                }
                const int last = min(r0 + RB, rb) - 1;  // row where the anticausal recursion (re)starts
                float a = 0.0f;
#pragma unroll
                for (int k = K + RB - 1; k >= 0; k--) {
                    const int y = w0 + k;
                    if (y == last) a = __fmul_rn(kAnti, cp[k]);
                    else if (y < last && y >= ra) a = anticausal_step(a, cp[k]);
                    if (k < RB && y >= yc0 && y < yc1) op[(size_t)k * dst_row] = a;
                }
            }
#pragma unroll
            for (int k = 0; k < K; k++) cp[k] = cp[RB + k];
        }
        if (tid < npad) {
            float *pp = op + (W - x0);  // column W + tid
            for (int k = 0; k < RB; k++) {
                const int y = w0 + k;
                if (y >= yc0 && y < yc1) pp[(size_t)k * dst_row] = 0.0f;
            }
        }
        op += (size_t)RB * dst_row;
        if (w0 + RB >= yc1) break;
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel 2: Z pass, single sweep per column
// ---------------------------------------------------------------------------------------------------
constexpr int ZB = 16;  // planes emitted per step
constexpr int Z_THREADS = 128;

// The sweep is instruction bound (ncu: 83 % issue-active), so a thread takes TWO adjacent columns as a packed pair
// (V = vt_f2): 64-bit loads and stores and FFMA2 / FMUL2 (two IEEE fp32 operations per lane and instruction) halve
// the instruction count per voxel.  V = float is the same code for planes with an odd number of columns.
template <typename V>
struct ZOps;
template <>
struct ZOps<float> {
    static __device__ __forceinline__ float zero() { return 0.0f; }
    static __device__ __forceinline__ float mulc(float c, float x) { return __fmul_rn(c, x); }
    static __device__ __forceinline__ float fmac(float c, float x, float y) { return __fmaf_rn(c, x, y); }
};
template <>
struct ZOps<vt_f2> {
    static __device__ __forceinline__ vt_f2 zero() { return 0ull; }
    static __device__ __forceinline__ vt_f2 mulc(float c, vt_f2 x)
    {
        vt_f2 r;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(vt_pk(c, c)), "l"(x));
        return r;
    }
    static __device__ __forceinline__ vt_f2 fmac(float c, vt_f2 x, vt_f2 y) { return vt_fma2(vt_pk(c, c), x, y); }
};
template <typename V>
__device__ __forceinline__ V causal_step_v(V s, V prev)
{
    return ZOps<V>::fmac(kPole, prev, ZOps<V>::mulc(kLambda, s));
}
template <typename V>
__device__ __forceinline__ V anticausal_step_v(V next, V c)
{
    return ZOps<V>::fmac(kPole, next, ZOps<V>::mulc(kNegPole, c));
}

// Z4 output (vt_resample_z4.cu): dst4[g][column] = float4 of planes 4g .. 4g+3 of that column (zeros past the last plane)
template <typename V>
__device__ __forceinline__ void store_z4(float4 *d4, size_t group, size_t fcols, size_t col, const V (&o)[4]);
template <>
__device__ __forceinline__ void store_z4<float>(float4 *d4, size_t group, size_t fcols, size_t col, const float (&o)[4])
{
    d4[group * fcols + col] = make_float4(o[0], o[1], o[2], o[3]);
}
template <>
__device__ __forceinline__ void store_z4<vt_f2>(float4 *d4, size_t group, size_t fcols, size_t col, const vt_f2 (&o)[4])
{
    float a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; k++) vt_unpk(o[k], a[k], b[k]);
    float4 *q = d4 + group * fcols + 2 * col;  // the pair's two columns are adjacent: 32 contiguous bytes
    q[0] = make_float4(a[0], a[1], a[2], a[3]);
    q[1] = make_float4(b[0], b[1], b[2], b[3]);
}

// src may equal dst (in place): no __restrict__ here.  `cols` = columns per plane in units of V.
// Z4OUT: dst is the Z4 layout of axis 0 (out of place only; z_begin and z_chunk are multiples of 4).
template <typename V, bool Z4OUT>
__global__ void __launch_bounds__(Z_THREADS) prefilter_z_kernel(const V *src, V *dst, int D, size_t cols, int z_chunk,
                                                                int z_begin, int z_end)
{
    using O = ZOps<V>;
    constexpr size_t VW = sizeof(V) / sizeof(float);
    float4 *d4 = (float4 *)dst;
    vt_pdl_wait();
    const size_t fcols = cols * VW;  // float columns per plane
    const size_t col = (size_t)blockIdx.x * Z_THREADS + threadIdx.x;
    if (col >= cols) return;
    const int zc0 = z_begin + blockIdx.y * z_chunk;  // this CTA emits planes [zc0, zc1)
    const int zc1 = min(zc0 + z_chunk, z_end);
    const V *s = src + col;
    V *d = dst + col;

    V cp[K + ZB];  // causal values of planes [zw, zw + K + ZB)
    V prev;
    int zw;            // plane index of cp[0]
    // ---- start-up: fill cp[0..K) ----
    if (zc0 <= K) {
        // the start of the line is within reach: exact start (InitialCausalCoefficient, bspline.h:2-19), then the
        // plain recursion up to the chunk
        V first[12];  // loaded together (independent), then summed in the reference's order
#pragma unroll
        for (int k = 0; k < 12; k++) first[k] = k < D ? s[(size_t)k * cols] : O::zero();
        float zn = kPole;
        V sum = first[0];
#pragma unroll
        for (int k = 0; k < 12; k++) {
            if (k < D) sum = O::fmac(zn, first[k], sum);
            zn = __fmul_rn(zn, kPole);
        }
        prev = O::mulc(kLambda, sum);
        for (int zz = 1; zz < zc0; zz++) prev = causal_step_v<V>(s[(size_t)zz * cols], prev);
        zw = zc0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int zz = zw + k;
            if (zz > 0 && zz < D) prev = causal_step_v<V>(s[(size_t)zz * cols], prev);
            cp[k] = prev;
        }
    } else {
        // warm-up K planes before the chunk, then the first K planes of the chunk
        const int zs = zc0 - K;  // > 0
        prev = O::mulc(kWarm, s[(size_t)zs * cols]);
#pragma unroll
        for (int k = 1; k < K; k++) prev = causal_step_v<V>(s[(size_t)(zs + k) * cols], prev);
        zw = zc0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int zz = zw + k;
            if (zz < D) prev = causal_step_v<V>(s[(size_t)zz * cols], prev);
            cp[k] = prev;
        }
    }
    // ---- steady state ----
    // raw[] holds the next ZB planes of the look-ahead region; it is reloaded for the following step right after
    // the causal recursion has consumed it, so those loads are in flight during the anticausal sweep and the stores
    // (2 x ZB loads outstanding per thread).  Steps whose whole window lies inside the volume and the chunk take a
    // guard-free path with one IMAD.WIDE per address (32-bit plane stride).
    const unsigned stride = (unsigned)cols;  // host: cols < 2^31
    const V *sp = s + (size_t)(zw + K) * cols;  // plane zw + K of this column
    V *dp = d + (size_t)zw * cols;              // plane zw
    V raw[ZB];
#pragma unroll
    for (int k = 0; k < ZB; k++) raw[k] = zw + K + k < D ? sp[(size_t)k * stride] : O::zero();
    for (; zw < zc1; zw += ZB) {
        sp += (size_t)ZB * stride;
        if (zw + K + ZB <= D && zw + ZB <= zc1) {  // uniform
#pragma unroll
            for (int k = 0; k < ZB; k++) {
                prev = causal_step_v<V>(raw[k], prev);
                cp[K + k] = prev;
            }
            if (zw + K + 2 * ZB <= D) {
#pragma unroll
                for (int k = 0; k < ZB; k++) raw[k] = sp[(size_t)k * stride];
            } else if (zw + ZB < zc1) {
#pragma unroll
                for (int k = 0; k < ZB; k++) raw[k] = zw + ZB + K + k < D ? sp[(size_t)k * stride] : O::zero();
            }
            V c = O::mulc(kAnti, cp[K + ZB - 1]);
            if constexpr (Z4OUT) {
                V o[4];
#pragma unroll
                for (int k = K + ZB - 2; k >= 0; k--) {
                    c = anticausal_step_v<V>(c, cp[k]);
                    if (k < ZB) {
                        o[k & 3] = c;
                        if ((k & 3) == 0) store_z4<V>(d4, (size_t)((zw + k) >> 2), fcols, col, o);
                    }
                }
            } else {
#pragma unroll
                for (int k = K + ZB - 2; k >= 0; k--) {
                    c = anticausal_step_v<V>(c, cp[k]);
                    if (k < ZB) dp[(size_t)k * stride] = c;
                }
            }
        } else {
            // causal values of the next ZB planes (look-ahead region moves forward)
#pragma unroll
            for (int k = 0; k < ZB; k++) {
                const int zz = zw + K + k;
                if (zz < D) prev = causal_step_v<V>(raw[k], prev);
                cp[K + k] = prev;
            }
            if (zw + ZB < zc1) {
#pragma unroll
                for (int k = 0; k < ZB; k++) raw[k] = zw + ZB + K + k < D ? sp[(size_t)k * stride] : O::zero();
            }
            // anticausal from the last available plane of the window back to zw
            const int last = min(zw + K + ZB, D) - 1;  // plane index where the recursion (re)starts
            V c = O::zero();
            V o[4];
#pragma unroll
            for (int k = K + ZB - 1; k >= 0; k--) {
                const int zz = zw + k;
                if (zz == last) c = O::mulc(kAnti, cp[k]);
                else if (zz < last) c = anticausal_step_v<V>(c, cp[k]);
                if constexpr (Z4OUT) {
                    if (k < ZB) {
                        o[k & 3] = zz < zc1 ? c : O::zero();  // planes past the end of the volume: zeros
                        if ((k & 3) == 0 && zz < zc1) store_z4<V>(d4, (size_t)(zz >> 2), fcols, col, o);
                    }
                } else {
                    if (k < ZB && zz < zc1) dp[(size_t)k * stride] = c;
                }
            }
        }
        dp += (size_t)ZB * stride;
#pragma unroll
        for (int k = 0; k < K; k++) cp[k] = cp[ZB + k];
    }
}

}  // namespace

int vt_prefilter_seq(float *d_vol, int d0, int d1, int d2, cudaStream_t st);  // vt_prefilter.cu

template <int NT>
int launch_xy(const float *d_src, float *d_dst, int D, int H, int W, long long dst_row, long long dst_plane, int z0, int z1,
              cudaStream_t st)
{
    const int x_strip = W <= NT ? W : NT - 2 * HX;
    const int strips = (W + x_strip - 1) / x_strip;
    const size_t smem = (size_t)NBUF * RB * NT * sizeof(float);
    const int nz = z1 - z0;
    // y-chunks: enough CTAs to fill the GPU, chunks of at least 64 rows (each pays 2*K rows of warm-up / look-ahead)
    int chunks = (200 + nz * strips - 1) / (nz * strips);
    const int max_chunks = H / 64 > 0 ? H / 64 : 1;
    if (chunks > max_chunks) chunks = max_chunks;
    if (const char *e = getenv("VT_XY_CHUNKS")) chunks = atoi(e) > 0 ? atoi(e) : chunks;  // tuning knob
    int y_chunk = (H + chunks - 1) / chunks;
    chunks = (H + y_chunk - 1) / y_chunk;
    if (nz > 65535 || chunks > 65535) return VT_ERR_UNSUPPORTED;
    // per device, so not cached in a static: a process may drive several GPUs
    VT_CUDA(cudaFuncSetAttribute(prefilter_xy_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VtProf prof(VT_K_PREFILTER_FUSED, st);
    VT_CUDA(vt_launch_pdl(prefilter_xy_kernel<NT>, dim3(strips, chunks, nz), dim3(NT), smem, st, d_src, d_dst, H, W, dst_row,
                          dst_plane, y_chunk, x_strip, z0));
    return VT_OK;
}

// X and Y passes of planes [z0, z1): d_src (dense) -> d_dst (padded strides).  Planes are independent.
int vt_prefilter_xy_range(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row, long long dst_plane,
                          int z0, int z1, cudaStream_t st)
{
    if (dst_row - d2 > 32) return VT_ERR_UNSUPPORTED;
    if (z1 <= z0) return VT_OK;
    int rc;
    if (d2 <= 128) rc = launch_xy<128>(d_src, d_dst, d0, d1, d2, dst_row, dst_plane, z0, z1, st);
    else if (d2 <= 256) rc = launch_xy<256>(d_src, d_dst, d0, d1, d2, dst_row, dst_plane, z0, z1, st);
    else rc = launch_xy<512>(d_src, d_dst, d0, d1, d2, dst_row, dst_plane, z0, z1, st);
    if (rc) return rc;
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

// Z pass producing planes [z0, z1) of d_dst from the XY-filtered volume d_src (depth d0, `cols` columns per plane).
// It reads planes [z0 - K, z1 + K) of d_src and nothing else (from plane 0 when z0 <= K).  d_src == d_dst (in place) is only valid for the whole range in one
// chunk; out of place the range is cut into `chunks` z-chunks (0 = choose).
static int z_range_impl(const float *d_src, float *d_dst, int d0, size_t cols, int z0, int z1, int chunks, bool z4out,
                        cudaStream_t st);
int vt_prefilter_z_range(const float *d_src, float *d_dst, int d0, size_t cols, int z0, int z1, int chunks, cudaStream_t st)
{
    return z_range_impl(d_src, d_dst, d0, cols, z0, z1, chunks, false, st);
}
// the same sweep writing the Z4 layout of axis 0 (vt_resample_z4.cu): d_dst4 holds ceil(d0/4) groups of `cols` float4
// (cols = the dense plane d1*d2 here: the XY-filtered source must be dense too).  z0 is a multiple of 4 and z1 a
// multiple of 4 or d0; out of place only.
int vt_prefilter_z_range_z4(const float *d_src, float *d_dst4, int d0, size_t cols, int z0, int z1, int chunks,
                            cudaStream_t st)
{
    if ((z0 & 3) || ((z1 & 3) && z1 != d0) || d_src == d_dst4) return VT_ERR_INVALID_ARG;
    return z_range_impl(d_src, d_dst4, d0, cols, z0, z1, chunks, true, st);
}
static int z_range_impl(const float *d_src, float *d_dst, int d0, size_t cols, int z0, int z1, int chunks, bool z4out,
                        cudaStream_t st)
{
    if (z1 <= z0) return VT_OK;
    // planes the kernel may read: everything up to K past the range (a pipelined caller has not produced the
    // rest of d_src yet); the anticausal restart at that artificial end is the usual |z|^12-accurate stand-in
    const int d_avail = z1 + K < d0 ? z1 + K : d0;
    // two columns per thread (packed pairs) when the planes allow 8-byte accesses
    static const bool no_pack = getenv("VT_Z_SCALAR") != nullptr;  // A/B knob
    const bool pack = !no_pack && cols % 2 == 0 && ((uintptr_t)d_src % 8) == 0 && ((uintptr_t)d_dst % (z4out ? 16 : 8)) == 0;
    const size_t vcols = pack ? cols / 2 : cols;
    const size_t bx = (vcols + Z_THREADS - 1) / Z_THREADS;
    if (bx > 0x7fffffffull || cols > 0x7fffffffull) return VT_ERR_UNSUPPORTED;
    const int nz = z1 - z0;
    if (d_src == d_dst) {
        if (z0 != 0 || z1 != d0) return VT_ERR_INVALID_ARG;
        chunks = 1;
    } else if (chunks <= 0) {
        // aim at >= ~300 000 threads, chunks of at least 64 planes (each pays 2*K planes of warm-up / look-ahead)
        chunks = (int)((300000 + vcols - 1) / vcols);
        const int max_chunks = nz / 64 > 0 ? nz / 64 : 1;
        if (chunks > max_chunks) chunks = max_chunks;
        if (const char *e = getenv("VT_Z_CHUNKS")) chunks = atoi(e) > 0 ? atoi(e) : chunks;  // tuning knob
    }
    int z_chunk = ((nz + chunks - 1) / chunks + ZB - 1) / ZB * ZB;
    chunks = (nz + z_chunk - 1) / z_chunk;
    if (chunks > 65535) return VT_ERR_UNSUPPORTED;
    {
        VtProf prof(VT_K_PREFILTER_Z, st);
        const dim3 grid((unsigned)bx, chunks);
        cudaError_t e;
        if (pack && z4out)
            e = vt_launch_pdl(prefilter_z_kernel<vt_f2, true>, grid, dim3(Z_THREADS), 0, st, (const vt_f2 *)d_src, (vt_f2 *)d_dst,
                              d_avail, vcols, z_chunk, z0, z1);
        else if (pack)
            e = vt_launch_pdl(prefilter_z_kernel<vt_f2, false>, grid, dim3(Z_THREADS), 0, st, (const vt_f2 *)d_src, (vt_f2 *)d_dst,
                              d_avail, vcols, z_chunk, z0, z1);
        else if (z4out)
            e = vt_launch_pdl(prefilter_z_kernel<float, true>, grid, dim3(Z_THREADS), 0, st, d_src, d_dst, d_avail, cols, z_chunk,
                              z0, z1);
        else
            e = vt_launch_pdl(prefilter_z_kernel<float, false>, grid, dim3(Z_THREADS), 0, st, d_src, d_dst, d_avail, cols, z_chunk,
                              z0, z1);
        VT_CUDA(e);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

// src -> dst (src != dst), dst possibly with padded strides.  With a workspace of d0 * dst_plane floats the XY kernel
// writes there and the Z sweep runs out of place, which allows it to be cut into z-chunks (more threads in flight:
// what a 250^3 volume, with only 63 000 columns, needs to cover the HBM latency); without one it runs in place.
int vt_prefilter_win(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row, long long dst_plane,
                     float *d_ws, size_t ws_bytes, cudaStream_t st)
{
    const size_t cols = (size_t)dst_plane;  // columns of the (padded) plane: pad columns hold zeros and stay zero
    const bool have_ws = d_ws && ws_bytes >= cols * (size_t)d0 * sizeof(float) && d_ws != d_dst && d_ws != d_src;
    float *xy_out = have_ws ? d_ws : d_dst;
    int rc = vt_prefilter_xy_range(d_src, xy_out, d0, d1, d2, dst_row, dst_plane, 0, d0, st);
    if (rc) return rc;
    return vt_prefilter_z_range(xy_out, d_dst, d0, cols, 0, d0, 0, st);
}
