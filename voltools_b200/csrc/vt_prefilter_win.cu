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
//   kernel 1 (xy): a CTA stages (rows y0-K .. y1+K) x (all of x) of one z-plane in shared memory (coalesced
//       loads), runs the X recursion with one thread per row over the COMPLETE row (same operation order as the
//       reference -> bit-identical X pass), then the Y recursion with one thread per column over the staged rows
//       (warm-up rows above and below unless the strip touches the volume face, where the reference's exact
//       start formulas apply), and writes rows y0..y1 to the destination (coalesced).  src -> dst, out of place.
//   kernel 2 (z): one thread per (y, x) column sweeps along z ONCE: the causal value is carried in a register;
//       the anticausal recursion is restarted every B = 16 planes from K planes ahead with the reference's own
//       start formula c = z/(z-1)*c+ (exact at the last plane, a |z|^12-accurate stand-in elsewhere) over a
//       register window.  Reads run ahead of writes, so it works in place.
//
// Accuracy vs the sequential variant (vt_prefilter.cu, bit-identical to the reference): <= ~3e-7 of the
// coefficient range (tests/test_gpu_parity.py::test_prefilter).
#include "vt_common.cuh"

namespace {

__device__ constexpr float kPole = -0.26794922351837158203f;
__device__ constexpr float kNegPole = 0.26794922351837158203f;
__device__ constexpr float kLambda = 5.9999995231628417969f;
__device__ constexpr float kAnti = 0.21132488548755645752f;   // z / (z - 1)
// steady-state start for a causal warm-up: c+ ~ lambda * s / (1 - z) if the signal were constant
__device__ constexpr float kWarm = 5.9999995231628417969f / 1.26794922351837158203f;

constexpr int K = 12;  // warm-up / look-ahead length

__device__ __forceinline__ float causal_step(float s, float prev)
{
    return __fmaf_rn(s, kLambda, -__fmul_rn(prev, kNegPole));
}
__device__ __forceinline__ float anticausal_step(float next, float c) { return __fmul_rn(kPole, __fsub_rn(next, c)); }

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
// kernel 1: X and Y passes of one strip of one plane, through shared memory
// ---------------------------------------------------------------------------------------------------
constexpr int XY_THREADS = 256;

__global__ void __launch_bounds__(XY_THREADS) prefilter_xy_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                                                  int H, int W, int rows_per_strip, int pitch)
{
    extern __shared__ float tile[];  // [rows][pitch], pitch odd
    const int z = blockIdx.y;
    const int y0 = blockIdx.x * rows_per_strip;
    const int y1 = min(y0 + rows_per_strip, H);        // rows [y0, y1) are written
    const int ya = max(y0 - K, 0), yb = min(y1 + K, H);  // rows [ya, yb) are staged
    const int rows = yb - ya;
    const size_t plane = (size_t)z * H * W;
    const int tid = threadIdx.x;

    // stage: the strip is contiguous in memory (complete rows): fully coalesced, any W
    const float *g = src + plane + (size_t)ya * W;
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < rows; r += XY_THREADS / 32) {
        const float *gr = g + (size_t)r * W;
        float *tr = tile + r * pitch;
        for (int c = lane; c < W; c += 32) tr[c] = __ldg(gr + c);
    }
    __syncthreads();

    // X: one thread per staged row, the whole row, reference operation order
    for (int r = tid; r < rows; r += XY_THREADS) {
        float *c = tile + r * pitch;
        float prev = causal_init(c, W, 1);
        c[0] = prev;
#pragma unroll 8
        for (int k = 1; k < W; k++) {
            prev = causal_step(c[k], prev);
            c[k] = prev;
        }
        prev = __fmul_rn(kAnti, prev);
        c[W - 1] = prev;
#pragma unroll 8
        for (int k = W - 2; k >= 0; k--) {
            prev = anticausal_step(prev, c[k]);
            c[k] = prev;
        }
    }
    __syncthreads();

    // Y: one thread per column over the staged rows
    for (int x = tid; x < W; x += XY_THREADS) {
        float *c = tile + x;
        float prev;
        if (ya == 0) prev = causal_init(c, rows, pitch);                 // true start of the line
        else prev = __fmul_rn(kWarm, c[0]);                             // warm-up start, K rows early
        c[0] = prev;
#pragma unroll 8
        for (int k = 1; k < rows; k++) {
            prev = causal_step(c[k * pitch], prev);
            c[k * pitch] = prev;
        }
        // true end of the line: the reference's start formula is exact; otherwise it is the look-ahead start
        prev = __fmul_rn(kAnti, prev);
        c[(rows - 1) * pitch] = prev;
#pragma unroll 8
        for (int k = rows - 2; k >= 0; k--) {
            prev = anticausal_step(prev, c[k * pitch]);
            c[k * pitch] = prev;
        }
    }
    __syncthreads();

    float *o = dst + plane + (size_t)y0 * W;
    const int roff = y0 - ya;
    for (int r = warp; r < y1 - y0; r += XY_THREADS / 32) {
        float *orow = o + (size_t)r * W;
        const float *tr = tile + (r + roff) * pitch;
        for (int c = lane; c < W; c += 32) orow[c] = tr[c];
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel 2: Z pass, single sweep per column
// ---------------------------------------------------------------------------------------------------
constexpr int ZB = 16;  // planes emitted per step
constexpr int Z_THREADS = 128;

// src may equal dst (in place): no __restrict__ here
__global__ void __launch_bounds__(Z_THREADS) prefilter_z_kernel(const float *src, float *dst, int D, size_t cols,
                                                                int z_chunk)
{
    const size_t col = (size_t)blockIdx.x * Z_THREADS + threadIdx.x;
    if (col >= cols) return;
    const int zc0 = blockIdx.y * z_chunk;          // this CTA emits planes [zc0, zc1)
    const int zc1 = min(zc0 + z_chunk, D);
    const float *s = src + col;
    float *d = dst + col;

    float cp[K + ZB];  // causal values of planes [zw, zw + K + ZB)
    float prev;
    int zw;            // plane index of cp[0]
    // ---- start-up: fill cp[0..K) ----
    if (zc0 == 0) {
        // exact start of the line (InitialCausalCoefficient, bspline.h:2-19)
        const int horizon = D < 12 ? D : 12;
        float zn = kPole, sum = s[0];
        for (int k = 0; k < horizon; k++) {
            sum = __fmaf_rn(zn, s[(size_t)k * cols], sum);
            zn = __fmul_rn(zn, kPole);
        }
        prev = __fmul_rn(kLambda, sum);
        zw = 0;
        cp[0] = prev;
#pragma unroll
        for (int k = 1; k < K; k++) {
            const int zz = zw + k;
            if (zz < D) prev = causal_step(s[(size_t)zz * cols], prev);
            cp[k] = prev;
        }
    } else {
        // warm-up K planes before the chunk, then the first K planes of the chunk
        const int zs = zc0 - K;  // >= 0 because chunks are >= K planes long
        prev = __fmul_rn(kWarm, s[(size_t)zs * cols]);
#pragma unroll
        for (int k = 1; k < K; k++) prev = causal_step(s[(size_t)(zs + k) * cols], prev);
        zw = zc0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int zz = zw + k;
            if (zz < D) prev = causal_step(s[(size_t)zz * cols], prev);
            cp[k] = prev;
        }
    }
    // ---- steady state ----
    for (; zw < zc1; zw += ZB) {
        // causal values of the next ZB planes (look-ahead region moves forward)
        float raw[ZB];
#pragma unroll
        for (int k = 0; k < ZB; k++) {
            const int zz = zw + K + k;
            raw[k] = zz < D ? s[(size_t)zz * cols] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < ZB; k++) {
            const int zz = zw + K + k;
            if (zz < D) prev = causal_step(raw[k], prev);
            cp[K + k] = prev;
        }
        // anticausal from the last available plane of the window back to zw
        const int last = min(zw + K + ZB, D) - 1;  // plane index where the recursion (re)starts
        float c = 0.0f;
#pragma unroll
        for (int k = K + ZB - 1; k >= 0; k--) {
            const int zz = zw + k;
            if (zz == last) c = __fmul_rn(kAnti, cp[k]);
            else if (zz < last) c = anticausal_step(c, cp[k]);
            if (k < ZB && zz < zc1) d[(size_t)zz * cols] = c;
        }
#pragma unroll
        for (int k = 0; k < K; k++) cp[k] = cp[ZB + k];
    }
}

}  // namespace

int vt_prefilter_seq(float *d_vol, int d0, int d1, int d2, cudaStream_t st);  // vt_prefilter.cu

// src -> dst (src != dst).  Returns VT_ERR_UNSUPPORTED when the plane strip does not fit shared memory.
int vt_prefilter_win(const float *d_src, float *d_dst, int d0, int d1, int d2, cudaStream_t st)
{
    const int D = d0, H = d1, W = d2;
    const int pitch = W | 1;
    // rows per strip: as many as fit ~100 KB of shared memory (2 CTAs per SM), at least 8
    const size_t budget = 100 * 1024;
    long rows_fit = (long)(budget / ((size_t)pitch * 4)) - 2 * K;
    if (rows_fit >= H) rows_fit = H;
    if (rows_fit < 8) {
        rows_fit = (long)((220 * 1024) / ((size_t)pitch * 4)) - 2 * K;  // one CTA per SM
        if (rows_fit < 4) return VT_ERR_UNSUPPORTED;
        if (rows_fit > H) rows_fit = H;
    }
    int rps = (int)rows_fit;
    const int strips = (H + rps - 1) / rps;
    rps = (H + strips - 1) / strips;  // balance
    const int staged = (rps + 2 * K) < H ? (rps + 2 * K) : H;
    const size_t smem = (size_t)staged * pitch * 4;
    if (D > 65535) return VT_ERR_UNSUPPORTED;
    static bool attr_set = false;
    if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(prefilter_xy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    {
        VtProf prof(VT_K_PREFILTER_FUSED, st);
        prefilter_xy_kernel<<<dim3(strips, D), XY_THREADS, smem, st>>>(d_src, d_dst, H, W, rps, pitch);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    {
        const size_t cols = (size_t)H * W;
        const size_t bx = (cols + Z_THREADS - 1) / Z_THREADS;
        if (bx > 0x7fffffffull) return VT_ERR_UNSUPPORTED;
        // the sweep runs in place (reads stay ahead of writes within a column), which rules out z-chunks: a
        // neighbouring chunk's warm-up would read planes this one has already overwritten.  One chunk.
        int chunks = 1;
        int z_chunk = (D + chunks - 1) / chunks;
        z_chunk = (z_chunk + ZB - 1) / ZB * ZB;  // whole steps
        chunks = (D + z_chunk - 1) / z_chunk;
        VtProf prof(VT_K_PREFILTER_Z, st);
        prefilter_z_kernel<<<dim3((unsigned)bx, chunks), Z_THREADS, 0, st>>>(d_dst, d_dst, D, cols, z_chunk);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}
