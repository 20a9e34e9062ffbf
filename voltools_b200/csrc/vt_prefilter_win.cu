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
// kernel 1: X and Y passes, marching down a plane
// ---------------------------------------------------------------------------------------------------
constexpr int RB = 16;   // rows per step
constexpr int SEG = 16;  // X segment length
// powers of the pole: kPow[j] = z^j
__device__ constexpr float kPow[17] = {1.0000000000e+00f,  -2.6794922352e-01f, 7.1796786384e-02f,  -1.9237893163e-02f,
                                       5.1547785351e-03f,  -1.3812189059e-03f, 3.7009653334e-04f,  -9.9167078735e-05f,
                                       2.6571741746e-05f,  -7.1198775683e-06f, 1.9077656660e-06f,  -5.1118432885e-07f,
                                       1.3697144399e-07f,  -3.6701392062e-08f, 9.8341095049e-09f,  -2.6350420058e-09f,
                                       7.0605745940e-10f};

// X pass pieces for one (row, segment) task; LEN is the compile-time segment length for full segments (SEG) so
// that the recursions unroll completely and the pole powers become immediates, or 0 for the ragged last segment.
template <int LEN>
__device__ __forceinline__ float xseg_causal(float *t, int len, int seg, bool line_start, int sw)
{
    float v;
    if (seg == 0) v = line_start ? causal_init(t, sw, 1) : __fmul_rn(kWarm, t[0]);
    else v = __fmul_rn(t[0], kLambda);
    t[0] = v;
    if (LEN) {
#pragma unroll
        for (int k = 1; k < LEN; k++) {
            v = causal_step(t[k], v);
            t[k] = v;
        }
    } else {
        for (int k = 1; k < len; k++) {
            v = causal_step(t[k], v);
            t[k] = v;
        }
    }
    return v;
}

template <int LEN>
__device__ __forceinline__ void xseg_anticausal(float *t, int len, float carry, bool line_end)
{
    if (LEN) {
        // full segment: never the end of the line's last segment unless line_end (then len == SEG too)
        float c = t[LEN - 1];  // k = 15 >= K: no carry correction
        float u = line_end ? __fmul_rn(kAnti, c) : __fmul_rn(kPole, -c);
        t[LEN - 1] = u;
#pragma unroll
        for (int k = LEN - 2; k >= 0; k--) {
            c = k < K ? fmaf(kPow[k + 1], carry, t[k]) : t[k];
            u = anticausal_step(u, c);
            t[k] = u;
        }
    } else {
        int k = len - 1;
        float c = k < K ? fmaf(kPow[k + 1], carry, t[k]) : t[k];
        float u = line_end ? __fmul_rn(kAnti, c) : __fmul_rn(kPole, -c);
        t[k] = u;
        for (k = len - 2; k >= 0; k--) {
            c = k < K ? fmaf(kPow[k + 1], carry, t[k]) : t[k];
            u = anticausal_step(u, c);
            t[k] = u;
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) prefilter_xy_kernel(const float *__restrict__ src, float *__restrict__ dst, int H,
                                                          int W, long long dst_row, long long dst_plane, int y_chunk,
                                                          int x_strip, int pitch, int nseg_max)
{
    extern __shared__ float smem[];
    float *tile = smem;                 // [RB][pitch], pitch odd
    float *ends = smem + RB * pitch;    // [RB][nseg_max]: causal value at the end of each X segment
    const int z = blockIdx.z;
    const int x0 = blockIdx.x * x_strip, x1 = min(x0 + x_strip, W);  // columns written by this CTA
    const int xa = max(x0 - K, 0), xb = min(x1 + K, W);              // columns staged (X warm-up on both sides)
    const int sw = xb - xa;
    const int yc0 = blockIdx.y * y_chunk, yc1 = min(yc0 + y_chunk, H);  // rows written
    const int ra = max(yc0 - K, 0), rb = min(yc1 + K, H);               // rows processed (Y warm-up / look-ahead)
    const int tid = threadIdx.x;
    const int nseg = (sw + SEG - 1) / SEG;
    const int last_len = sw - (nseg - 1) * SEG;
    const bool stager = tid < sw;
    const bool has_col = x0 + tid < x1;   // this thread sweeps column x0 + tid along y
    const int cx = x0 + tid - xa;         // its column inside the tile
    const int npad = (x1 == W) ? (int)(dst_row - W) : 0;  // pad columns (written as zeros) belong to the last strip
    // running pointers
    const float *sp = src + (size_t)z * H * W + (size_t)ra * W + xa + tid;  // row being prefetched, this thread's column
    float *op = dst + (size_t)z * dst_plane + (long long)(ra - K) * dst_row + x0 + tid;  // row r0 - K of the output

    float pf[RB];
#pragma unroll
    for (int i = 0; i < RB; i++) pf[i] = (stager && ra + i < rb) ? __ldg(sp + (size_t)i * W) : 0.0f;
    sp += (size_t)RB * W;
    float cp[K + RB];  // causal Y values of rows [r0 - K, r0 + RB)
#pragma unroll
    for (int k = 0; k < K + RB; k++) cp[k] = 0.0f;
    float prev = 0.0f;

    for (int r0 = ra;; r0 += RB) {
        const int nrows = min(RB, rb - r0);  // <= 0 once the rows are exhausted (flush steps)
        __syncthreads();                     // the previous step's column sweep is done with the tile
        if (nrows > 0) {
            if (stager) {
#pragma unroll
                for (int i = 0; i < RB; i++) tile[i * pitch + tid] = pf[i];
            }
            if (r0 + 2 * RB <= rb) {  // uniform: the next step is a full one
                if (stager) {
#pragma unroll
                    for (int i = 0; i < RB; i++) pf[i] = __ldg(sp + (size_t)i * W);
                }
            } else {
#pragma unroll
                for (int i = 0; i < RB; i++) pf[i] = (stager && r0 + RB + i < rb) ? __ldg(sp + (size_t)i * W) : 0.0f;
            }
            sp += (size_t)RB * W;
            __syncthreads();
            // ---- X, causal: local recursion per (row, segment) ----
            for (int task = tid; task < RB * nseg; task += NT) {
                const int row = task & (RB - 1), seg = task >> 4;
                if (row >= nrows) continue;
                float *t = tile + row * pitch + seg * SEG;
                float v;
                if (seg < nseg - 1 || last_len == SEG) v = xseg_causal<SEG>(t, SEG, seg, xa == 0, sw);
                else v = xseg_causal<0>(t, last_len, seg, xa == 0, sw);
                ends[row * nseg_max + seg] = v;
            }
            __syncthreads();
            // ---- X, anticausal: local recursion on the carry-corrected causal values ----
            for (int task = tid; task < RB * nseg; task += NT) {
                const int row = task & (RB - 1), seg = task >> 4;
                if (row >= nrows) continue;
                float *t = tile + row * pitch + seg * SEG;
                const float carry = seg > 0 ? ends[row * nseg_max + seg - 1] : 0.0f;
                const bool line_end = seg == nseg - 1 && xb == W;
                if (seg < nseg - 1 || last_len == SEG) xseg_anticausal<SEG>(t, SEG, carry, line_end);
                else xseg_anticausal<0>(t, last_len, carry, line_end);
            }
            __syncthreads();
            // ---- X, anticausal carry: the first sample of the next segment feeds the last 12 of this one ----
            for (int task = tid; task < RB * (nseg - 1); task += NT) {
                const int row = task & (RB - 1), seg = task >> 4;
                if (row >= nrows) continue;
                float *t = tile + row * pitch + seg * SEG;
                const float carry = t[SEG];
#pragma unroll
                for (int k = SEG - K; k < SEG; k++) t[k] = fmaf(kPow[SEG - k], carry, t[k]);
            }
            __syncthreads();
        }
        // ---- Y: one thread per column, rows r0 .. r0+nrows-1 enter the window ----
        const int w0 = r0 - K;  // row of cp[0]
        if (has_col) {
            const float *c = tile + cx;
            if (r0 > ra && nrows == RB && w0 >= ra) {
                // steady state (uniform): a full step strictly inside the processed rows
#pragma unroll
                for (int k = 0; k < RB; k++) {
                    prev = causal_step(c[k * pitch], prev);
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
                    prev = ra == 0 ? causal_init(c, min(rb, RB), pitch) : __fmul_rn(kWarm, c[0]);
                    cp[K] = prev;
                    kstart = 1;
                }
#pragma unroll
                for (int k = 0; k < RB; k++) {
                    if (k >= kstart) {
                        if (r0 + k < rb) prev = causal_step(c[k * pitch], prev);
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
    // raw[] holds the next ZB planes of the look-ahead region; it is reloaded for the following step right after
    // the causal recursion has consumed it, so those loads are in flight during the anticausal sweep and the stores
    // (2 x ZB loads outstanding per thread: enough to cover the HBM latency at ~13 warps per SM)
    float raw[ZB];
#pragma unroll
    for (int k = 0; k < ZB; k++) {
        const int zz = zw + K + k;
        raw[k] = zz < D ? s[(size_t)zz * cols] : 0.0f;
    }
    for (; zw < zc1; zw += ZB) {
        // causal values of the next ZB planes (look-ahead region moves forward)
#pragma unroll
        for (int k = 0; k < ZB; k++) {
            const int zz = zw + K + k;
            if (zz < D) prev = causal_step(raw[k], prev);
            cp[K + k] = prev;
        }
        if (zw + ZB < zc1) {
#pragma unroll
            for (int k = 0; k < ZB; k++) {
                const int zz = zw + ZB + K + k;
                raw[k] = zz < D ? s[(size_t)zz * cols] : 0.0f;
            }
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

template <int NT>
int launch_xy(const float *d_src, float *d_dst, int D, int H, int W, long long dst_row, long long dst_plane, cudaStream_t st)
{
    const int x_strip = W <= NT ? W : NT - 2 * K;
    const int strips = (W + x_strip - 1) / x_strip;
    const int sw_max = W <= NT ? W : NT;
    const int pitch = sw_max | 1;
    const int nseg_max = (sw_max + SEG - 1) / SEG;
    const size_t smem = ((size_t)RB * pitch + (size_t)RB * nseg_max) * sizeof(float);
    // y-chunks: ~1000+ CTAs in flight, chunks of at least 64 rows (each pays 2*K rows of warm-up / look-ahead)
    int chunks = (1200 + D * strips - 1) / (D * strips);
    const int max_chunks = H / 64 > 0 ? H / 64 : 1;
    if (chunks > max_chunks) chunks = max_chunks;
    int y_chunk = (H + chunks - 1) / chunks;
    chunks = (H + y_chunk - 1) / y_chunk;
    if (D > 65535 || chunks > 65535) return VT_ERR_UNSUPPORTED;
    VtProf prof(VT_K_PREFILTER_FUSED, st);
    prefilter_xy_kernel<NT><<<dim3(strips, chunks, D), NT, smem, st>>>(d_src, d_dst, H, W, dst_row, dst_plane, y_chunk,
                                                                      x_strip, pitch, nseg_max);
    return VT_OK;
}

// src -> dst (src != dst), dst possibly with padded strides.
int vt_prefilter_win(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row, long long dst_plane,
                     cudaStream_t st)
{
    const int D = d0, H = d1, W = d2;
    if (dst_row - W > 32) return VT_ERR_UNSUPPORTED;
    int rc;
    if (W <= 128) rc = launch_xy<128>(d_src, d_dst, D, H, W, dst_row, dst_plane, st);
    else if (W <= 256) rc = launch_xy<256>(d_src, d_dst, D, H, W, dst_row, dst_plane, st);
    else rc = launch_xy<512>(d_src, d_dst, D, H, W, dst_row, dst_plane, st);
    if (rc) return rc;
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    {
        // columns of the (padded) plane: pad columns hold zeros and simply stay zero
        const size_t cols = (size_t)dst_plane;
        const size_t bx = (cols + Z_THREADS - 1) / Z_THREADS;
        if (bx > 0x7fffffffull) return VT_ERR_UNSUPPORTED;
        // the sweep runs in place (reads stay ahead of writes within a column), which rules out z-chunks: a
        // neighbouring chunk's warm-up would read planes this one has already overwritten.  One chunk.
        const int z_chunk = (D + ZB - 1) / ZB * ZB;
        VtProf prof(VT_K_PREFILTER_Z, st);
        prefilter_z_kernel<<<dim3((unsigned)bx, 1), Z_THREADS, 0, st>>>(d_dst, d_dst, D, cols, z_chunk);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}
