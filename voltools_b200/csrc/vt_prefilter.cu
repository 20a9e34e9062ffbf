// vt_prefilter.cu -- cubic B-spline prefilter (samples -> interpolation coefficients), in place.
//
// Replaces _bspline_prefilter (voltools/transforms.py:290-309) and SamplesToCoefficients3DX/Y/Z +
// ConvertToInterpolationCoefficients (voltools/kernels/bspline.h:2-99).  Per line of N samples:
//     z = sqrt(3)-2, lambda = (1-z)(1-1/z)
//     c+[0] = lambda * (s[0] + sum_{n<min(12,N)} z^(n+1) s[n]);   c+[n] = lambda*s[n] + z*c+[n-1]
//     c[N-1] = z/(z-1) * c+[N-1];                                  c[n]  = z * (c[n+1] - c+[n])
// along X (fastest axis), then Y, then Z.
//
// Variant 1 ("sequential"): one thread per line, causal sweep then anticausal sweep through global memory,
//   with exactly the reference's compiled operation order -> bit-identical coefficients.  Any shape.
//   Y/Z passes are coalesced (adjacent threads = adjacent x); the X pass stages 32x32 tiles through shared
//   memory so global accesses are coalesced as well.  16 B/voxel/pass of traffic.
// Variant 2 ("windowed"): see vt_prefilter_win.cu.
#include "vt_common.cuh"

namespace {

// SASS immediates of the reference's prefilter cubin (0xbe8930a4, 0x40bfffff, 0x3e58658d)
__device__ constexpr float kPole = -0.26794922351837158203f;
__device__ constexpr float kNegPole = 0.26794922351837158203f;
__device__ constexpr float kLambda = 5.9999995231628417969f;
__device__ constexpr float kAnti = 0.21132488548755645752f;

// causal step as compiled: fma(s, lambda, -(prev * |pole|))
__device__ __forceinline__ float causal_step(float s, float prev)
{
    return __fmaf_rn(s, kLambda, -__fmul_rn(prev, kNegPole));
}
// anticausal step: pole * (next - c)
__device__ __forceinline__ float anticausal_step(float next, float c) { return __fmul_rn(kPole, __fsub_rn(next, c)); }

// Y / Z passes: thread per line, `lane` index runs along x (stride 1) so a warp touches contiguous memory.
//   line(i, j) starts at base + i*stride_outer + j ; elements are stride_line apart ; n elements.
__global__ void __launch_bounds__(256) prefilter_strided_seq(float *__restrict__ vol, int n, size_t stride_line,
                                                             int n_inner, int n_outer, size_t stride_outer)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n_inner || i >= n_outer) return;
    float *c = vol + (size_t)i * stride_outer + j;
    const int horizon = n < 12 ? n : 12;
    float zn = kPole, sum = c[0];
    for (int k = 0; k < horizon; k++) {
        sum = __fmaf_rn(zn, c[k * stride_line], sum);
        zn = __fmul_rn(zn, kPole);
    }
    float prev = __fmul_rn(kLambda, sum);
    c[0] = prev;
    for (int k = 1; k < n; k++) {
        prev = causal_step(c[k * stride_line], prev);
        c[k * stride_line] = prev;
    }
    prev = __fmul_rn(kAnti, prev);
    c[(size_t)(n - 1) * stride_line] = prev;
    for (int k = n - 2; k >= 0; k--) {
        prev = anticausal_step(prev, c[k * stride_line]);
        c[k * stride_line] = prev;
    }
}

// X pass: a warp owns 32 consecutive rows (row = a line along x).  It walks the rows in chunks of 32
// columns: the chunk is read coalesced (lane = column) into a padded shared tile, then lane r runs the
// recursion of row r over the 32 columns out of shared memory, and the tile is written back coalesced.
constexpr int XW = 4;  // warps per block
__global__ void __launch_bounds__(32 * XW) prefilter_x_seq(float *__restrict__ vol, int n, size_t n_rows)
{
    __shared__ float tile[XW][32][33];
    const int lane = threadIdx.x, w = threadIdx.y;
    const size_t row0 = ((size_t)blockIdx.x * XW + w) * 32;
    if (row0 >= n_rows) return;
    const int rows = (int)min((size_t)32, n_rows - row0);
    float(*t)[33] = tile[w];
    float *base = vol + row0 * n;
    const int nchunks = (n + 31) / 32;

    // ---- causal initialisation: horizon = min(12, n) samples of each row (first chunk) ----
    float prev = 0.0f;
    // forward sweep
    for (int ch = 0; ch < nchunks; ch++) {
        const int x0 = ch * 32;
        const int cols = min(32, n - x0);
        for (int r = 0; r < rows; r++)
            if (lane < cols) t[r][lane] = base[(size_t)r * n + x0 + lane];
        __syncwarp();
        if (lane < rows) {
            int k = 0;
            if (ch == 0) {
                const int horizon = n < 12 ? n : 12;
                float zn = kPole, sum = t[lane][0];
                for (int q = 0; q < horizon; q++) {
                    sum = __fmaf_rn(zn, t[lane][q], sum);
                    zn = __fmul_rn(zn, kPole);
                }
                prev = __fmul_rn(kLambda, sum);
                t[lane][0] = prev;
                k = 1;
            }
            for (; k < cols; k++) {
                prev = causal_step(t[lane][k], prev);
                t[lane][k] = prev;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; r++)
            if (lane < cols) base[(size_t)r * n + x0 + lane] = t[r][lane];
        __syncwarp();
    }
    // backward sweep
    for (int ch = nchunks - 1; ch >= 0; ch--) {
        const int x0 = ch * 32;
        const int cols = min(32, n - x0);
        for (int r = 0; r < rows; r++)
            if (lane < cols) t[r][lane] = base[(size_t)r * n + x0 + lane];
        __syncwarp();
        if (lane < rows) {
            int k = cols - 1;
            if (ch == nchunks - 1) {
                prev = __fmul_rn(kAnti, t[lane][k]);
                t[lane][k] = prev;
                k--;
            }
            for (; k >= 0; k--) {
                prev = anticausal_step(prev, t[lane][k]);
                t[lane][k] = prev;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; r++)
            if (lane < cols) base[(size_t)r * n + x0 + lane] = t[r][lane];
        __syncwarp();
    }
}

}  // namespace

int vt_prefilter_win(float *d_vol, int d0, int d1, int d2, cudaStream_t st);  // vt_prefilter_win.cu

int vt_prefilter_seq(float *d_vol, int d0, int d1, int d2, cudaStream_t st)
{
    const size_t D = d0, H = d1, W = d2;
    // X: rows = D*H lines of W
    {
        const size_t n_rows = D * H;
        const size_t blocks = (n_rows + 32 * XW - 1) / (32 * XW);
        if (blocks > 0x7fffffffull) return VT_ERR_UNSUPPORTED;
        VtProf prof(VT_K_PREFILTER_X, st);
        prefilter_x_seq<<<(unsigned)blocks, dim3(32, XW), 0, st>>>(d_vol, d2, n_rows);
        vt_count_launch();
    }
    // Y: for each z (outer), W columns (inner), line stride W, n = H
    {
        dim3 grid((unsigned)((W + 255) / 256), (unsigned)D);
        if (D > 65535) return VT_ERR_UNSUPPORTED;
        VtProf prof(VT_K_PREFILTER_Y, st);
        prefilter_strided_seq<<<grid, 256, 0, st>>>(d_vol, d1, W, d2, d0, H * W);
        vt_count_launch();
    }
    // Z: H*W columns (inner, contiguous), line stride H*W, n = D
    {
        const size_t cols = H * W;
        const size_t bx = (cols + 255) / 256;
        if (bx > 0x7fffffffull) return VT_ERR_UNSUPPORTED;
        // inner index runs over the whole (y, x) plane: treat the plane as one row of H*W columns
        VtProf prof(VT_K_PREFILTER_Z, st);
        prefilter_strided_seq<<<dim3((unsigned)bx, 1), 256, 0, st>>>(d_vol, d0, cols, (int)cols, 1, 0);
        vt_count_launch();
    }
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}
