/*
 * ref_harness.cu -- drives the REFERENCE's own CUDA kernels on a real GPU.
 *
 * TEST INFRASTRUCTURE ONLY (oracle).  Built by oracle/build_ref.py into oracle/_ref/libvt_ref_gpu.so.
 * The kernels themselves are NOT in this file and NOT in this repository: build_ref.py captures the
 * source strings the reference hands to cupy.RawKernel (voltools/transforms.py:237-284, :293-296),
 * compiles them unmodified with nvcc against the reference's headers where they lie
 * (/root/reference/voltools/kernels) and leaves the cubins in the git-ignored oracle/_ref/.
 *
 * This harness only replays, with the CUDA runtime/driver API, what the reference's Python does with
 * CuPy around those kernels:
 *   texture object .......... transforms.py:184-192 (float32 channel, 3-D CUDA array of extent (W,H,D),
 *                             border addressing on all axes, linear filter, element read, unnormalised)
 *   prefilter launches ...... transforms.py:301-307 + utils/general.py:9-33 (pow-2-divisor launch dims)
 *   array upload ............ transforms.py:197-199
 *   transform launch ........ transforms.py:203-212 + utils/general.py:36-58
 *   output=None semantics ... transforms.py:207-210 (zero-filled result)
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "ref_harness: %s failed: %s\n", #x, cudaGetErrorString(e_));      \
            return 100 + (int)e_;                                                              \
        }                                                                                      \
    } while (0)
#define CKD(x)                                                                                 \
    do {                                                                                       \
        CUresult e_ = (x);                                                                     \
        if (e_ != CUDA_SUCCESS) {                                                              \
            const char *s_ = 0;                                                                \
            cuGetErrorString(e_, &s_);                                                         \
            fprintf(stderr, "ref_harness: %s failed: %s\n", #x, s_ ? s_ : "?");               \
            return 1000 + (int)e_;                                                             \
        }                                                                                      \
    } while (0)

static char g_dir[4096] = "";
static CUmodule g_mod[4] = {0, 0, 0, 0}; /* 0 linear, 1 cubic, 2 cubic_simple, 3 prefilter */
static const char *g_names[4] = {"transform_linear.cubin", "transform_cubic.cubin",
                                 "transform_cubic_simple.cubin", "prefilter.cubin"};

static int load_module(int which)
{
    if (g_mod[which]) return 0;
    char path[4200];
    snprintf(path, sizeof path, "%s/%s", g_dir, g_names[which]);
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "ref_harness: cannot open %s\n", path);
        return 2;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void *buf = malloc(n);
    if (fread(buf, 1, n, f) != (size_t)n) return 3;
    fclose(f);
    CK(cudaFree(0)); /* make sure the primary context is current */
    CKD(cuModuleLoadData(&g_mod[which], buf));
    free(buf);
    return 0;
}

/* utils/general.py:9-33 */
static unsigned pow2_divider(unsigned n)
{
    if (n == 0) return 0;
    unsigned d = 1;
    while ((n & d) == 0) d <<= 1;
    return d;
}
static unsigned u_min(unsigned a, unsigned b) { return a < b ? a : b; }

/* utils/general.py:36-58 */
static void elementwise_dims(size_t n, int sm_count, unsigned *blocks, unsigned *threads)
{
    const size_t min_threads = 32, max_threads = 128;
    const size_t max_blocks = (size_t)4 * 8 * sm_count;
    if (n < min_threads) {
        *blocks = 1;
        *threads = (unsigned)min_threads;
    } else if (n < max_blocks * min_threads) {
        *blocks = (unsigned)((n + min_threads - 1) / min_threads);
        *threads = (unsigned)min_threads;
    } else if (n < max_blocks * max_threads) {
        *blocks = (unsigned)max_blocks;
        size_t grp = (n + min_threads - 1) / min_threads;
        *threads = (unsigned)(((grp + max_blocks - 1) / max_blocks) * min_threads);
    } else {
        *blocks = (unsigned)max_blocks;
        *threads = (unsigned)max_threads;
    }
}

/* transforms.py:290-309 on a device buffer, in place */
static int prefilter_device(float *d_vol, int D, int H, int W, cudaStream_t st)
{
    int rc = load_module(3);
    if (rc) return rc;
    CUfunction fx, fy, fz;
    CKD(cuModuleGetFunction(&fx, g_mod[3], "SamplesToCoefficients3DX"));
    CKD(cuModuleGetFunction(&fy, g_mod[3], "SamplesToCoefficients3DY"));
    CKD(cuModuleGetFunction(&fz, g_mod[3], "SamplesToCoefficients3DZ"));
    unsigned dim_x = u_min(u_min(pow2_divider(W), pow2_divider(H)), 64);
    unsigned dim_y = u_min(u_min(pow2_divider(D), pow2_divider(H)), 512 / dim_x);
    unsigned pitch = (unsigned)(W * sizeof(float)); /* volume.strides[1] */
    int shape[3] = {W, H, D};                     /* volume.shape[::-1] as int32 -> uint3 */
    int *d_shape;
    CK(cudaMalloc(&d_shape, sizeof shape));
    CK(cudaMemcpyAsync(d_shape, shape, sizeof shape, cudaMemcpyHostToDevice, st));
    void *args[3] = {&d_vol, &pitch, &d_shape};
    CKD(cuLaunchKernel(fx, H / dim_x, D / dim_y, 1, dim_x, dim_y, 1, 0, st, args, 0));
    CKD(cuLaunchKernel(fy, W / dim_x, D / dim_y, 1, dim_x, dim_y, 1, 0, st, args, 0));
    CKD(cuLaunchKernel(fz, W / dim_x, H / dim_y, 1, dim_x, dim_y, 1, 0, st, args, 0));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_shape));
    return 0;
}

struct RefTex {
    cudaArray_t arr;
    cudaTextureObject_t tex;
};

/* transforms.py:184-199 */
static int make_texture(const float *d_vol, int D, int H, int W, RefTex *out)
{
    cudaChannelFormatDesc ch = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    CK(cudaMalloc3DArray(&out->arr, &ch, make_cudaExtent(W, H, D)));
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr((void *)d_vol, W * sizeof(float), W, H);
    cp.dstArray = out->arr;
    cp.extent = make_cudaExtent(W, H, D);
    cp.kind = cudaMemcpyDeviceToDevice;
    CK(cudaMemcpy3D(&cp));
    cudaResourceDesc res;
    memset(&res, 0, sizeof res);
    res.resType = cudaResourceTypeArray;
    res.res.array.array = out->arr;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&out->tex, &res, &td, NULL));
    return 0;
}

static int free_texture(RefTex *t)
{
    CK(cudaDestroyTextureObject(t->tex));
    CK(cudaFreeArray(t->arr));
    return 0;
}

/* my own probe kernel: raw tex3D at caller-supplied coordinates (x,y,z) -- used to pin the oracle's
 * texture-unit emulation (fixed-point conversion rule) against the hardware. */
__global__ void probe_tex3d(cudaTextureObject_t tex, const float *xyz, long n, float *out)
{
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = tex3D<float>(tex, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}

extern "C" {

int ref_init(const char *cubin_dir)
{
    snprintf(g_dir, sizeof g_dir, "%s", cubin_dir);
    CK(cudaFree(0));
    return 0;
}

/* _bspline_prefilter on a host volume, result written back in place */
int ref_prefilter(float *h_vol, int D, int H, int W)
{
    size_t n = (size_t)D * H * W;
    float *d;
    CK(cudaMalloc(&d, n * 4));
    CK(cudaMemcpy(d, h_vol, n * 4, cudaMemcpyHostToDevice));
    int rc = prefilter_device(d, D, H, W, 0);
    if (rc) return rc;
    CK(cudaMemcpy(h_vol, d, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    return 0;
}

/*
 * affine(volume, m, interpolation, output=...) GPU branch, transforms.py:164-226.
 * fn: 0 linearTex3D, 1 cubicTex3D, 2 cubicTex3DSimple; prefilter: filt_* modes.
 * h_out must be pre-initialised by the caller: zeros reproduce output=None, anything else reproduces a
 * caller-supplied `output` (out-of-bounds voxels keep their contents).
 * iters > 0: additionally times `iters` launches of the transform kernel alone (CUDA events) -> *ms_kernel
 * (mean per launch), and the three prefilter launches -> *ms_prefilter.
 */
int ref_affine(const float *h_vol, int D, int H, int W, const float *h_m16, int fn, int prefilter,
               float *h_out, int iters, float *ms_kernel, float *ms_prefilter)
{
    size_t n = (size_t)D * H * W;
    int rc = load_module(fn);
    if (rc) return rc;
    CUfunction f;
    CKD(cuModuleGetFunction(&f, g_mod[fn], "transform"));
    float *d_vol, *d_out, *d_m;
    unsigned *d_dims;
    CK(cudaMalloc(&d_vol, n * 4));
    CK(cudaMalloc(&d_out, n * 4));
    CK(cudaMalloc(&d_m, 64));
    CK(cudaMalloc(&d_dims, 16));
    CK(cudaMemcpy(d_vol, h_vol, n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    if (prefilter) {
        if (iters > 0 && ms_prefilter) {
            /* time on a scratch copy so the data used for the result is filtered exactly once */
            float *d_tmp;
            CK(cudaMalloc(&d_tmp, n * 4));
            CK(cudaMemcpy(d_tmp, d_vol, n * 4, cudaMemcpyDeviceToDevice));
            rc = prefilter_device(d_tmp, D, H, W, 0); /* warm-up (module load) */
            if (rc) return rc;
            CK(cudaMemcpy(d_tmp, d_vol, n * 4, cudaMemcpyDeviceToDevice));
            CK(cudaEventRecord(e0, 0));
            rc = prefilter_device(d_tmp, D, H, W, 0);
            if (rc) return rc;
            CK(cudaEventRecord(e1, 0));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(ms_prefilter, e0, e1));
            CK(cudaFree(d_tmp));
        }
        rc = prefilter_device(d_vol, D, H, W, 0);
        if (rc) return rc;
    }
    RefTex t;
    rc = make_texture(d_vol, D, H, W, &t);
    if (rc) return rc;
    unsigned dims[4] = {(unsigned)D, (unsigned)H, (unsigned)W, 0}; /* cp.asarray(shape, uint32) read as uint4 */
    CK(cudaMemcpy(d_dims, dims, 12, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_m, h_m16, 64, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_out, h_out, n * 4, cudaMemcpyHostToDevice));
    int dev, sms;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned blocks, threads;
    elementwise_dims(n, sms, &blocks, &threads);
    void *args[4] = {&d_out, &t.tex, &d_m, &d_dims};
    CKD(cuLaunchKernel(f, blocks, 1, 1, threads, 1, 1, 0, 0, args, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, n * 4, cudaMemcpyDeviceToHost));
    if (iters > 0 && ms_kernel) {
        for (int i = 0; i < 3; i++) CKD(cuLaunchKernel(f, blocks, 1, 1, threads, 1, 1, 0, 0, args, 0));
        CK(cudaEventRecord(e0, 0));
        for (int i = 0; i < iters; i++) CKD(cuLaunchKernel(f, blocks, 1, 1, threads, 1, 1, 0, 0, args, 0));
        CK(cudaEventRecord(e1, 0));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        *ms_kernel = ms / iters;
    }
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
    rc = free_texture(&t);
    if (rc) return rc;
    CK(cudaFree(d_vol));
    CK(cudaFree(d_out));
    CK(cudaFree(d_m));
    CK(cudaFree(d_dims));
    return 0;
}

/* raw hardware trilinear fetches through the reference's texture configuration */
int ref_tex3d_sample(const float *h_vol, int D, int H, int W, const float *h_xyz, long n, float *h_out)
{
    size_t nv = (size_t)D * H * W;
    float *d_vol, *d_xyz, *d_out;
    CK(cudaMalloc(&d_vol, nv * 4));
    CK(cudaMalloc(&d_xyz, n * 12));
    CK(cudaMalloc(&d_out, n * 4));
    CK(cudaMemcpy(d_vol, h_vol, nv * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_xyz, h_xyz, n * 12, cudaMemcpyHostToDevice));
    RefTex t;
    int rc = make_texture(d_vol, D, H, W, &t);
    if (rc) return rc;
    probe_tex3d<<<(unsigned)((n + 255) / 256), 256>>>(t.tex, d_xyz, n, d_out);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out, d_out, n * 4, cudaMemcpyDeviceToHost));
    rc = free_texture(&t);
    if (rc) return rc;
    CK(cudaFree(d_vol));
    CK(cudaFree(d_xyz));
    CK(cudaFree(d_out));
    return 0;
}

} /* extern "C" */
