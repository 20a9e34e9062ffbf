// vt_resample_tex.cu -- "texture" kernel family: general affine matrices through a hardware texture object.
//
// For the two interpolators whose arithmetic IS the texture unit (linear = one trilinear fetch, cubic_tex = Ruijters'
// eight trilinear fetches) and a matrix the slice family cannot take, emulating the unit's 1.8 fixed-point
// coordinates and integer texel weights in software costs ~150 (linear) / ~570 (cubic_tex) instructions per voxel
// (vt_resample_brick.cu, ncu: profiles/r01i_brick_ncu_full.txt, issue-bound).  The unit does the same work in one
// TEX instruction per fetch, with bit-identical results by construction.  So this family keeps a second resident
// copy of the sampled volume as a 3-D CUDA array behind a texture object with the reference's descriptor
// (voltools/transforms.py:184-192: float32 channel, border addressing, linear filter, element read, unnormalised
// coordinates) and launches 3-D tiles of output voxels so that a warp's fetches stay inside a few texture-cache
// lines (the reference maps a warp to 32 consecutive voxels of one output row, transforms.py:258-264).
//
// The handle (vt_tex) owns the CUDA array; it is created once per resident volume (StaticVolume) or per call
// (transform()), and the copy into it costs one 8 B/voxel device pass -- the same pass the reference makes
// (transforms.py:197-199).
//
// Replaces the reference's `transform` kernel (voltools/transforms.py:253-282) + linearTex3D / cubicTex3D
// (voltools/kernels/helper_interpolation.h:3-40).  cubic_simple never comes here: its 64 point fetches are cheaper
// from shared memory (brick family).
#include <new>

#include "vt_common.cuh"

struct vt_tex {
    cudaArray_t arr;
    cudaTextureObject_t tex;
    int s0, s1, s2;
    int device;
};

namespace {

constexpr int TZ = 8, TY = 8, TX = 16;  // output tile of a CTA
constexpr int NT = 256;                 // 16 x 8 columns x 2 z-groups, VPT voxels each along z
constexpr int VPT = TZ * TY * TX / NT;  // 4

// Thread mapping: a warp covers 4 x 8 columns (compact in the input volume too) and each thread marches VPT voxels
// along axis 0.  Measured on B200 at 512^3 (tools/perf_probe.py): 8 x 4, 2 x 16 columns per warp and VPT = 2 / 8
// all land within 4 % of this one (340 Gvox/s linear, 66 Gvox/s cubic_tex: the unit's fetch rate, ~1.2 and ~1.8
// float32 trilinear fetches per clock per SM, is the limit, not the mapping); one voxel per thread is 25 % slower.
// MODE: 0 = out-of-bounds voxels are skipped, 1 = written as zero, 2 = rotate-and-project (voxels are summed along axis 0)
template <int INTERP, int MODE>
__global__ void __launch_bounds__(NT) vt_tex_kernel(const __grid_constant__ VtResampleParams P, cudaTextureObject_t tex)
{
    const int tid = threadIdx.x;
    const int nz = P.z_end - P.z_begin;
    const int nzt = (nz + TZ - 1) / TZ;
    const int mat = blockIdx.z / nzt;
    const int zt = blockIdx.z - mat * nzt;
    const VtMat &M = P.mats[mat];
    const int lane = tid & 31, warp = tid >> 5;
    const int tx = (warp & 1) * 8 + (lane & 7);
    const int ty = ((warp >> 1) & 1) * 4 + (lane >> 3);
    const int tz = warp >> 2;
    const int a1 = blockIdx.y * TY + ty, a2 = blockIdx.x * TX + tx;
    if (a1 >= P.o1 || a2 >= P.o2) return;
    const int a0_0 = P.z_begin + zt * TZ + tz * VPT;
    const float f0 = (float)P.s0, f1 = (float)P.s1, f2 = (float)P.s2;
    const float fa1 = (float)a1, fa2 = (float)a2;
    float *__restrict__ dst = P.dst + (size_t)mat * P.dst_batch_stride + ((size_t)a1 * P.o2 + a2);
    const size_t oplane = (size_t)P.o1 * P.o2;
    constexpr bool project = MODE == 2;  // sum along axis 0 instead of storing
    constexpr bool OOB_ZERO = MODE == 1;
    float acc = 0.0f;
#pragma unroll
    for (int v = 0; v < VPT; v++) {
        const int a0 = a0_0 + v;
        if (a0 >= P.z_end) break;
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
        if (INTERP == VT_LINEAR) {
            r = tex3D<float>(tex, p2, p1, p0);  // helper_interpolation.h:3-6
        } else {
            // cubicTex3D, helper_interpolation.h:8-40: same fetch and combine order
            float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
            vt_ruijters(p2, g0x, g1x, h0x, h1x);
            vt_ruijters(p1, g0y, g1y, h0y, h1y);
            vt_ruijters(p0, g0z, g1z, h0z, h1z);
            float t000 = tex3D<float>(tex, h0x, h0y, h0z), t100 = tex3D<float>(tex, h1x, h0y, h0z);
            t000 = __fmaf_rn(g0x, t000, __fmul_rn(g1x, t100));
            float t010 = tex3D<float>(tex, h0x, h1y, h0z), t110 = tex3D<float>(tex, h1x, h1y, h0z);
            t010 = __fmaf_rn(g0x, t010, __fmul_rn(g1x, t110));
            t000 = __fmaf_rn(g0y, t000, __fmul_rn(g1y, t010));
            float t001 = tex3D<float>(tex, h0x, h0y, h1z), t101 = tex3D<float>(tex, h1x, h0y, h1z);
            t001 = __fmaf_rn(g0x, t001, __fmul_rn(g1x, t101));
            float t011 = tex3D<float>(tex, h0x, h1y, h1z), t111 = tex3D<float>(tex, h1x, h1y, h1z);
            t011 = __fmaf_rn(g0x, t011, __fmul_rn(g1x, t111));
            t001 = __fmaf_rn(g0y, t001, __fmul_rn(g1y, t011));
            r = __fmaf_rn(g0z, t000, __fmul_rn(g1z, t001));
        }
        if (project) acc += r;
        else dst[(size_t)a0 * oplane] = r;
    }
    if (project && acc != 0.0f) atomicAdd(dst, acc);
}

template <int INTERP>
int launch(const VtResampleParams &P, cudaTextureObject_t tex, cudaStream_t st)
{
    const int nz = P.z_end - P.z_begin;
    const int nzt = (nz + TZ - 1) / TZ;
    dim3 grid((P.o2 + TX - 1) / TX, (P.o1 + TY - 1) / TY, nzt * P.n_mats);
    if (grid.y > 65535u || grid.z > 65535u) return VT_ERR_UNSUPPORTED;
    {
        VtProf prof(INTERP == VT_LINEAR ? VT_K_TEX_LINEAR : VT_K_TEX_CUBIC, st);
        if (P.flags & VT_INTERNAL_PROJECT) vt_tex_kernel<INTERP, 2><<<grid, NT, 0, st>>>(P, tex);
        else if (P.flags & VT_OOB_ZERO) vt_tex_kernel<INTERP, 1><<<grid, NT, 0, st>>>(P, tex);
        else vt_tex_kernel<INTERP, 0><<<grid, NT, 0, st>>>(P, tex);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

}  // namespace

int vt_launch_tex(const VtResampleParams &P, const vt_tex *t, int interp, cudaStream_t st)
{
    if (P.z_end <= P.z_begin || P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    if (interp == VT_LINEAR) return launch<VT_LINEAR>(P, t->tex, st);
    if (interp == VT_CUBIC_TEX) return launch<VT_CUBIC_TEX>(P, t->tex, st);
    return VT_ERR_INVALID_ARG;
}

int vt_tex_upload_impl(vt_tex *t, const float *d_src, long long row, long long plane, cudaStream_t st)
{
    if (!t || !d_src || row < t->s2 || plane < row * t->s1 || plane % row != 0) return VT_ERR_INVALID_ARG;
    cudaMemcpy3DParms cp;
    memset(&cp, 0, sizeof cp);
    cp.srcPtr = make_cudaPitchedPtr((void *)d_src, (size_t)row * sizeof(float), (size_t)t->s2, (size_t)(plane / row));
    cp.dstArray = t->arr;
    cp.extent = make_cudaExtent((size_t)t->s2, (size_t)t->s1, (size_t)t->s0);
    cp.kind = cudaMemcpyDeviceToDevice;
    VT_CUDA(cudaMemcpy3DAsync(&cp, st));
    return VT_OK;
}

int vt_tex_create_impl(int s0, int s1, int s2, int device, vt_tex **out)
{
    if (!out || s0 < 1 || s1 < 1 || s2 < 1) return VT_ERR_INVALID_ARG;
    vt_tex *t = new (std::nothrow) vt_tex();
    if (!t) return VT_ERR_ALLOC;
    memset(t, 0, sizeof *t);
    t->s0 = s0; t->s1 = s1; t->s2 = s2; t->device = device;
    const cudaChannelFormatDesc ch = cudaCreateChannelDesc(32, 0, 0, 0, cudaChannelFormatKindFloat);
    cudaError_t e = cudaMalloc3DArray(&t->arr, &ch, make_cudaExtent((size_t)s2, (size_t)s1, (size_t)s0), 0);
    if (e != cudaSuccess) {
        delete t;
        cudaGetLastError();
        return e == cudaErrorMemoryAllocation ? VT_ERR_ALLOC : VT_ERR_UNSUPPORTED;  // extent beyond the 3-D texture limits
    }
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = t->arr;
    cudaTextureDesc td;  // transforms.py:187-191
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    e = cudaCreateTextureObject(&t->tex, &rd, &td, nullptr);
    if (e != cudaSuccess) {
        cudaFreeArray(t->arr);
        delete t;
        return 1000 + (int)e;
    }
    *out = t;
    return VT_OK;
}

int vt_tex_destroy_impl(vt_tex *t)
{
    if (!t) return VT_OK;
    cudaDestroyTextureObject(t->tex);
    cudaFreeArray(t->arr);
    delete t;
    return VT_OK;
}
