// vt_api.cu -- the extern "C" surface of libvoltools_b200.so (see include/voltools_b200.h).
#include <cuda.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "vt_common.cuh"

int vt_launch_gather(const VtResampleParams &P, int interp, cudaStream_t st);   // vt_resample_gather.cu
int vt_launch_brick(const VtResampleParams &P, int interp, cudaStream_t st);    // vt_resample_brick.cu
int vt_brick_supported(const VtResampleParams &P, int interp);                  // vt_resample_brick.cu
int vt_launch_slice(VtResampleParams &P, int interp, cudaStream_t st);          // vt_resample_slice.cu
int vt_slice_supported(const VtResampleParams &P, int interp);                  // vt_resample_slice.cu
int vt_slice_plan_impl(const VtResampleParams &P, int interp, int sms, int *chunks, int *z_chunk, int *tma, int *box_w,
                       int *box_h, int *shapes, int *pitches);                                    // vt_resample_slice.cu
size_t vt_slice_project_workspace_bytes(int s0, int s1, int s2);                                   // vt_resample_slice.cu
int vt_launch_slice_project(VtResampleParams &P, int interp, float *d_workspace, cudaStream_t st);  // vt_resample_slice.cu
int vt_prefilter_seq(float *d_vol, int d0, int d1, int d2, cudaStream_t st);    // vt_prefilter.cu
int vt_prefilter_win(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row, long long dst_plane,
                     float *d_ws, size_t ws_bytes, cudaStream_t st);  // vt_prefilter_win.cu
int vt_prefilter_xy_range(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row, long long dst_plane,
                          int z0, int z1, cudaStream_t st);  // vt_prefilter_win.cu
int vt_prefilter_z_range(const float *d_src, float *d_dst, int d0, size_t cols, int z0, int z1, int chunks,
                         cudaStream_t st);  // vt_prefilter_win.cu

int vt_prefilter_z_range_z4(const float *d_src, float *d_dst4, int d0, size_t cols, int z0, int z1, int chunks,
                            cudaStream_t st);  // vt_prefilter_win.cu
int vt_z4_axis(const VtResampleParams &P, int interp);                                                     // vt_resample_z4.cu
size_t vt_z4_floats(int s0, int s1, int s2, int axis);                                                     // vt_resample_z4.cu
int vt_pack_z4_impl(const float *d_src, int s0, int s1, int s2, long long row, long long plane, float *d_dst4, int axis,
                    cudaStream_t st);                                                                      // vt_resample_z4.cu
int vt_z4_plan_impl(const VtResampleParams &P, int axis, int interp, int sms, int *chunks, int *m_chunk, int *box_w,
                    int *box_h, int *shapes, int *pitches, float *wavefronts);                             // vt_resample_z4.cu
int vt_launch_z4(const VtResampleParams &P, const float *d_src4, int axis, int interp, cudaStream_t st);   // vt_resample_z4.cu
bool vt_z4_axis_accepts(const VtResampleParams &P, int axis);                                              // vt_resample_z4.cu

struct vt_tex;
int vt_launch_tex(const VtResampleParams &P, const vt_tex *t, int interp, cudaStream_t st);            // vt_resample_tex.cu
int vt_tex_create_impl(int s0, int s1, int s2, int device, vt_tex **out);                              // vt_resample_tex.cu
int vt_tex_upload_impl(vt_tex *t, const float *d_src, long long row, long long plane, cudaStream_t st);  // vt_resample_tex.cu
int vt_tex_destroy_impl(vt_tex *t);                                                                    // vt_resample_tex.cu
struct vt_tex {  // layout shared with vt_resample_tex.cu
    cudaArray_t arr;
    cudaTextureObject_t tex;
    int s0, s1, s2;
    int device;
};

static std::atomic<long long> g_launches{0};
void vt_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------------
// per-kernel device timing
// ---------------------------------------------------------------------------------------------------
#define VT_HOST_MAX_CHUNKS 16

namespace {
struct ProfRec {
    int id;
    cudaEvent_t e0, e1;
};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRec *> g_prof_recs;
double g_prof_ms[VT_K_COUNT];
long long g_prof_n[VT_K_COUNT];
const char *const g_prof_names[VT_K_COUNT] = {
    "prefilter_x", "prefilter_y", "prefilter_z", "prefilter_fused", "gather_linear", "gather_cubic_tex",
    "gather_cubic_simple", "brick_linear", "brick_cubic_tex", "brick_cubic_simple", "slice_linear",
    "slice_cubic_tex", "slice_cubic_simple", "tex_linear", "tex_cubic", "plane_sum", "project_2d", "z4_linear",
    "z4_cubic_tex", "z4_cubic_simple", "pack_z4", "pad_rows"};

void prof_drain_locked()
{
    for (ProfRec *r : g_prof_recs) {
        float ms = 0.0f;
        if (cudaEventSynchronize(r->e1) == cudaSuccess && cudaEventElapsedTime(&ms, r->e0, r->e1) == cudaSuccess) {
            g_prof_ms[r->id] += ms;
            g_prof_n[r->id] += 1;
        }
        cudaEventDestroy(r->e0);
        cudaEventDestroy(r->e1);
        delete r;
    }
    g_prof_recs.clear();
}
}  // namespace

VtProf::VtProf(int id_, cudaStream_t st_) : id(id_), st(st_), rec(nullptr)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec *r = new (std::nothrow) ProfRec();
    if (!r) return;
    r->id = id;
    if (cudaEventCreate(&r->e0) != cudaSuccess || cudaEventCreate(&r->e1) != cudaSuccess) {
        delete r;
        return;
    }
    cudaEventRecord(r->e0, st);
    rec = r;
}

VtProf::~VtProf()
{
    if (!rec) return;
    ProfRec *r = (ProfRec *)rec;
    cudaEventRecord(r->e1, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
}

int vt_encode_tmap_nd(void *tmap, const void *base, int rank, const unsigned long long *dims, const unsigned long long *strides,
                      const unsigned *box)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = nullptr;
    if (rank < 3 || rank > 5) return VT_ERR_INVALID_ARG;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        VT_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) return VT_ERR_UNSUPPORTED;
        fn = (encode_fn)p;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; i++) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estr[i] = 1;
        if (i + 1 < rank) gstr[i] = strides[i];
    }
    const CUresult cr = fn((CUtensorMap *)tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void *)base, gdim, gstr, bx,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return cr == CUDA_SUCCESS ? VT_OK : 2000 + (int)cr;
}

int vt_encode_tmap_3d(void *tmap, const void *base, const unsigned long long dims[3], const unsigned long long strides[2],
                      const unsigned box[3])
{
    return vt_encode_tmap_nd(tmap, base, 3, dims, strides, box);
}

namespace {

// scoped device switch: unlike the reference's switch_to_device (voltools/utils/general.py:84-88) the
// caller's current device is restored on exit.
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    int status = VT_OK;
    explicit DeviceGuard(int device)
    {
        if (device < 0) return;
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) { status = 1000 + (int)e; return; }
        if (prev != device) {
            e = cudaSetDevice(device);
            if (e != cudaSuccess) { status = 1000 + (int)e; return; }
            changed = true;
        }
    }
    ~DeviceGuard()
    {
        if (changed) cudaSetDevice(prev);
    }
};

int fill_params(VtResampleParams &P, const float *d_src, int s0, int s1, int s2, long long src_row, long long src_plane,
                float *d_dst, int o0, int o1, int o2, long long dst_batch_stride, unsigned flags, int z_begin, int z_end)
{
    if (src_row < s2 || src_plane < src_row * s1) return VT_ERR_INVALID_ARG;
    P.src_row = src_row;
    P.src_plane = src_plane;
    if (!d_src || !d_dst) return VT_ERR_INVALID_ARG;
    if (s0 < 1 || s1 < 1 || s2 < 1 || o0 < 1 || o1 < 1 || o2 < 1) return VT_ERR_INVALID_ARG;
    if (z_begin < 0 || z_end > o0 || z_begin > z_end) return VT_ERR_INVALID_ARG;
    // the reference indexes voxels with 32-bit integers (transforms.py:258-264); so do the kernels' planes
    if (src_plane > 0x7fffffffLL || (long long)o1 * o2 > 0x7fffffffLL) return VT_ERR_UNSUPPORTED;
    P.src = d_src;
    P.dst = d_dst;
    P.s0 = s0; P.s1 = s1; P.s2 = s2;
    P.o0 = o0; P.o1 = o1; P.o2 = o2;
    P.dst_batch_stride = dst_batch_stride;
    P.z_begin = z_begin;
    P.z_end = z_end;
    P.flags = flags;
    return VT_OK;
}

void copy_mats(VtResampleParams &P, const float *h_mats, int first, int count)
{
    P.n_mats = count;
    for (int k = 0; k < count; k++)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) P.mats[k].r[r][c] = h_mats[(size_t)(first + k) * 16 + r * 4 + c];
}

int choose_family(const VtResampleParams &P, int interp, unsigned flags)
{
    const unsigned forced = flags & 0xf0u;
    if (forced == VT_KERNEL_GATHER) return 1;
    if (forced == VT_KERNEL_SLICE) return vt_slice_supported(P, interp) ? 3 : -1;
    if (forced == VT_KERNEL_BRICK) return vt_brick_supported(P, interp) ? 2 : -1;
    if (vt_slice_supported(P, interp)) return 3;
    return vt_brick_supported(P, interp) ? 2 : 1;
}

}  // namespace

bool vt_pdl_enabled()
{
    static const bool on = !(getenv("VT_PDL") && atoi(getenv("VT_PDL")) == 0);
    return on;
}

extern "C" {

int vt_abi_version(void) { return VT_ABI_VERSION; }

const char *vt_error_string(int status)
{
    static thread_local char buf[256];
    switch (status) {
        case VT_OK: return "ok";
        case VT_ERR_INVALID_ARG: return "invalid argument";
        case VT_ERR_UNSUPPORTED: return "unsupported shape or configuration";
        case VT_ERR_NO_DEVICE: return "no CUDA device";
        case VT_ERR_ALLOC: return "allocation failed";
    }
    if (status >= 2000) {
        snprintf(buf, sizeof buf, "CUDA driver error %d", status - 2000);
        return buf;
    }
    if (status >= 1000) {
        snprintf(buf, sizeof buf, "CUDA error %d: %s", status - 1000, cudaGetErrorString((cudaError_t)(status - 1000)));
        return buf;
    }
    return "unknown status";
}

int vt_device_count(int *count)
{
    if (!count) return VT_ERR_INVALID_ARG;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        cudaGetLastError();
        return VT_ERR_NO_DEVICE;
    }
    return VT_OK;
}

long long vt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int vt_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain_locked();
    if (on) {
        for (int i = 0; i < VT_K_COUNT; i++) {
            g_prof_ms[i] = 0.0;
            g_prof_n[i] = 0;
        }
    }
    g_prof_on.store(on ? 1 : 0);
    return VT_OK;
}

int vt_profile_kernel_count(void) { return VT_K_COUNT; }

const char *vt_profile_kernel_name(int kernel)
{
    return (kernel >= 0 && kernel < VT_K_COUNT) ? g_prof_names[kernel] : "";
}

int vt_profile_read(int kernel, double *ms_total, long long *launches)
{
    if (kernel < 0 || kernel >= VT_K_COUNT || !ms_total || !launches) return VT_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    prof_drain_locked();
    *ms_total = g_prof_ms[kernel];
    *launches = g_prof_n[kernel];
    return VT_OK;
}

size_t vt_prefilter_workspace_bytes(int d0, int d1, int d2, long long dst_row_stride, long long dst_plane_stride)
{
    if (d0 < 1 || d1 < 1 || d2 < 1 || dst_row_stride < d2 || dst_plane_stride < dst_row_stride * d1) return 0;
    return (size_t)d0 * (size_t)dst_plane_stride * sizeof(float);
}

int vt_prefilter_strided_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row_stride,
                             long long dst_plane_stride, int variant, int device, void *stream)
{
    return vt_prefilter_ws_f32(d_src, d_dst, d0, d1, d2, dst_row_stride, dst_plane_stride, nullptr, 0, variant, device,
                               stream);
}

int vt_prefilter_ws_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row_stride,
                        long long dst_plane_stride, void *d_workspace, size_t workspace_bytes, int variant, int device,
                        void *stream)
{
    if (!d_src || !d_dst || d0 < 1 || d1 < 1 || d2 < 1) return VT_ERR_INVALID_ARG;
    if (variant != 0 && variant != 1) return VT_ERR_INVALID_ARG;
    if (dst_row_stride < d2 || dst_plane_stride < dst_row_stride * d1) return VT_ERR_INVALID_ARG;
    const bool dense = dst_row_stride == d2 && dst_plane_stride == (long long)d1 * d2;
    if (!dense && (variant != 0 || d_src == (const float *)d_dst)) return VT_ERR_UNSUPPORTED;
    DeviceGuard g(device);
    if (g.status) return g.status;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 0 && d_src != d_dst) {
        const int rc = vt_prefilter_win(d_src, d_dst, d0, d1, d2, dst_row_stride, dst_plane_stride, (float *)d_workspace,
                                        workspace_bytes, st);
        if (rc != VT_ERR_UNSUPPORTED || !dense) return rc;  // rows too long for shared memory: sequential kernels
    }
    if (d_src != d_dst)
        VT_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)d0 * d1 * d2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return vt_prefilter_seq(d_dst, d0, d1, d2, st);
}

int vt_prefilter_planes_f32(const float *d_src, float *d_workspace, float *d_dst, int d0, int d1, int d2,
                            long long dst_row_stride, long long dst_plane_stride, int xy_begin, int xy_end, int z_begin,
                            int z_end, int device, void *stream)
{
    if (!d_src || !d_workspace || !d_dst || d0 < 1 || d1 < 1 || d2 < 1) return VT_ERR_INVALID_ARG;
    if (dst_row_stride < d2 || dst_plane_stride < dst_row_stride * d1) return VT_ERR_INVALID_ARG;
    if (d_workspace == d_dst || d_workspace == (float *)d_src || d_dst == (float *)d_src) return VT_ERR_INVALID_ARG;
    if (xy_begin < 0 || xy_end > d0 || xy_begin > xy_end || z_begin < 0 || z_end > d0 || z_begin > z_end)
        return VT_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (g.status) return g.status;
    int rc = vt_prefilter_xy_range(d_src, d_workspace, d0, d1, d2, dst_row_stride, dst_plane_stride, xy_begin, xy_end,
                                   (cudaStream_t)stream);
    if (rc) return rc;
    return vt_prefilter_z_range(d_workspace, d_dst, d0, (size_t)dst_plane_stride, z_begin, z_end, 1, (cudaStream_t)stream);
}

int vt_prefilter_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, int variant, int device, void *stream)
{
    return vt_prefilter_strided_f32(d_src, d_dst, d0, d1, d2, d2, (long long)d1 * d2, variant, device, stream);
}

int vt_affine_plan(int s0, int s1, int s2, int o0, int o1, int o2, const void *d_src, const float *h_mats, int n_mats,
                   int interp, unsigned flags, int *family)
{
    if (!family || !h_mats || n_mats < 1) return VT_ERR_INVALID_ARG;
    VtResampleParams P;
    float dummy;
    int rc = fill_params(P, d_src ? (const float *)d_src : &dummy, s0, s1, s2, s2, (long long)s1 * s2, &dummy, o0, o1, o2, 0,
                         flags, 0, o0);
    if (rc) return rc;
    if (!d_src) P.src = nullptr;  // (only inspected for its alignment: a null pointer counts as aligned)
    copy_mats(P, h_mats, 0, n_mats < VT_MAX_BATCH ? n_mats : VT_MAX_BATCH);
    const int f = choose_family(P, interp, flags);
    if (f < 0) return VT_ERR_UNSUPPORTED;
    *family = f;
    return VT_OK;
}

int vt_slice_plan(const void *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride, int o0,
                  int o1, int o2, const float *h_mats, int n_mats, int interp, unsigned flags, int sms, int *chunks,
                  int *z_chunk, int *tma, int *box_w, int *box_h, int *shapes, int *pitches)
{
    if (!h_mats || n_mats < 1 || n_mats > VT_MAX_BATCH || !chunks || !z_chunk || !tma || !box_w || !box_h || sms < 1)
        return VT_ERR_INVALID_ARG;
    VtResampleParams P;
    float dummy;
    int rc = fill_params(P, d_src ? (const float *)d_src : &dummy, s0, s1, s2, src_row_stride, src_plane_stride, &dummy, o0,
                         o1, o2, 0, flags, 0, o0);
    if (rc) return rc;
    if (!d_src) P.src = nullptr;  // alignment of a null pointer: aligned
    copy_mats(P, h_mats, 0, n_mats);
    return vt_slice_plan_impl(P, interp, sms, chunks, z_chunk, tma, box_w, box_h, shapes, pitches);
}

int vt_affine_strided_f32(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                          float *d_dst, int o0, int o1, int o2, long long dst_batch_stride, const float *h_mats, int n_mats,
                          int interp, unsigned flags, int z_begin, int z_end, int device, void *stream)
{
    if (!h_mats || n_mats < 0) return VT_ERR_INVALID_ARG;
    if (interp != VT_LINEAR && interp != VT_CUBIC_TEX && interp != VT_CUBIC_SIMPLE) return VT_ERR_INVALID_ARG;
    if (n_mats == 0) return VT_OK;
    VtResampleParams P;
    int rc = fill_params(P, d_src, s0, s1, s2, src_row_stride, src_plane_stride, d_dst, o0, o1, o2, dst_batch_stride, flags,
                         z_begin, z_end);
    if (rc) return rc;
    DeviceGuard g(device);
    if (g.status) return g.status;
    cudaStream_t st = (cudaStream_t)stream;
    for (int first = 0; first < n_mats; first += VT_MAX_BATCH) {
        const int count = (n_mats - first) < VT_MAX_BATCH ? (n_mats - first) : VT_MAX_BATCH;
        copy_mats(P, h_mats, first, count);
        P.dst = d_dst + (size_t)first * dst_batch_stride;
        const int family = choose_family(P, interp, flags);
        if (family < 0) return VT_ERR_UNSUPPORTED;
        rc = family == 3 ? vt_launch_slice(P, interp, st)
             : family == 2 ? vt_launch_brick(P, interp, st) : vt_launch_gather(P, interp, st);
        if (rc) return rc;
    }
    return VT_OK;
}

int vt_affine_f32(const float *d_src, int s0, int s1, int s2, float *d_dst, int o0, int o1, int o2,
                  long long dst_batch_stride, const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin,
                  int z_end, int device, void *stream)
{
    return vt_affine_strided_f32(d_src, s0, s1, s2, s2, (long long)s1 * s2, d_dst, o0, o1, o2, dst_batch_stride, h_mats,
                                 n_mats, interp, flags, z_begin, z_end, device, stream);
}

// ---------------------------------------------------------------------------------------------------
// slice4 family: the Z4 layout (vt_resample_z4.cu)
// ---------------------------------------------------------------------------------------------------
size_t vt_z4_bytes(int s0, int s1, int s2, int axis)
{
    if (s0 < 1 || s1 < 1 || s2 < 1 || axis < 0 || axis > 2) return 0;
    return vt_z4_floats(s0, s1, s2, axis) * sizeof(float);
}

int vt_pack_z4_f32(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                   float *d_dst4, int axis, int device, void *stream)
{
    if (!d_src || !d_dst4 || s0 < 1 || s1 < 1 || s2 < 1 || axis < 0 || axis > 2) return VT_ERR_INVALID_ARG;
    if (src_row_stride < s2 || src_plane_stride < src_row_stride * s1 || ((uintptr_t)d_dst4 % 16) != 0) return VT_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (g.status) return g.status;
    return vt_pack_z4_impl(d_src, s0, s1, s2, src_row_stride, src_plane_stride, d_dst4, axis, (cudaStream_t)stream);
}

__global__ void vt_pad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, int w, int row, size_t rows);
static inline unsigned pad_rows_blocks(size_t rows);

int vt_pad_rows_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row_stride, int device, void *stream)
{
    if (!d_src || !d_dst || d_src == d_dst || d0 < 1 || d1 < 1 || d2 < 1 || dst_row_stride < d2) return VT_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (g.status) return g.status;
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_row_stride > 0x7fffffffLL) return VT_ERR_INVALID_ARG;
    const size_t rows = (size_t)d0 * d1;
    {
        VtProf prof(VT_K_PAD_ROWS, st);
        VT_CUDA(vt_launch_pdl(vt_pad_rows_kernel, dim3(pad_rows_blocks(rows)), dim3(64, 4), 0, st, d_src, d_dst, d2,
                              (int)dst_row_stride, rows));
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

int vt_prefilter_z4_f32(const float *d_src, float *d_dst4, int d0, int d1, int d2, void *d_workspace, size_t workspace_bytes,
                        int device, void *stream)
{
    if (!d_src || !d_dst4 || !d_workspace || d0 < 1 || d1 < 1 || d2 < 1) return VT_ERR_INVALID_ARG;
    if (workspace_bytes < (size_t)d0 * d1 * d2 * sizeof(float) || ((uintptr_t)d_dst4 % 16) != 0) return VT_ERR_INVALID_ARG;
    if (d_workspace == (void *)d_src || d_workspace == (void *)d_dst4 || d_src == d_dst4) return VT_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (g.status) return g.status;
    cudaStream_t st = (cudaStream_t)stream;
    const long long plane = (long long)d1 * d2;
    int rc = vt_prefilter_xy_range(d_src, (float *)d_workspace, d0, d1, d2, d2, plane, 0, d0, st);
    if (rc) return rc;
    return vt_prefilter_z_range_z4((const float *)d_workspace, d_dst4, d0, (size_t)plane, 0, d0, 0, st);
}

int vt_z4_axis_of(int s0, int s1, int s2, int o0, int o1, int o2, const float *h_mats, int n_mats, int interp, int *axis)
{
    if (!axis || !h_mats || n_mats < 1) return VT_ERR_INVALID_ARG;
    VtResampleParams P;
    float dummy;
    int rc = fill_params(P, &dummy, s0, s1, s2, s2, (long long)s1 * s2, &dummy, o0, o1, o2, 0, 0, 0, o0);
    if (rc) return rc;
    (void)interp;
    unsigned mask = 7u;  // axes every chunk of VT_MAX_BATCH matrices accepts
    for (int first = 0; first < n_mats && mask; first += VT_MAX_BATCH) {
        copy_mats(P, h_mats, first, (n_mats - first) < VT_MAX_BATCH ? (n_mats - first) : VT_MAX_BATCH);
        for (int a = 0; a < 3; a++)
            if ((mask >> a & 1u) && !vt_z4_axis_accepts(P, a)) mask &= ~(1u << a);
    }
    *axis = (mask & 1u) ? 0 : ((mask & 2u) ? 1 : ((mask & 4u) ? 2 : -1));
    return VT_OK;
}

int vt_affine_z4_f32(const float *d_src4, int axis, int s0, int s1, int s2, float *d_dst, int o0, int o1, int o2,
                     long long dst_batch_stride, const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin,
                     int z_end, int device, void *stream)
{
    if (!h_mats || n_mats < 0 || !d_src4 || axis < 0 || axis > 2) return VT_ERR_INVALID_ARG;
    if (interp != VT_LINEAR && interp != VT_CUBIC_TEX && interp != VT_CUBIC_SIMPLE) return VT_ERR_INVALID_ARG;
    if (n_mats == 0) return VT_OK;
    VtResampleParams P;
    int rc = fill_params(P, d_src4, s0, s1, s2, s2, (long long)s1 * s2, d_dst, o0, o1, o2, dst_batch_stride, flags, z_begin,
                         z_end);
    if (rc) return rc;
    DeviceGuard g(device);
    if (g.status) return g.status;
    for (int first = 0; first < n_mats; first += VT_MAX_BATCH) {
        const int count = (n_mats - first) < VT_MAX_BATCH ? (n_mats - first) : VT_MAX_BATCH;
        copy_mats(P, h_mats, first, count);
        P.dst = d_dst + (size_t)first * dst_batch_stride;
        // the batch must leave `axis` alone (any of the axes it leaves alone will do)
        if (!vt_z4_axis_accepts(P, axis)) return VT_ERR_UNSUPPORTED;
        rc = vt_launch_z4(P, d_src4, axis, interp, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return VT_OK;
}

int vt_z4_plan(int axis, int s0, int s1, int s2, int o0, int o1, int o2, const float *h_mats, int n_mats, int interp, int sms,
               int *chunks, int *m_chunk, int *box_w, int *box_h, int *shapes, int *pitches, float *wavefronts)
{
    if (!h_mats || n_mats < 1 || n_mats > VT_MAX_BATCH || !chunks || !m_chunk || !box_w || !box_h || sms < 1 || axis < 0 ||
        axis > 2)
        return VT_ERR_INVALID_ARG;
    VtResampleParams P;
    float dummy;
    int rc = fill_params(P, &dummy, s0, s1, s2, s2, (long long)s1 * s2, &dummy, o0, o1, o2, 0, 0, 0, o0);
    if (rc) return rc;
    copy_mats(P, h_mats, 0, n_mats);
    if (!vt_z4_axis_accepts(P, axis)) return VT_ERR_UNSUPPORTED;
    return vt_z4_plan_impl(P, axis, interp, sms, chunks, m_chunk, box_w, box_h, shapes, pitches, wavefronts);
}

// ---------------------------------------------------------------------------------------------------
// rotate-and-project
// ---------------------------------------------------------------------------------------------------
size_t vt_project_workspace_bytes(int s0, int s1, int s2)
{
    if (s0 < 1 || s1 < 1 || s2 < 1) return 0;
    return vt_slice_project_workspace_bytes(s0, s1, s2);
}

int vt_project_strided_f32(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                           float *d_proj, int o0, int o1, int o2, long long proj_batch_stride, const float *h_mats,
                           int n_mats, int interp, unsigned flags, int z_begin, int z_end, void *d_workspace,
                           size_t workspace_bytes, int device, void *stream)
{
    if (!h_mats || n_mats < 0) return VT_ERR_INVALID_ARG;
    if (interp != VT_LINEAR && interp != VT_CUBIC_TEX && interp != VT_CUBIC_SIMPLE) return VT_ERR_INVALID_ARG;
    if (proj_batch_stride < (long long)o1 * o2) return VT_ERR_INVALID_ARG;
    if (n_mats == 0) return VT_OK;
    // OOB voxels add nothing to a sum: the OOB policy bits are meaningless here
    flags &= ~(unsigned)VT_OOB_ZERO;
    VtResampleParams P;
    int rc = fill_params(P, d_src, s0, s1, s2, src_row_stride, src_plane_stride, d_proj, o0, o1, o2, proj_batch_stride, flags,
                         z_begin, z_end);
    if (rc) return rc;
    DeviceGuard g(device);
    if (g.status) return g.status;
    cudaStream_t st = (cudaStream_t)stream;
    const bool have_ws = d_workspace && workspace_bytes >= vt_slice_project_workspace_bytes(s0, s1, s2);
    for (int first = 0; first < n_mats; first += VT_MAX_BATCH) {
        const int count = (n_mats - first) < VT_MAX_BATCH ? (n_mats - first) : VT_MAX_BATCH;
        copy_mats(P, h_mats, first, count);
        P.dst = d_proj + (size_t)first * proj_batch_stride;
        P.flags = flags;
        int family = choose_family(P, interp, flags);
        if (family < 0) return VT_ERR_UNSUPPORTED;
        if (family == 3 && have_ws && s1 <= 65535 && o1 <= 65535 * 16) {  // (grid limits of the plane-sum / 2-D kernels)
            rc = vt_launch_slice_project(P, interp, (float *)d_workspace, st);
        } else {
            // general matrices (or no workspace): the resampling kernels add into the zeroed image
            if (family == 3) family = vt_brick_supported(P, interp) ? 2 : 1;
            // one memset per image keeps a padded proj_batch_stride intact (an empty z range leaves the zeros)
            for (int k = 0; k < count; k++)
                VT_CUDA(cudaMemsetAsync(P.dst + (size_t)k * proj_batch_stride, 0, (size_t)o1 * o2 * sizeof(float), st));
            P.flags = flags | VT_INTERNAL_PROJECT;
            rc = family == 2 ? vt_launch_brick(P, interp, st) : vt_launch_gather(P, interp, st);
        }
        if (rc) return rc;
    }
    return VT_OK;
}

int vt_project_tex_f32(const vt_tex *t, float *d_proj, int o0, int o1, int o2, long long proj_batch_stride,
                       const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin, int z_end, void *stream)
{
    if (!t || !h_mats || n_mats < 0) return VT_ERR_INVALID_ARG;
    if (interp != VT_LINEAR && interp != VT_CUBIC_TEX) return VT_ERR_INVALID_ARG;
    if (flags & VT_WEIGHTS_EXACT) return VT_ERR_UNSUPPORTED;
    if (proj_batch_stride < (long long)o1 * o2) return VT_ERR_INVALID_ARG;
    if (n_mats == 0) return VT_OK;
    flags = (flags & ~(unsigned)VT_OOB_ZERO) | VT_INTERNAL_PROJECT;
    VtResampleParams P;
    float dummy;
    int rc = fill_params(P, &dummy, t->s0, t->s1, t->s2, t->s2, (long long)t->s1 * t->s2, d_proj, o0, o1, o2,
                         proj_batch_stride, flags, z_begin, z_end);
    if (rc) return rc;
    DeviceGuard g(t->device);
    if (g.status) return g.status;
    cudaStream_t st = (cudaStream_t)stream;
    for (int first = 0; first < n_mats; first += VT_MAX_BATCH) {
        const int count = (n_mats - first) < VT_MAX_BATCH ? (n_mats - first) : VT_MAX_BATCH;
        copy_mats(P, h_mats, first, count);
        P.dst = d_proj + (size_t)first * proj_batch_stride;
        for (int k = 0; k < count; k++)
            VT_CUDA(cudaMemsetAsync(P.dst + (size_t)k * proj_batch_stride, 0, (size_t)o1 * o2 * sizeof(float), st));
        rc = vt_launch_tex(P, t, interp, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return VT_OK;
}

// ---------------------------------------------------------------------------------------------------
// texture family
// ---------------------------------------------------------------------------------------------------
int vt_tex_create(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                  int device, void *stream, vt_tex **out)
{
    if (!out) return VT_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (g.status) return g.status;
    if (device < 0) VT_CUDA(cudaGetDevice(&device));
    vt_tex *t = nullptr;
    int rc = vt_tex_create_impl(s0, s1, s2, device, &t);
    if (rc) return rc;
    if (d_src) {
        rc = vt_tex_upload_impl(t, d_src, src_row_stride, src_plane_stride, (cudaStream_t)stream);
        if (rc) {
            vt_tex_destroy_impl(t);
            return rc;
        }
    }
    *out = t;
    return VT_OK;
}

int vt_tex_upload(vt_tex *t, const float *d_src, long long src_row_stride, long long src_plane_stride, void *stream)
{
    if (!t) return VT_ERR_INVALID_ARG;
    DeviceGuard g(t->device);
    if (g.status) return g.status;
    return vt_tex_upload_impl(t, d_src, src_row_stride, src_plane_stride, (cudaStream_t)stream);
}

int vt_tex_destroy(vt_tex *t)
{
    if (!t) return VT_OK;
    DeviceGuard g(t->device);
    return vt_tex_destroy_impl(t);
}

int vt_affine_tex_f32(const vt_tex *t, float *d_dst, int o0, int o1, int o2, long long dst_batch_stride,
                      const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin, int z_end, void *stream)
{
    if (!t || !h_mats || n_mats < 0) return VT_ERR_INVALID_ARG;
    if (interp != VT_LINEAR && interp != VT_CUBIC_TEX) return VT_ERR_INVALID_ARG;
    if (flags & VT_WEIGHTS_EXACT) return VT_ERR_UNSUPPORTED;  // the unit's weights are what they are
    if (n_mats == 0) return VT_OK;
    VtResampleParams P;
    float dummy;  // fill_params wants a source pointer; this family reads the texture
    int rc = fill_params(P, &dummy, t->s0, t->s1, t->s2, t->s2, (long long)t->s1 * t->s2, d_dst, o0, o1, o2,
                         dst_batch_stride, flags, z_begin, z_end);
    if (rc) return rc;
    DeviceGuard g(t->device);
    if (g.status) return g.status;
    for (int first = 0; first < n_mats; first += VT_MAX_BATCH) {
        const int count = (n_mats - first) < VT_MAX_BATCH ? (n_mats - first) : VT_MAX_BATCH;
        copy_mats(P, h_mats, first, count);
        P.dst = d_dst + (size_t)first * dst_batch_stride;
        rc = vt_launch_tex(P, t, interp, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return VT_OK;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer path
// ---------------------------------------------------------------------------------------------------
struct vt_host_ctx {
    int device;
    cudaStream_t st_in, st_k, st_out;
    float *d_src, *d_dst, *d_coef, *d_ws;
    size_t cap_src, cap_dst, cap_coef, cap_ws;
    cudaEvent_t ev_k;
    cudaEvent_t ev_in[VT_HOST_MAX_CHUNKS];
};

// Calls in flight on different contexts of one device (transform() from several host threads) take turns on the
// host-to-device link: an upload waits for the previous call's upload (an event chain per device, no host blocking), so
// that it runs next to that call's DOWNLOAD -- PCIe is full duplex -- instead of sharing the upload direction with it
// and leaving the download direction idle.  Measured at 250^3 filt_bspline, 8 volumes from 2 threads: 1.63 ms per
// volume without the chain (1.82 ms from one thread; duplex floor 1.29 ms).
constexpr int VT_MAX_DEVICES = 64;
std::mutex g_upload_mu[VT_MAX_DEVICES];     // one per device: calls on different GPUs do not serialise each other
cudaEvent_t g_upload_done[VT_MAX_DEVICES];  // created on first use, never destroyed (process lifetime)

int vt_host_ctx_create(int device, vt_host_ctx **out)
{
    if (!out) return VT_ERR_INVALID_ARG;
    DeviceGuard g(device);
    if (g.status) return g.status;
    vt_host_ctx *c = new (std::nothrow) vt_host_ctx();
    if (!c) return VT_ERR_ALLOC;
    memset(c, 0, sizeof *c);
    // anything that fails below releases what was created so far
    auto build = [&]() -> int {
        if (device < 0) VT_CUDA(cudaGetDevice(&device));
        c->device = device;
        VT_CUDA(cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
        VT_CUDA(cudaStreamCreateWithFlags(&c->st_k, cudaStreamNonBlocking));
        VT_CUDA(cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
        VT_CUDA(cudaEventCreateWithFlags(&c->ev_k, cudaEventDisableTiming));
        for (int i = 0; i < VT_HOST_MAX_CHUNKS; i++) VT_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        return VT_OK;
    };
    const int rc = build();
    if (rc) {
        if (c->ev_k) cudaEventDestroy(c->ev_k);
        for (int i = 0; i < VT_HOST_MAX_CHUNKS; i++)
            if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->st_in) cudaStreamDestroy(c->st_in);
        if (c->st_k) cudaStreamDestroy(c->st_k);
        if (c->st_out) cudaStreamDestroy(c->st_out);
        delete c;
        return rc;
    }
    *out = c;
    return VT_OK;
}

// releases the device buffers of a context (they grow to the largest volume seen: d_src, d_coef, d_ws, d_dst --
// 16 GiB after one 1024^3 filt_* call); the context stays usable and re-allocates on demand
int vt_host_ctx_trim(vt_host_ctx *c)
{
    if (!c) return VT_OK;
    DeviceGuard g(c->device);
    if (g.status) return g.status;
    cudaStreamSynchronize(c->st_in);
    cudaStreamSynchronize(c->st_k);
    cudaStreamSynchronize(c->st_out);
    cudaFree(c->d_src);
    cudaFree(c->d_dst);
    cudaFree(c->d_coef);
    cudaFree(c->d_ws);
    c->d_src = c->d_dst = c->d_coef = c->d_ws = nullptr;
    c->cap_src = c->cap_dst = c->cap_coef = c->cap_ws = 0;
    return VT_OK;
}

int vt_host_ctx_destroy(vt_host_ctx *c)
{
    if (!c) return VT_OK;
    DeviceGuard g(c->device);
    cudaStreamSynchronize(c->st_in);
    cudaStreamSynchronize(c->st_k);
    cudaStreamSynchronize(c->st_out);
    cudaFree(c->d_src);
    cudaFree(c->d_dst);
    cudaFree(c->d_coef);
    cudaFree(c->d_ws);
    cudaEventDestroy(c->ev_k);
    for (int i = 0; i < VT_HOST_MAX_CHUNKS; i++) cudaEventDestroy(c->ev_in[i]);
    cudaStreamDestroy(c->st_in);
    cudaStreamDestroy(c->st_k);
    cudaStreamDestroy(c->st_out);
    delete c;
    return VT_OK;
}

// dense rows -> rows padded to `row` floats (pad columns zero); blockDim = (64, 4): threadIdx.y picks a row, the 64 lanes
// walk along it (no per-element division: the one-thread-per-element version was instruction bound)
__global__ void __launch_bounds__(256) vt_pad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, int w, int row,
                                                          size_t rows)
{
    vt_pdl_wait();
    for (size_t r = (size_t)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (size_t)gridDim.x * blockDim.y) {
        const float *s = src + r * (size_t)w;
        float *d = dst + r * (size_t)row;
        for (int x0 = threadIdx.x; x0 < row; x0 += 4 * 64) {  // four loads in flight per thread
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = x0 + 64 * k < w ? __ldg(s + x0 + 64 * k) : 0.0f;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (x0 + 64 * k < row) d[x0 + 64 * k] = v[k];
        }
    }
}
static inline unsigned pad_rows_blocks(size_t rows) { return (unsigned)((rows + 3) / 4 < 148 * 128 ? (rows + 3) / 4 : 148 * 128); }

static int ensure(float **p, size_t *cap, size_t bytes)
{
    if (*cap >= bytes) return VT_OK;
    if (*p) VT_CUDA(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    VT_CUDA(cudaMalloc(p, bytes));
    *cap = bytes;
    return VT_OK;
}

// Pipeline of the host-buffer path.  The volume is uploaded in z-chunks on the copy stream; as each chunk lands the
// compute stream runs the XY prefilter of its planes (planes are independent), the Z prefilter of every plane whose
// 12-plane look-ahead is now complete, the resampling of every output plane whose input planes are final, and the
// download stream ships those output planes.  Upload and download overlap (PCIe is full duplex), so a call costs
// about one upload plus a two-chunk tail instead of upload + kernels + download.  The early resampling needs a
// matrix of the slice family (output plane z reads input planes z + t0 - 1 .. z + t0 + 1); for a general matrix
// every output plane can read any input plane, so the resampling waits for the whole volume and only the download
// is overlapped (in z-slabs).
static int host_affine_impl(vt_host_ctx *c, const float *h_src, int s0, int s1, int s2, float *h_dst, int o0, int o1, int o2,
                            const float *h_m16, int interp, int prefilter, unsigned flags, cudaEvent_t *dbg);

int vt_host_affine_f32(vt_host_ctx *c, const float *h_src, int s0, int s1, int s2, float *h_dst, int o0, int o1, int o2,
                       const float *h_m16, int interp, int prefilter, unsigned flags)
{
    if (!c || !h_src || !h_dst || !h_m16) return VT_ERR_INVALID_ARG;
    if (s0 < 1 || s1 < 1 || s2 < 1 || o0 < 1 || o1 < 1 || o2 < 1) return VT_ERR_INVALID_ARG;
    if (interp != VT_LINEAR && interp != VT_CUBIC_TEX && interp != VT_CUBIC_SIMPLE) return VT_ERR_INVALID_ARG;
    DeviceGuard g(c->device);
    if (g.status) return g.status;
    static const bool debug = getenv("VT_HOST_DEBUG") != nullptr;  // prints the timeline of the three streams
    cudaEvent_t dbg[4] = {nullptr, nullptr, nullptr, nullptr};
    if (debug)
        for (auto &e : dbg) cudaEventCreate(&e);
    const int rc = host_affine_impl(c, h_src, s0, s1, s2, h_dst, o0, o1, o2, h_m16, interp, prefilter, flags,
                                    debug ? dbg : nullptr);
    if (rc) {
        // ONE exit for every failure inside the pipeline: copies on the caller's arrays may still be in flight, and the
        // caller is free to release h_src / h_dst as soon as this returns
        cudaStreamSynchronize(c->st_in);
        cudaStreamSynchronize(c->st_k);
        cudaStreamSynchronize(c->st_out);
    }
    for (auto &e : dbg)
        if (e) cudaEventDestroy(e);
    return rc;
}

static int host_affine_impl(vt_host_ctx *c, const float *h_src, int s0, int s1, int s2, float *h_dst, int o0, int o1, int o2,
                            const float *h_m16, int interp, int prefilter, unsigned flags, cudaEvent_t *dbg)
{
    const bool debug = dbg != nullptr;
    const size_t plane_in = (size_t)s1 * s2, plane_out = (size_t)o1 * o2;
    // device copies keep their rows padded to 16 bytes so that the resampling kernels can stage with TMA whatever
    // the width is; the upload itself does the padding (2-D copy), the prefilter writes padded rows directly
    const long long row = ((long long)s2 + 3) / 4 * 4, plane = row * s1;
    int rc = ensure(&c->d_dst, &c->cap_dst, (size_t)o0 * plane_out * 4);
    if (rc) return rc;
    rc = ensure(&c->d_coef, &c->cap_coef, (size_t)plane * s0 * 4);
    if (rc) return rc;
    // an unfiltered volume whose rows are not a multiple of 16 bytes is uploaded densely and padded on the device
    // (a 2-D copy with 1000-byte rows runs at a seventh of the PCIe rate)
    const bool pad = !prefilter && row != s2;
    if (prefilter || pad) {
        rc = ensure(&c->d_src, &c->cap_src, plane_in * s0 * 4);
        if (rc) return rc;
    }
    if (prefilter) {
        rc = ensure(&c->d_ws, &c->cap_ws, (size_t)plane * s0 * 4);
        if (rc) return rc;
    }
    // output=None semantics (transforms.py:207-210): skipped voxels are zero -> fused as VT_OOB_ZERO
    const unsigned fl = (flags & ~1u) | VT_OOB_ZERO;
    // can output planes be produced before the whole input is there?
    VtResampleParams P;
    rc = fill_params(P, c->d_coef, s0, s1, s2, row, plane, c->d_dst, o0, o1, o2, 0, fl, 0, o0);
    if (rc) return rc;
    copy_mats(P, h_m16, 0, 1);
    const bool slice = choose_family(P, interp, fl) == 3;
    const int t0 = slice ? (int)P.mats[0].r[0][3] : 0;
    const int margin = interp == VT_LINEAR ? 0 : 1;
    const int KZ = 12;  // look-ahead of the Z prefilter (vt_prefilter_win.cu)

    // Chunk plan: a handful of large chunks (>= 32 planes; measured on B200 / PCIe 5 x16 at 250^3 filt_bspline: 2.41 ms
    // with one chunk, 1.83 ms with 4-8 equal ones, no better with 16 -- per-chunk event and launch latencies start to
    // show), then a short tail of shrinking chunks: what is still to be downloaded after the last upload has landed is
    // the last chunk plus the prefilter's look-ahead, and nothing overlaps that download.
    int bounds[VT_HOST_MAX_CHUNKS + 1];
    int nch = 0;
    bounds[0] = 0;
    {
        const int tail[3] = {24, 16, 10};
        const int tail_total = s0 >= 128 ? 50 : 0;
        const int body = s0 - tail_total;
        int nbody = body / 40;
        if (nbody > 8) nbody = 8;
        if (nbody < 1) nbody = 1;
        if (const char *e = getenv("VT_HOST_CHUNKS")) nbody = atoi(e) > 0 ? (atoi(e) > 12 ? 12 : atoi(e)) : nbody;  // tuning knob
        for (int i = 1; i <= nbody; i++) bounds[++nch] = (int)((long long)body * i / nbody);
        if (tail_total)
            for (int i = 0; i < 3; i++) { bounds[nch + 1] = bounds[nch] + tail[i]; nch++; }
    }
    if (debug) cudaEventRecord(dbg[0], c->st_in);
    // 1) the whole upload, chunk by chunk, on the copy stream, after the previous call's upload on this device
    static const bool chain = getenv("VT_HOST_NO_CHAIN") == nullptr;  // A/B knob
    std::unique_lock<std::mutex> upload_turn(g_upload_mu[(unsigned)c->device % VT_MAX_DEVICES], std::defer_lock);
    cudaEvent_t *turn = (chain && c->device < VT_MAX_DEVICES) ? &g_upload_done[c->device] : nullptr;
    if (turn) {
        upload_turn.lock();
        if (*turn) VT_CUDA(cudaStreamWaitEvent(c->st_in, *turn, 0));
        else VT_CUDA(cudaEventCreateWithFlags(turn, cudaEventDisableTiming));
    }
    for (int i = 0; i < nch; i++) {
        const int h0 = bounds[i], h1 = bounds[i + 1];
        float *up = (prefilter || pad) ? c->d_src : c->d_coef;  // dense either way
        VT_CUDA(cudaMemcpyAsync(up + (size_t)h0 * plane_in, h_src + (size_t)h0 * plane_in,
                                (size_t)(h1 - h0) * plane_in * 4, cudaMemcpyHostToDevice, c->st_in));
        VT_CUDA(cudaEventRecord(c->ev_in[i], c->st_in));
    }
    if (turn) {
        VT_CUDA(cudaEventRecord(*turn, c->st_in));
        upload_turn.unlock();
    }
    if (debug) cudaEventRecord(dbg[1], c->st_in);
    // 2) kernels behind it on the compute stream, downloads behind those on the download stream
    int z_done = 0, o_done = 0;
    for (int i = 0; i < nch; i++) {
        const int h0 = bounds[i], h1 = bounds[i + 1];
        VT_CUDA(cudaStreamWaitEvent(c->st_k, c->ev_in[i], 0));
        int ready = h1;  // sampled planes [0, ready) are final
        if (prefilter) {
            rc = vt_prefilter_xy_range(c->d_src, c->d_ws, s0, s1, s2, row, plane, h0, h1, c->st_k);
            if (rc == VT_ERR_UNSUPPORTED) return rc;  // (rows too long for the windowed kernels: not on this path)
            if (rc) return rc;
            const int zn = h1 == s0 ? s0 : h1 - KZ;
            if (zn > z_done) {
                rc = vt_prefilter_z_range(c->d_ws, c->d_coef, s0, (size_t)plane, z_done, zn, nch > 1 ? 1 : 0, c->st_k);
                if (rc) return rc;
                z_done = zn;
            }
            ready = z_done;
        } else if (pad) {
            const size_t rows = (size_t)(h1 - h0) * s1;
            vt_pad_rows_kernel<<<pad_rows_blocks(rows), dim3(64, 4), 0, c->st_k>>>(c->d_src + (size_t)h0 * plane_in, c->d_coef + (size_t)h0 * plane,
                                                           s2, (int)row, rows);
            vt_count_launch();
            VT_CUDA(cudaGetLastError());
        }
        int on = o_done;
        if (ready == s0) on = o0;
        else if (slice) {
            const long long lim = (long long)ready - t0 - margin;
            on = (int)(lim < o_done ? o_done : (lim > o0 ? o0 : lim));
        }
        if (on <= o_done) continue;
        // a general matrix produces everything at the end: cut it into slabs so that the download still overlaps
        const int pieces = (!slice && on - o_done >= 8) ? 8 : 1;
        const int base = o_done, span = on - o_done;
        for (int pc = 0; pc < pieces; pc++) {
            const int z0 = base + (int)((long long)span * pc / pieces), z1 = base + (int)((long long)span * (pc + 1) / pieces);
            if (z1 <= z0) continue;
            rc = vt_affine_strided_f32(c->d_coef, s0, s1, s2, row, plane, c->d_dst, o0, o1, o2, 0, h_m16, 1, interp, fl, z0,
                                       z1, -1, c->st_k);
            if (rc) return rc;
            VT_CUDA(cudaEventRecord(c->ev_k, c->st_k));
            VT_CUDA(cudaStreamWaitEvent(c->st_out, c->ev_k, 0));
            VT_CUDA(cudaMemcpyAsync(h_dst + (size_t)z0 * plane_out, c->d_dst + (size_t)z0 * plane_out,
                                    (size_t)(z1 - z0) * plane_out * 4, cudaMemcpyDeviceToHost, c->st_out));
        }
        o_done = on;
    }
    if (debug) {
        cudaEventRecord(dbg[2], c->st_k);
        cudaEventRecord(dbg[3], c->st_out);
    }
    VT_CUDA(cudaStreamSynchronize(c->st_out));
    VT_CUDA(cudaStreamSynchronize(c->st_k));
    if (debug) {
        float t1 = 0, t2 = 0, t3 = 0;
        cudaEventElapsedTime(&t1, dbg[0], dbg[1]);
        cudaEventElapsedTime(&t2, dbg[0], dbg[2]);
        cudaEventElapsedTime(&t3, dbg[0], dbg[3]);
        fprintf(stderr, "vt_host_affine_f32: %d chunks; since the first upload started: uploads done %.3f ms, kernels done "
                        "%.3f ms, downloads done %.3f ms\n", nch, t1, t2, t3);
    }
    return VT_OK;
}

}  // extern "C"
