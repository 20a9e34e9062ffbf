/*
 * voltools_b200.h -- C ABI of libvoltools_b200.so: the B200-native (sm_100a) replacement for the device
 * side of voltools' affine volume-resampling path.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / cupy types), returns an
 * int status (0 = VT_OK, otherwise pass it to vt_error_string) and is stream-ordered and non-blocking
 * unless stated.  Device buffers are owned by the caller; the library allocates device memory only inside
 * the opaque contexts created by vt_host_ctx_create().
 *
 * Conventions (identical to the reference): a volume is a C-contiguous float32 array of numpy shape
 * (d0, d1, d2); a transform is a row-major float32 4x4 mapping OUTPUT index (a0,a1,a2,1) to INPUT index
 * (voltools/utils/matrices.py:111-154); the sample point of output voxel a is M*a, evaluated with the
 * reference kernel's float32 recipe (voltools/transforms.py:264-274).
 *
 * Each function names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef VOLTOOLS_B200_H
#define VOLTOOLS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VT_ABI_VERSION 6

/* status codes: 0 ok; 1..99 library errors; 1000+e = cudaError_t e; 2000+e = CUresult e */
#define VT_OK 0
#define VT_ERR_INVALID_ARG 1
#define VT_ERR_UNSUPPORTED 2
#define VT_ERR_NO_DEVICE 3
#define VT_ERR_ALLOC 4

/* interpolation device function, as selected by voltools/transforms.py:11-17 (_INTERPOLATIONS):
 *   'linear' -> VT_LINEAR; 'bspline','filt_bspline' -> VT_CUBIC_TEX; '*_simple' -> VT_CUBIC_SIMPLE.
 * The filt_* names differ only by running vt_prefilter_f32 on the volume first.                      */
#define VT_LINEAR 0       /* linearTex3D       voltools/kernels/helper_interpolation.h:3-6   */
#define VT_CUBIC_TEX 1    /* cubicTex3D        voltools/kernels/helper_interpolation.h:8-40  */
#define VT_CUBIC_SIMPLE 2 /* cubicTex3DSimple  voltools/kernels/helper_interpolation.h:42-68 */

/* flags for vt_affine_f32 (OR together) */
#define VT_OOB_SKIP 0x0u       /* out-of-bounds output voxels are not written (reference: transforms.py:276-278) */
#define VT_OOB_ZERO 0x1u       /* ...are written as 0: fuses the reference's fill(0)/cp.zeros (transforms.py:208,
                                  volume.py:73) into the kernel when the library owns a fresh output            */
#define VT_WEIGHTS_TEX_HW 0x0u /* texture-unit compatible filter weights for VT_LINEAR / VT_CUBIC_TEX: the B200 unit's
                                  1.8 fixed-point coordinates and 8-bit integer texel weights, reproduced in
                                  software (DESIGN.md "texture unit model") -- the parity policy, default         */
#define VT_WEIGHTS_EXACT 0x4u  /* exact float32 fractions (more accurate than the reference; NOT parity)         */
#define VT_KERNEL_AUTO 0x00u   /* pick the kernel family from shape/alignment/matrix                             */
#define VT_KERNEL_GATHER 0x10u /* force: direct global gathers through L1                                        */
#define VT_KERNEL_BRICK 0x20u  /* force: shared-memory brick cache, general matrices (VT_ERR_UNSUPPORTED if impossible) */
#define VT_KERNEL_SLICE 0x30u  /* force: plane-marching kernels for matrices that leave axis 0 alone (ditto)         */

#define VT_STAGE_CP_ASYNC 0x100u /* slice/brick families: stage with per-element cp.async even where TMA box loads are
                                    possible (16-byte aligned source rows); diagnostic / A-B measurements            */

#define VT_MAX_BATCH 32 /* matrices per launch held in kernel parameters; larger batches are chunked */

int vt_abi_version(void);
const char *vt_error_string(int status);

/* replaces voltools/utils/general.py:61-80 (get_available_devices): number of CUDA devices */
int vt_device_count(int *count);

/*
 * Cubic B-spline prefilter: samples d_src -> interpolation coefficients d_dst, X (fastest axis) then Y then Z.
 * Replaces _bspline_prefilter (voltools/transforms.py:290-309) and the kernels SamplesToCoefficients3DX/Y/Z
 * (voltools/kernels/bspline.h:58-99); any shape >= 1 per axis (no power-of-two launch constraint).
 *   d_src, d_dst  device pointers, (d0,d1,d2) float32 C-contiguous.  d_dst == d_src filters in place (what the
 *           reference does); a distinct d_dst leaves the samples untouched, saves the caller a copy and enables
 *           the fast windowed kernels.
 *   variant 0 = default: windowed kernels when d_dst != d_src (16 B/voxel of traffic, coefficients within ~3e-7
 *           of the range of variant 1), otherwise variant 1;
 *           1 = sequential two-sweep kernels with the reference's exact operation order (bit-identical
 *           coefficients), in place or out of place
 *   device  CUDA ordinal, or -1 for the current device
 *   stream  cudaStream_t (NULL = legacy default stream)
 */
int vt_prefilter_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, int variant, int device, void *stream);

/*
 * Same, writing the coefficients with padded strides (in elements): row y of plane z starts at
 * d_dst + z*dst_plane_stride + y*dst_row_stride; the pad columns d2..dst_row_stride-1 are written as zeros.
 * A row stride that is a multiple of 4 floats (16 bytes) lets vt_affine_strided_f32 stage the volume with TMA
 * whatever d2 is (e.g. 250 -> 252).  Only variant 0 with d_dst != d_src supports padded strides.
 * The reference has no counterpart: its resident copy is an opaque CUDA array (voltools/transforms.py:185-199).
 */
int vt_prefilter_strided_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row_stride,
                             long long dst_plane_stride, int variant, int device, void *stream);

/*
 * The unfiltered counterpart: a dense (d0,d1,d2) volume copied into rows of dst_row_stride elements (pad columns zero),
 * plane stride d1*dst_row_stride -- the layout vt_affine_strided_f32 can stage with TMA for any d2.  One 8 B/voxel pass;
 * the reference pays the same pass for its CUDA-array copy (voltools/transforms.py:197-199, voltools/volume.py:46-48).
 */
int vt_pad_rows_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row_stride, int device, void *stream);

/*
 * Same with a caller-owned device workspace of vt_prefilter_workspace_bytes(...) bytes (distinct from d_src and
 * d_dst; contents undefined afterwards).  With it variant 0 runs its Z sweep out of place and in z-chunks, which is
 * what keeps small volumes (<= 256^3: too few columns to cover the HBM latency with one thread per column) at
 * speed.  d_workspace == NULL or too small: identical to vt_prefilter_strided_f32.  Same coefficients either way
 * to ~1e-7 of the range.  The library itself never allocates device memory on this path.
 */
size_t vt_prefilter_workspace_bytes(int d0, int d1, int d2, long long dst_row_stride, long long dst_plane_stride);
int vt_prefilter_ws_f32(const float *d_src, float *d_dst, int d0, int d1, int d2, long long dst_row_stride,
                        long long dst_plane_stride, void *d_workspace, size_t workspace_bytes, int variant, int device,
                        void *stream);

/*
 * The `transform` kernel launch (voltools/transforms.py:253-282 launched at :212 and volume.py:78), for a
 * batch of matrices over one resident source volume.
 *   d_src              sampled volume (s0,s1,s2): raw samples, or coefficients from vt_prefilter_f32
 *   d_dst              n_mats output volumes of shape (o0,o1,o2); matrix k writes at d_dst + k*dst_batch_stride
 *                      (elements).  The reference always has (o0,o1,o2) == (s0,s1,s2).
 *   h_mats             HOST pointer, n_mats row-major float32 4x4 matrices; copied into kernel parameters
 *                      (replaces the per-call cp.asarray(transform_m) H2D copy, volume.py:70)
 *   interp             VT_LINEAR / VT_CUBIC_TEX / VT_CUBIC_SIMPLE
 *   flags              VT_OOB_* | VT_WEIGHTS_* | VT_KERNEL_*
 *   z_begin, z_end     only output planes a0 in [z_begin, z_end) are produced (z-slab sharding); pass 0, o0
 */
int vt_affine_f32(const float *d_src, int s0, int s1, int s2, float *d_dst, int o0, int o1, int o2,
                  long long dst_batch_stride, const float *h_mats, int n_mats, int interp, unsigned flags,
                  int z_begin, int z_end, int device, void *stream);

/* vt_affine_f32 on a source volume with padded strides (elements), as written by vt_prefilter_strided_f32 */
int vt_affine_strided_f32(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                          float *d_dst, int o0, int o1, int o2, long long dst_batch_stride, const float *h_mats, int n_mats,
                          int interp, unsigned flags, int z_begin, int z_end, int device, void *stream);

/*
 * Streaming form of the windowed prefilter, for callers that pipeline it with an upload or a broadcast (the library's
 * own host path, the multi-GPU layer): X and Y passes of sample planes [xy_begin, xy_end) of d_src into the workspace
 * (planes are independent), then the Z pass producing coefficient planes [z_begin, z_end) of d_dst from the workspace.
 * The Z pass reads workspace planes [z_begin - 12, z_end + 12) (clipped to the volume), so those must have been
 * produced by this or an earlier call.  Workspace and d_dst use the same padded strides; all
 * three buffers are distinct.  Covering [0, d0) with consecutive calls gives the coefficients of
 * vt_prefilter_strided_f32 (variant 0) to ~1e-7 of their range.
 */
int vt_prefilter_planes_f32(const float *d_src, float *d_workspace, float *d_dst, int d0, int d1, int d2,
                            long long dst_row_stride, long long dst_plane_stride, int xy_begin, int xy_end, int z_begin,
                            int z_end, int device, void *stream);

/*
 * Texture family: the sampled volume as a 3-D CUDA array behind a texture object with the reference's descriptor
 * (voltools/transforms.py:184-199, voltools/volume.py:37-50: float32 channel, border addressing, linear filter,
 * unnormalised coordinates).  For VT_LINEAR and VT_CUBIC_TEX under a GENERAL matrix the hardware unit beats its
 * software emulation (results are the unit's own, i.e. the reference's, bit for bit); matrices of the slice family
 * and VT_CUBIC_SIMPLE stay on vt_affine_f32.  The handle owns the array (the one place besides vt_host_ctx where the
 * library allocates device memory); vt_tex_create(d_src != NULL) also uploads (one 8 B/voxel device pass, stream
 * ordered), vt_tex_upload re-uploads a volume of the same shape.
 */
typedef struct vt_tex vt_tex;
int vt_tex_create(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                  int device, void *stream, vt_tex **tex);
int vt_tex_upload(vt_tex *tex, const float *d_src, long long src_row_stride, long long src_plane_stride, void *stream);
int vt_tex_destroy(vt_tex *tex);
/* the `transform` kernel launch (voltools/transforms.py:212, volume.py:78) on a texture handle; arguments as
 * vt_affine_f32 (VT_WEIGHTS_EXACT and VT_CUBIC_SIMPLE are not available here) */
int vt_affine_tex_f32(const vt_tex *tex, float *d_dst, int o0, int o1, int o2, long long dst_batch_stride,
                      const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin, int z_end, void *stream);

/*
 * Slice4 family: the `transform` kernel launch (voltools/transforms.py:212, volume.py:78) for matrices that leave one
 * axis m alone -- M[m] = e_m + integer shift, M[r][m] = 0 otherwise: every rotation about axis m through the centre --
 * on a sampled volume stored in the "Z4" layout of that axis,
 *        L[g][y][x][j] = V[index 4g+j along axis m][y][x],   j = 0..3,  zero past the end of axis m,
 * (y, x) = the two other axes in ascending order; vt_z4_bytes() bytes, 16-byte aligned.  The layout takes the place of
 * the reference's CUDA array (voltools/transforms.py:184-199): one TMA box delivers four consecutive planes per texel
 * and every tap is a 16-byte shared-memory load serving four output planes.
 *   vt_pack_z4_f32       plain (possibly row-padded) volume -> Z4 layout of `axis` (one 8 B/voxel pass; the reference
 *                        pays the same pass for its array copy, transforms.py:197-199)
 *   vt_prefilter_z4_f32  vt_prefilter_f32 (variant 0) whose Z sweep writes the Z4 layout of axis 0 directly; the
 *                        caller-owned workspace holds d0*d1*d2 floats
 *   vt_z4_axis_of        *axis = the lowest axis every matrix of the batch leaves alone in the required way, or -1
 *   vt_affine_z4_f32     arguments as vt_affine_f32; VT_ERR_UNSUPPORTED if a matrix does not leave `axis` alone.
 *                        Results are bit-identical to vt_affine_f32's slice kernels.
 *   vt_z4_plan           host-only introspection (no device is touched): chunks along the march axis, TMA box, and per
 *                        matrix the quarter-warp shape (0..3 = 1x8, 2x4, 4x2, 8x1 columns per 8 lanes), the
 *                        shared-memory pitch in texels and the simulated wavefronts per quarter-warp load
 */
size_t vt_z4_bytes(int s0, int s1, int s2, int axis);
int vt_pack_z4_f32(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                   float *d_dst4, int axis, int device, void *stream);
int vt_prefilter_z4_f32(const float *d_src, float *d_dst4, int d0, int d1, int d2, void *d_workspace, size_t workspace_bytes,
                        int device, void *stream);
int vt_z4_axis_of(int s0, int s1, int s2, int o0, int o1, int o2, const float *h_mats, int n_mats, int interp, int *axis);
int vt_affine_z4_f32(const float *d_src4, int axis, int s0, int s1, int s2, float *d_dst, int o0, int o1, int o2,
                     long long dst_batch_stride, const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin,
                     int z_end, int device, void *stream);
int vt_z4_plan(int axis, int s0, int s1, int s2, int o0, int o1, int o2, const float *h_mats, int n_mats, int interp, int sms,
               int *chunks, int *m_chunk, int *box_w, int *box_h, int *shapes, int *pitches, float *wavefronts);

/*
 * Rotate-and-project: the transformed volume summed over axis 0, without ever writing the volume -- what
 * examples/projections.py:20-26 computes as `static_volume.transform(rotation=...).sum(axis=0)` (a transform kernel launch,
 * an N*4 B store, and a reduction that reads it back).  d_proj receives n_mats images of shape (o1, o2), image k at
 * d_proj + k*proj_batch_stride (elements), fully overwritten:
 *        proj_k[a1][a2] = sum over a0 in [z_begin, z_end) of transform_k(src)[a0][a1][a2],   out-of-bounds voxels adding 0
 * (so z-slabs computed on different GPUs add up to the full projection).  Other arguments as vt_affine_strided_f32; the
 * VT_OOB_* bits are ignored.
 *   - matrices that leave axis 0 alone (the example's `rotation=(i,0,0), 'sxyz'`; every tilt about axis 0): by linearity the
 *     in-plane interpolation is applied once to sums of input planes -- one 4 B/voxel read of the source and a 2-D resample.
 *     Needs the caller-owned workspace of vt_project_workspace_bytes(s0,s1,s2) bytes; deterministic.
 *   - any other matrix (or no workspace): the brick / gather kernels accumulate each thread's run along axis 0 in a
 *     register and add it to the image with red.global.add.f32 (summation order, hence the last bits, vary run to run).
 * Float32 summation order differs from transform-then-sum by ~1e-7 of the projection's range either way.
 * vt_project_tex_f32: the same on a texture handle (general matrices, VT_LINEAR / VT_CUBIC_TEX).
 */
size_t vt_project_workspace_bytes(int s0, int s1, int s2);
int vt_project_strided_f32(const float *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride,
                           float *d_proj, int o0, int o1, int o2, long long proj_batch_stride, const float *h_mats,
                           int n_mats, int interp, unsigned flags, int z_begin, int z_end, void *d_workspace,
                           size_t workspace_bytes, int device, void *stream);
int vt_project_tex_f32(const vt_tex *tex, float *d_proj, int o0, int o1, int o2, long long proj_batch_stride,
                       const float *h_mats, int n_mats, int interp, unsigned flags, int z_begin, int z_end, void *stream);

/* which kernel family vt_affine_f32 would run for these arguments: 1 = gather, 2 = brick, 3 = slice */
int vt_affine_plan(int s0, int s1, int s2, int o0, int o1, int o2, const void *d_src, const float *h_mats,
                   int n_mats, int interp, unsigned flags, int *family);

/*
 * Host-only introspection (no device is touched): what the slice family would decide for this launch on a GPU with
 * `sms` multiprocessors -- z-chunks per column tile and their length, TMA or cp.async staging, the TMA box, and per
 * matrix the warp shape (0 = 2x16, 1 = 4x8, 2 = 8x4 lanes over the CTA's 16x16 columns) and shared-memory pitch.
 * VT_ERR_UNSUPPORTED if the matrices are not of the slice family.  `d_src` is only inspected for its alignment.
 * Used by the CPU-side tests to pin the tuned heuristics.
 */
int vt_slice_plan(const void *d_src, int s0, int s1, int s2, long long src_row_stride, long long src_plane_stride, int o0,
                  int o1, int o2, const float *h_mats, int n_mats, int interp, unsigned flags, int sms, int *chunks,
                  int *z_chunk, int *tma, int *box_w, int *box_h, int *shapes, int *pitches);

/*
 * Host-buffer path: numpy in -> numpy out, as transforms.affine() with output=None
 * (voltools/transforms.py:180-223: H2D, [prefilter], kernel, D2H).  Blocking.  The context owns three streams and
 * device buffers that grow to the largest volume seen (vt_host_ctx_trim releases them; the context stays usable), and
 * pipelines the copies with the kernels in z-chunks.  It copies straight from / to the caller's arrays: the copies only
 * overlap the kernels when h_src and h_dst are page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory);
 * with pageable arrays the CUDA runtime stages every copy through its own bounce buffer and blocks the calling thread.
 * The Python layer therefore stages pageable numpy inputs through a pinned buffer and returns pinned-backed results.
 * On failure every copy that touches the caller's arrays has completed when the call returns.
 */
typedef struct vt_host_ctx vt_host_ctx;
int vt_host_ctx_create(int device, vt_host_ctx **ctx);
int vt_host_ctx_destroy(vt_host_ctx *ctx);
int vt_host_ctx_trim(vt_host_ctx *ctx);
int vt_host_affine_f32(vt_host_ctx *ctx, const float *h_src, int s0, int s1, int s2, float *h_dst, int o0, int o1,
                       int o2, const float *h_m16, int interp, int prefilter, unsigned flags);

/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
long long vt_launch_count(void);

/*
 * Per-kernel device timing, for roofline reporting (replaces the reference's `profile=` event pair around
 * the whole call, voltools/transforms.py:167-169, :214-219, at kernel granularity).  While enabled every
 * kernel launch of the library is bracketed by CUDA events on its own stream.  vt_profile_read blocks until
 * the recorded launches of that kernel have finished and returns their summed duration and their count
 * since the last vt_profile_enable(1).
 */
int vt_profile_enable(int on);
int vt_profile_kernel_count(void);
const char *vt_profile_kernel_name(int kernel);
int vt_profile_read(int kernel, double *ms_total, long long *launches);

#ifdef __cplusplus
}
#endif
#endif /* VOLTOOLS_B200_H */
