// vt_resample_gather.cu -- "gather" kernel family: per-output-voxel inverse-affine coordinate generation
// fused with the trilinear / 8-fetch / 64-tap gathers, reading the source volume directly from global
// memory through L1 (ld.global.nc).  Works for any shape, alignment and matrix; it is the fallback of the
// TMA brick family (vt_resample_brick.cu) and the first-correct path the others are validated against.
//
// Replaces the reference's `transform` kernel (voltools/transforms.py:253-282) and its three device
// interpolators (voltools/kernels/helper_interpolation.h:3-68).  Differences in structure, not in results:
//   * 3-D output tiles (32 x 8 x TZ) instead of a 1-D grid-stride loop: neighbouring threads sample
//     neighbouring input voxels in all three axes, and index decomposition needs no division;
//   * the matrix (and a whole batch of them) lives in kernel parameters (constant bank), not in global memory;
//   * the texture unit is reproduced in software (vt_common.cuh: 1.8 fixed-point coordinates, 8-bit integer
//     texel weights, border = 0), so the source is the
//     plain linear buffer: no CUDA-array copy (transforms.py:197-199) is needed.
#include "vt_common.cuh"

namespace {

constexpr int TX = 32, TY = 8, TZ = 8;

struct SrcView {
    const float *__restrict__ p;
    int s0, s1, s2;
    size_t row, plane;  // element strides
    __device__ __forceinline__ float at(int i0, int i1, int i2) const
    {
        // cudaAddressModeBorder (transforms.py:187-189): texels outside the array read as 0
        if ((unsigned)i0 >= (unsigned)s0 || (unsigned)i1 >= (unsigned)s1 || (unsigned)i2 >= (unsigned)s2) return 0.0f;
        return __ldg(p + (size_t)i0 * plane + (size_t)i1 * row + i2);
    }
};

// tex3D<float>(tex, x, y, z): x along axis 2, y along axis 1, z along axis 0
template <int RULE>
__device__ __forceinline__ float tex3d_emul(const SrcView &s, float x, float y, float z)
{
    int i2, i1, i0;
    float c000, c001, c010, c011, c100, c101, c110, c111;
    int a, b, c;
    float ax, ay, az;
    if (RULE == 0) {
        vt_tex_fix_hw(x, i2, a);
        vt_tex_fix_hw(y, i1, b);
        vt_tex_fix_hw(z, i0, c);
    } else {
        vt_tex_fix<2>(x, i2, ax);
        vt_tex_fix<2>(y, i1, ay);
        vt_tex_fix<2>(z, i0, az);
    }
    const bool interior = i0 >= 0 && i1 >= 0 && i2 >= 0 && i0 + 1 < s.s0 && i1 + 1 < s.s1 && i2 + 1 < s.s2;
    if (interior) {
        const float *q = s.p + (size_t)i0 * s.plane + (size_t)i1 * s.row + i2;
        const size_t sy = s.row, sz = s.plane;
        c000 = __ldg(q);           c001 = __ldg(q + 1);
        c010 = __ldg(q + sy);      c011 = __ldg(q + sy + 1);
        c100 = __ldg(q + sz);      c101 = __ldg(q + sz + 1);
        c110 = __ldg(q + sz + sy); c111 = __ldg(q + sz + sy + 1);
    } else {
        c000 = s.at(i0, i1, i2);         c001 = s.at(i0, i1, i2 + 1);
        c010 = s.at(i0, i1 + 1, i2);     c011 = s.at(i0, i1 + 1, i2 + 1);
        c100 = s.at(i0 + 1, i1, i2);     c101 = s.at(i0 + 1, i1, i2 + 1);
        c110 = s.at(i0 + 1, i1 + 1, i2); c111 = s.at(i0 + 1, i1 + 1, i2 + 1);
    }
    if (RULE == 0) {
        int wn[4], wf[4];
        vt_tex_hw_side(a, b, 256 - c, wn);
        vt_tex_hw_side(a, b, c, wf);
        float r = __fmul_rn(vt_u2f(wn[0]), c000);
        r = __fmaf_rn(vt_u2f(wn[1]), c001, r);
        r = __fmaf_rn(vt_u2f(wn[2]), c010, r);
        r = __fmaf_rn(vt_u2f(wn[3]), c011, r);
        r = __fmaf_rn(vt_u2f(wf[0]), c100, r);
        r = __fmaf_rn(vt_u2f(wf[1]), c101, r);
        r = __fmaf_rn(vt_u2f(wf[2]), c110, r);
        r = __fmaf_rn(vt_u2f(wf[3]), c111, r);
        return __fmul_rn(r, 1.0f / 256.0f);
    }
    const float bx = 1.0f - ax, by = 1.0f - ay, bz = 1.0f - az;
    const float r00 = bx * c000 + ax * c001, r01 = bx * c010 + ax * c011;
    const float r10 = bx * c100 + ax * c101, r11 = bx * c110 + ax * c111;
    const float s0 = by * r00 + ay * r01, s1 = by * r10 + ay * r11;
    return bz * s0 + az * s1;
}

// cubicTex3D, helper_interpolation.h:8-40 (same fetch / combine order)
template <int RULE>
__device__ __forceinline__ float cubic_tex(const SrcView &s, float x, float y, float z)
{
    float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
    vt_ruijters(x, g0x, g1x, h0x, h1x);
    vt_ruijters(y, g0y, g1y, h0y, h1y);
    vt_ruijters(z, g0z, g1z, h0z, h1z);
    float t000 = tex3d_emul<RULE>(s, h0x, h0y, h0z), t100 = tex3d_emul<RULE>(s, h1x, h0y, h0z);
    t000 = __fmaf_rn(g0x, t000, __fmul_rn(g1x, t100));
    float t010 = tex3d_emul<RULE>(s, h0x, h1y, h0z), t110 = tex3d_emul<RULE>(s, h1x, h1y, h0z);
    t010 = __fmaf_rn(g0x, t010, __fmul_rn(g1x, t110));
    t000 = __fmaf_rn(g0y, t000, __fmul_rn(g1y, t010));
    float t001 = tex3d_emul<RULE>(s, h0x, h0y, h1z), t101 = tex3d_emul<RULE>(s, h1x, h0y, h1z);
    t001 = __fmaf_rn(g0x, t001, __fmul_rn(g1x, t101));
    float t011 = tex3d_emul<RULE>(s, h0x, h1y, h1z), t111 = tex3d_emul<RULE>(s, h1x, h1y, h1z);
    t011 = __fmaf_rn(g0x, t011, __fmul_rn(g1x, t111));
    t001 = __fmaf_rn(g0y, t001, __fmul_rn(g1y, t011));
    return __fmaf_rn(g0z, t000, __fmul_rn(g1z, t001));
}

// cubicTex3DSimple, helper_interpolation.h:42-68: 64 point fetches at texel centres (exact texels), the
// reference's weight products and accumulation order, so this path is bit-identical to the reference.
__device__ __forceinline__ float cubic_simple(const SrcView &s, float x, float y, float z)
{
    const float cgx = __fadd_rn(x, -0.5f), cgy = __fadd_rn(y, -0.5f), cgz = __fadd_rn(z, -0.5f);
    const float fx0 = floorf(cgx), fy0 = floorf(cgy), fz0 = floorf(cgz);
    const float fx = __fsub_rn(cgx, fx0), fy = __fsub_rn(cgy, fy0), fz = __fsub_rn(cgz, fz0);
    const int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
    float wx[4], wy[4], wz[4];
    vt_bspline4(fx, wx);
    vt_bspline4(fy, wy);
    vt_bspline4(fz, wz);
    const bool interior = iz >= 1 && iy >= 1 && ix >= 1 && iz + 2 < s.s0 && iy + 2 < s.s1 && ix + 2 < s.s2;
    float result = 0.0f;
    if (interior) {
        const float *q = s.p + (size_t)(iz - 1) * s.plane + (size_t)(iy - 1) * s.row + (ix - 1);
        const size_t sy = s.row, sz = s.plane;
#pragma unroll
        for (int kz = 0; kz < 4; kz++) {
#pragma unroll
            for (int ky = 0; ky < 4; ky++) {
                const float byz = __fmul_rn(wy[ky], wz[kz]);
                const float *r = q + kz * sz + ky * sy;
#pragma unroll
                for (int kx = 0; kx < 4; kx++) result = __fmaf_rn(__fmul_rn(wx[kx], byz), __ldg(r + kx), result);
            }
        }
    } else {
#pragma unroll
        for (int kz = 0; kz < 4; kz++) {
#pragma unroll
            for (int ky = 0; ky < 4; ky++) {
                const float byz = __fmul_rn(wy[ky], wz[kz]);
#pragma unroll
                for (int kx = 0; kx < 4; kx++)
                    result = __fmaf_rn(__fmul_rn(wx[kx], byz), s.at(iz - 1 + kz, iy - 1 + ky, ix - 1 + kx), result);
            }
        }
    }
    return result;
}

// MODE: 0 = out-of-bounds voxels are skipped, 1 = written as zero, 2 = rotate-and-project (voxels are summed along axis 0)
template <int INTERP, int RULE, int MODE>
__global__ void __launch_bounds__(TX *TY) vt_gather_kernel(const __grid_constant__ VtResampleParams P)
{
    const int nz = P.z_end - P.z_begin;
    const int nzt = (nz + TZ - 1) / TZ;
    const int mat = blockIdx.z / nzt;
    const int zt = blockIdx.z - mat * nzt;
    const int a2 = blockIdx.x * TX + threadIdx.x;
    const int a1 = blockIdx.y * TY + threadIdx.y;
    if (a2 >= P.o2 || a1 >= P.o1) return;
    const VtMat &M = P.mats[mat];
    const SrcView s{P.src, P.s0, P.s1, P.s2, (size_t)P.src_row, (size_t)P.src_plane};
    const float f0 = (float)P.s0, f1 = (float)P.s1, f2 = (float)P.s2;
    float *__restrict__ dst = P.dst + (size_t)mat * P.dst_batch_stride;
    const float fa1 = (float)a1, fa2 = (float)a2;
    const int zlo = P.z_begin + zt * TZ;
    const int zhi = min(zlo + TZ, P.z_end);
    constexpr bool project = MODE == 2;  // sum along axis 0 instead of storing
    constexpr bool OOB_ZERO = MODE == 1;
    float acc = 0.0f;
    for (int a0 = zlo; a0 < zhi; a0++) {
        const float fa0 = (float)a0;
        const float p0 = vt_row_finish(M.r[0], vt_row_base(M.r[0], fa0, fa1), fa2);
        const float p1 = vt_row_finish(M.r[1], vt_row_base(M.r[1], fa0, fa1), fa2);
        const float p2 = vt_row_finish(M.r[2], vt_row_base(M.r[2], fa0, fa1), fa2);
        const size_t o = ((size_t)a0 * P.o1 + a1) * P.o2 + a2;
        // transforms.py:276-278
        if (p2 < 0 || p1 < 0 || p0 < 0 || p2 >= f2 || p1 >= f1 || p0 >= f0) {
            if (OOB_ZERO && !project) dst[o] = 0.0f;
            continue;
        }
        float r;
        if (INTERP == VT_LINEAR) r = tex3d_emul<RULE>(s, p2, p1, p0);
        else if (INTERP == VT_CUBIC_TEX) r = cubic_tex<RULE>(s, p2, p1, p0);
        else r = cubic_simple(s, p2, p1, p0);
        if (project) acc += r;
        else dst[o] = r;
    }
    if (project && acc != 0.0f) atomicAdd(dst + (size_t)a1 * P.o2 + a2, acc);
}

template <int INTERP, int RULE>
int launch2(const VtResampleParams &P, cudaStream_t st)
{
    const int nz = P.z_end - P.z_begin;
    const int nzt = (nz + TZ - 1) / TZ;
    dim3 grid((P.o2 + TX - 1) / TX, (P.o1 + TY - 1) / TY, nzt * P.n_mats);
    dim3 block(TX, TY, 1);
    if (grid.y > 65535u || grid.z > 65535u) return VT_ERR_UNSUPPORTED;
    {
        VtProf prof(VT_K_GATHER_LINEAR + INTERP, st);
        if (P.flags & VT_INTERNAL_PROJECT) vt_gather_kernel<INTERP, RULE, 2><<<grid, block, 0, st>>>(P);
        else if (P.flags & VT_OOB_ZERO) vt_gather_kernel<INTERP, RULE, 1><<<grid, block, 0, st>>>(P);
        else vt_gather_kernel<INTERP, RULE, 0><<<grid, block, 0, st>>>(P);
    }
    vt_count_launch();
    VT_CUDA(cudaGetLastError());
    return VT_OK;
}

template <int INTERP>
int launch1(const VtResampleParams &P, cudaStream_t st)
{
    if (INTERP == VT_CUBIC_SIMPLE) return launch2<INTERP, 0>(P, st);  // no texture weights involved
    if (P.flags & VT_WEIGHTS_EXACT) return launch2<INTERP, 2>(P, st);
    return launch2<INTERP, 0>(P, st);
}

}  // namespace

int vt_launch_gather(const VtResampleParams &P, int interp, cudaStream_t st)
{
    if (P.z_end <= P.z_begin || P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    switch (interp) {
        case VT_LINEAR: return launch1<VT_LINEAR>(P, st);
        case VT_CUBIC_TEX: return launch1<VT_CUBIC_TEX>(P, st);
        case VT_CUBIC_SIMPLE: return launch1<VT_CUBIC_SIMPLE>(P, st);
    }
    return VT_ERR_INVALID_ARG;
}
