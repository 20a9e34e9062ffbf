// vt_resample_slice.cu -- "slice" kernel family: transforms that leave axis 0 alone.
//
// When the matrix maps output axis 0 straight onto input axis 0 with an integer offset
//        M[0] = (1, 0, 0, t0),  t0 integer,  M[1][0] = M[2][0] = 0
// (every rotation about axis 0 through the centre of the volume: BASELINE configs[0..2], the reference's
// README sweep `rotation=(0, i, 0)` with 'rzxz'), the sample point of output voxel (a0, a1, a2) is
// (a0 + t0, q1(a1,a2), q2(a1,a2)): the in-plane position and therefore every in-plane interpolation weight is
// the same for the whole column a0 = 0..o0-1, and the fraction along axis 0 is exactly 0.  The kernel exploits
// that (the general kernels cannot):
//   * a CTA owns a 16x16 tile of (a1, a2) columns and marches along axis 0;
//   * the in-plane footprint of the tile (its rotated bounding box + filter support) is the same rectangle in
//     every input plane; it is staged plane by plane into a shared-memory ring with cp.async (zero-filled
//     outside the source = the texture's border mode), using per-thread offsets computed once;
//   * weights are computed once per column with the reference's exact float32 recipe and kept in registers;
//   * per plane a thread does 4 (linear) or 16 (cubic) shared-memory loads; the three axis-0 taps of the cubic
//     modes come from a register sliding window over the per-plane sums.
// Results: same arithmetic as the gather family up to float32 summation order (<= ~1e-7 of the range).
//
// Replaces the reference's `transform` kernel (voltools/transforms.py:253-282) + linearTex3D / cubicTex3D /
// cubicTex3DSimple (voltools/kernels/helper_interpolation.h:3-68) for this class of matrices.
#include "vt_common.cuh"

namespace {

constexpr int TS = 16;            // tile edge in (a1, a2)
constexpr int NT = TS * TS;       // threads per CTA, one column each
constexpr int BMAX = 32;          // max footprint edge (texels)
constexpr int PITCH = BMAX + 1;   // shared-memory row pitch (odd: spreads rows over banks)
constexpr int STAGE = BMAX * PITCH;
constexpr int NSTAGE = 3;
constexpr int EPT = (BMAX * BMAX + NT - 1) / NT;  // footprint elements per thread (4)

__device__ __forceinline__ void cp_async4_zfill(float *smem_dst, const float *gmem_src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// in-plane coordinate of a column, reference recipe (transforms.py:264-274) with the a0 term dropped: it is an
// exact no-op because M[r][0] == 0 (fma(a0, 0, t) == t).
__device__ __forceinline__ float inplane_coord(const float *row, float a1, float a2)
{
    float t = __fmul_rn(a1, row[1]);
    t = __fmaf_rn(a2, row[2], t);
    t = __fadd_rn(row[3], t);
    return __fadd_rn(t, 0.5f);
}

template <int INTERP>
struct Taps;

// ---- linear: one plane, 4 taps with the texture unit's integer weights (c = 0 -> S = 256) ----------------
template <>
struct Taps<VT_LINEAR> {
    static constexpr int LO = 0, HI = 2;  // footprint margins relative to floor(p - 0.5)
    static constexpr int PLANES_BEFORE = 0, PLANES_AFTER = 0;
    float w[4];
    int off;
    template <int RULE>
    __device__ void init(float p1, float p2, int ylo, int xlo)
    {
        int by, bx;
        if (RULE == 0) {
            int a, b;
            vt_tex_fix_hw(p2, bx, a);
            vt_tex_fix_hw(p1, by, b);
            int wi[4];
            vt_tex_hw_side(a, b, 256, wi);
#pragma unroll
            for (int k = 0; k < 4; k++) w[k] = vt_u2f(wi[k]) * (1.0f / 256.0f);
        } else {
            float ax, ay;
            vt_tex_fix<2>(p2, bx, ax);
            vt_tex_fix<2>(p1, by, ay);
            w[0] = (1.0f - ax) * (1.0f - ay);
            w[1] = ax * (1.0f - ay);
            w[2] = (1.0f - ax) * ay;
            w[3] = ax * ay;
        }
        off = (by - ylo) * PITCH + (bx - xlo);
    }
    __device__ __forceinline__ float plane(const float *s) const
    {
        float r = w[0] * s[off];
        r = fmaf(w[1], s[off + 1], r);
        r = fmaf(w[2], s[off + PITCH], r);
        r = fmaf(w[3], s[off + PITCH + 1], r);
        return r;
    }
};

// ---- cubic_simple: 16 in-plane taps, float32 B-spline weights; axis-0 weights B(-1), B(0), B(1) ------------
template <>
struct Taps<VT_CUBIC_SIMPLE> {
    static constexpr int LO = -1, HI = 2;
    static constexpr int PLANES_BEFORE = 1, PLANES_AFTER = 1;
    float w[16];
    int off;
    float wz0, wz1, wz2;
    template <int RULE>
    __device__ void init(float p1, float p2, int ylo, int xlo)
    {
        const float cgx = __fadd_rn(p2, -0.5f), cgy = __fadd_rn(p1, -0.5f);
        const float fx0 = floorf(cgx), fy0 = floorf(cgy);
        const float fx = __fsub_rn(cgx, fx0), fy = __fsub_rn(cgy, fy0);
        float wx[4], wy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            wx[k] = vt_bspline(__fsub_rn((float)(k - 1), fx));
            wy[k] = vt_bspline(__fsub_rn((float)(k - 1), fy));
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) w[j * 4 + i] = __fmul_rn(wx[i], wy[j]);
        off = ((int)fy0 - 1 - ylo) * PITCH + ((int)fx0 - 1 - xlo);
        wz0 = vt_bspline(-1.0f);  // fraction along axis 0 is exactly 0
        wz1 = vt_bspline(0.0f);
        wz2 = vt_bspline(1.0f);
    }
    __device__ __forceinline__ float plane(const float *s) const
    {
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) r = fmaf(w[j * 4 + i], s[off + j * PITCH + i], r);
        return r;
    }
};

// ---- cubic_tex: Ruijters' 8 trilinear fetches with the texture unit's integer weights ---------------------
// With fraction 0 along axis 0: h0z = idx + 0.3 -> texels idx-1 (S = 256-c0) and idx (S = c0), combined with g0z;
// h1z = idx + 1.5 -> texel idx+1 (S = 256-c1 = 256), combined with g1z (texel idx+2 has S = c1 = 0).
// Each of the three planes gets its own set of 16 pre-multiplied tap weights.
template <>
struct Taps<VT_CUBIC_TEX> {
    static constexpr int LO = -1, HI = 2;
    static constexpr int PLANES_BEFORE = 1, PLANES_AFTER = 1;
    float wa[16], wb[16], wc[16];
    int offy[4], offx[4];  // footprint row offsets (already * PITCH) and column offsets of the 4x4 taps
    template <int RULE>
    __device__ void init(float p1, float p2, int ylo, int xlo)
    {
        float g0x, g1x, h0x, h1x, g0y, g1y, h0y, h1y, g0z, g1z, h0z, h1z;
        vt_ruijters(p2, g0x, g1x, h0x, h1x);
        vt_ruijters(p1, g0y, g1y, h0y, h1y);
        vt_ruijters(8.5f, g0z, g1z, h0z, h1z);  // any texel centre: the fraction along axis 0 is exactly 0
        const float gx[2] = {g0x, g1x}, gy[2] = {g0y, g1y};
        const float hx[2] = {h0x, h1x}, hy[2] = {h0y, h1y};
        if (RULE == 0) {
            int bz0, c0, bz1, c1;
            vt_tex_fix_hw(h0z, bz0, c0);  // (7, 205)
            vt_tex_fix_hw(h1z, bz1, c1);  // (9, 0)
            const int S[3] = {256 - c0, c0, 256 - c1};
            const float gz[3] = {g0z, g0z, g1z};
            int ax[2], bx[2], ay[2], by[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                vt_tex_fix_hw(hx[k], bx[k], ax[k]);
                vt_tex_fix_hw(hy[k], by[k], ay[k]);
                offx[2 * k] = bx[k] - xlo;
                offx[2 * k + 1] = bx[k] + 1 - xlo;
                offy[2 * k] = (by[k] - ylo) * PITCH;
                offy[2 * k + 1] = (by[k] + 1 - ylo) * PITCH;
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                float *w = p == 0 ? wa : (p == 1 ? wb : wc);
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        int wi[4];
                        vt_tex_hw_side(ax[i], ay[j], S[p], wi);
                        const float g = __fmul_rn(__fmul_rn(gx[i], gy[j]), gz[p]) * (1.0f / 256.0f);
                        w[(2 * j) * 4 + 2 * i] = g * vt_u2f(wi[0]);
                        w[(2 * j) * 4 + 2 * i + 1] = g * vt_u2f(wi[1]);
                        w[(2 * j + 1) * 4 + 2 * i] = g * vt_u2f(wi[2]);
                        w[(2 * j + 1) * 4 + 2 * i + 1] = g * vt_u2f(wi[3]);
                    }
            }
        } else {
            // exact float32 weights: alphas are the exact fractions of h - 0.5
            int bz0, bz1;
            float c0, c1;
            vt_tex_fix<2>(h0z, bz0, c0);
            vt_tex_fix<2>(h1z, bz1, c1);
            const float S[3] = {1.0f - c0, c0, 1.0f - c1};
            const float gz[3] = {g0z, g0z, g1z};
            int bx[2], by[2];
            float ax[2], ay[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                vt_tex_fix<2>(hx[k], bx[k], ax[k]);
                vt_tex_fix<2>(hy[k], by[k], ay[k]);
                offx[2 * k] = bx[k] - xlo;
                offx[2 * k + 1] = bx[k] + 1 - xlo;
                offy[2 * k] = (by[k] - ylo) * PITCH;
                offy[2 * k + 1] = (by[k] + 1 - ylo) * PITCH;
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                float *w = p == 0 ? wa : (p == 1 ? wb : wc);
#pragma unroll
                for (int j = 0; j < 2; j++)
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const float g = gx[i] * gy[j] * gz[p] * S[p];
                        w[(2 * j) * 4 + 2 * i] = g * (1.0f - ax[i]) * (1.0f - ay[j]);
                        w[(2 * j) * 4 + 2 * i + 1] = g * ax[i] * (1.0f - ay[j]);
                        w[(2 * j + 1) * 4 + 2 * i] = g * (1.0f - ax[i]) * ay[j];
                        w[(2 * j + 1) * 4 + 2 * i + 1] = g * ax[i] * ay[j];
                    }
            }
        }
    }
    __device__ __forceinline__ void plane3(const float *s, float &qa, float &qb, float &qc) const
    {
        qa = qb = qc = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float t = s[offy[j] + offx[i]];
                qa = fmaf(wa[j * 4 + i], t, qa);
                qb = fmaf(wb[j * 4 + i], t, qb);
                qc = fmaf(wc[j * 4 + i], t, qc);
            }
    }
};

template <int INTERP, int RULE, bool OOB_ZERO>
__global__ void __launch_bounds__(NT) vt_slice_kernel(const __grid_constant__ VtResampleParams P, int z_chunk)
{
    __shared__ float ring[NSTAGE][STAGE];
    using T = Taps<INTERP>;
    const int tid = threadIdx.x;
    const int ntx = (P.o2 + TS - 1) / TS;
    const int tile_y = blockIdx.x / ntx, tile_x = blockIdx.x - tile_y * ntx;
    const int mat = blockIdx.z;
    const VtMat &M = P.mats[mat];
    const int t0 = (int)M.r[0][3];
    const int zc0 = P.z_begin + blockIdx.y * z_chunk;
    const int zc1 = min(zc0 + z_chunk, P.z_end);
    const int a1_0 = tile_y * TS, a2_0 = tile_x * TS;
    const int a1_1 = min(a1_0 + TS, P.o1) - 1, a2_1 = min(a2_0 + TS, P.o2) - 1;

    // footprint of the tile: extremes are at the corners (the float recipe is monotone in a1 and in a2)
    float y_min, y_max, x_min, x_max;
    {
        const float c1a = (float)a1_0, c1b = (float)a1_1, c2a = (float)a2_0, c2b = (float)a2_1;
        const float y00 = inplane_coord(M.r[1], c1a, c2a), y01 = inplane_coord(M.r[1], c1a, c2b);
        const float y10 = inplane_coord(M.r[1], c1b, c2a), y11 = inplane_coord(M.r[1], c1b, c2b);
        const float x00 = inplane_coord(M.r[2], c1a, c2a), x01 = inplane_coord(M.r[2], c1a, c2b);
        const float x10 = inplane_coord(M.r[2], c1b, c2a), x11 = inplane_coord(M.r[2], c1b, c2b);
        y_min = fminf(fminf(y00, y01), fminf(y10, y11));
        y_max = fmaxf(fmaxf(y00, y01), fmaxf(y10, y11));
        x_min = fminf(fminf(x00, x01), fminf(x10, x11));
        x_max = fmaxf(fmaxf(x00, x01), fmaxf(x10, x11));
    }
    // clamp the footprint to one texel around the source: everything further out is border (zero) anyway and
    // columns that far out are out of bounds; keeps the integer conversions safe for wild matrices
    y_min = fmaxf(y_min, -2.0f); x_min = fmaxf(x_min, -2.0f);
    y_max = fminf(y_max, (float)P.s1 + 2.0f); x_max = fminf(x_max, (float)P.s2 + 2.0f);
    const int ylo = (int)floorf(y_min - 0.5f) + T::LO, yhi = (int)floorf(y_max - 0.5f) + T::HI;
    const int xlo = (int)floorf(x_min - 0.5f) + T::LO, xhi = (int)floorf(x_max - 0.5f) + T::HI;
    int bh = min(yhi - ylo + 1, BMAX), bw = min(xhi - xlo + 1, BMAX);  // host guarantees <= BMAX
    if (bh <= 0 || bw <= 0) bh = bw = 0;  // tile entirely outside the source: nothing to stage

    // per-thread share of the footprint: constant along the march
    int goff[EPT], soff[EPT];
    bool gval[EPT];
    const int nel = bh * bw;
#pragma unroll
    for (int k = 0; k < EPT; k++) {
        const int e = tid + k * NT;
        const int r = bw > 0 ? e / bw : 0, c = e - r * bw;
        const int y = ylo + r, x = xlo + c;
        soff[k] = e < nel ? r * PITCH + c : -1;
        gval[k] = e < nel && (unsigned)y < (unsigned)P.s1 && (unsigned)x < (unsigned)P.s2;
        goff[k] = gval[k] ? y * P.s2 + x : 0;
    }
    const size_t plane_elems = (size_t)P.s1 * P.s2;
    auto issue = [&](int q) {  // stage input plane q (may lie outside the source: all zero)
        float *dst = ring[((q % NSTAGE) + NSTAGE) % NSTAGE];
        const bool zin = (unsigned)q < (unsigned)P.s0;
        const float *src = P.src + (zin ? (size_t)q * plane_elems : 0);
#pragma unroll
        for (int k = 0; k < EPT; k++)
            if (soff[k] >= 0) cp_async4_zfill(dst + soff[k], src + goff[k], zin && gval[k]);
        cp_async_commit();
    };

    // this thread's column
    const int ty = tid / TS, tx = tid - ty * TS;
    const int a1 = a1_0 + ty, a2 = a2_0 + tx;
    const bool live = a1 < P.o1 && a2 < P.o2;
    const float p1 = inplane_coord(M.r[1], (float)a1, (float)a2);
    const float p2 = inplane_coord(M.r[2], (float)a1, (float)a2);
    // transforms.py:276-278 for the two in-plane axes
    const bool inplane = live && !(p2 < 0 || p1 < 0 || p2 >= (float)P.s2 || p1 >= (float)P.s1);
    T taps;
    if (inplane) taps.template init<RULE>(p1, p2, ylo, xlo);
    float *__restrict__ dst = P.dst + (size_t)mat * P.dst_batch_stride + ((size_t)a1 * P.o2 + a2);
    const size_t oplane = (size_t)P.o1 * P.o2;

    // input planes needed: q = z + t0 + d, d in [-PLANES_BEFORE, PLANES_AFTER]
    const int q_first = zc0 + t0 - T::PLANES_BEFORE, q_last = zc1 - 1 + t0 + T::PLANES_AFTER;
    issue(q_first);
    if (q_first + 1 <= q_last) issue(q_first + 1); else cp_async_commit();
    float s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;  // sliding window of per-plane sums
    for (int q = q_first; q <= q_last; q++) {
        cp_async_wait<1>();
        __syncthreads();  // plane q has landed for every thread; everyone is done with plane q-1
        if (q + 2 <= q_last) issue(q + 2); else cp_async_commit();
        const float *s = ring[((q % NSTAGE) + NSTAGE) % NSTAGE];
        const int z = q - t0 - T::PLANES_AFTER;  // output plane completed by input plane q
        float r = 0.0f;
        if (inplane) {
            if constexpr (INTERP == VT_LINEAR) {
                r = taps.plane(s);
            } else if constexpr (INTERP == VT_CUBIC_SIMPLE) {
                const float pq = taps.plane(s);
                // the reference accumulates kz = -1, 0, 1 in that order (helper_interpolation.h:51)
                r = fmaf(taps.wz2, pq, fmaf(taps.wz1, s1, __fmul_rn(taps.wz0, s2)));
                s2 = s1;
                s1 = pq;
            } else {
                float qa, qb, qc;
                taps.plane3(s, qa, qb, qc);
                r = (s3 + s1) + qc;  // s3 = A-sum of plane q-2, s1 = B-sum of plane q-1
                s3 = s2;             // s2 = A-sum of plane q-1
                s2 = qa;
                s1 = qb;
            }
        }
        if (z >= zc0 && live) {
            const bool zok = (unsigned)(z + t0) < (unsigned)P.s0;  // 0 <= p0 < s0 with p0 = z + t0 + 0.5
            if (inplane && zok) dst[(size_t)z * oplane] = r;
            else if (OOB_ZERO) dst[(size_t)z * oplane] = 0.0f;
        }
    }
}

// host side: can this batch run on the slice family?
bool slice_ok(const VtResampleParams &P, int interp)
{
    if (P.s0 >= 16384) return false;  // h0z = idx + 0.3 must keep its 1/256 quantum (see Taps<VT_CUBIC_TEX>)
    for (int k = 0; k < P.n_mats; k++) {
        const VtMat &M = P.mats[k];
        if (M.r[0][0] != 1.0f || M.r[0][1] != 0.0f || M.r[0][2] != 0.0f) return false;
        if (M.r[1][0] != 0.0f || M.r[2][0] != 0.0f) return false;
        const float t0 = M.r[0][3];
        if (!(fabsf(t0) < 16384.0f) || t0 != floorf(t0)) return false;
        for (int r = 1; r <= 2; r++) {
            const float ext = (fabsf(M.r[r][1]) + fabsf(M.r[r][2])) * (float)(TS - 1);
            if (!(ext <= (float)(BMAX - 5) - 0.1f)) return false;  // footprint edge <= ext + 5 texels
            if (!(fabsf(M.r[r][3]) < 1e6f)) return false;
        }
    }
    return true;
}

template <int INTERP, int RULE>
int launch2(const VtResampleParams &P, cudaStream_t st)
{
    const int nz = P.z_end - P.z_begin;
    const int tiles = ((P.o1 + TS - 1) / TS) * ((P.o2 + TS - 1) / TS);
    // enough CTAs for ~3 waves of 148 SMs x resident CTAs, without making z-chunks so short that the
    // warm-up planes (2 per chunk for the cubic modes) cost more than a few percent
    int chunks = (148 * 8 * 3 + tiles * P.n_mats - 1) / (tiles * P.n_mats);
    chunks = max(1, min(chunks, nz / 32 > 0 ? nz / 32 : 1));
    const int z_chunk = (nz + chunks - 1) / chunks;
    chunks = (nz + z_chunk - 1) / z_chunk;
    dim3 grid(tiles, chunks, P.n_mats);
    if (chunks > 65535 || P.n_mats > 65535) return VT_ERR_UNSUPPORTED;
    {
        VtProf prof(VT_K_SLICE_LINEAR + INTERP, st);
        if (P.flags & VT_OOB_ZERO) vt_slice_kernel<INTERP, RULE, true><<<grid, NT, 0, st>>>(P, z_chunk);
        else vt_slice_kernel<INTERP, RULE, false><<<grid, NT, 0, st>>>(P, z_chunk);
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

int vt_slice_supported(const VtResampleParams &P, int interp) { return slice_ok(P, interp) ? 1 : 0; }

int vt_launch_slice(const VtResampleParams &P, int interp, cudaStream_t st)
{
    if (P.z_end <= P.z_begin || P.o1 <= 0 || P.o2 <= 0 || P.n_mats <= 0) return VT_OK;
    switch (interp) {
        case VT_LINEAR: return launch1<VT_LINEAR>(P, st);
        case VT_CUBIC_TEX: return launch1<VT_CUBIC_TEX>(P, st);
        case VT_CUBIC_SIMPLE: return launch1<VT_CUBIC_SIMPLE>(P, st);
    }
    return VT_ERR_INVALID_ARG;
}
