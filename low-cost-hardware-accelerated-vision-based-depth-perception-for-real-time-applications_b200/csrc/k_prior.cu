// Plane fitting per triangle, disparity-candidate grid, and rasterisation of the triangle lists into
// per-pixel "owner" maps (which triangle's plane prior a pixel uses).
//
// Replaces Elas::computeDisparityPlanes -> Matrix::solve (src/serial_includes/elas/elas.cpp:503-575,
// src/common_includes/elas/matrix.cpp:418-510), Elas::createGrid (elas.cpp:577-653) and the scan conversion
// loop of Elas::computeDisparity (elas.cpp:838-941).
#include "svb_internal.h"

namespace svb {

namespace {

// ---- planes -------------------------------------------------------------------------------------
// Gauss-Jordan with full pivoting on a 3x3 double system, operation for operation as matrix.cpp:418-510
// (m = 3, n = 1, eps = 1e-20).  Explicit round-to-nearest intrinsics keep ptxas from contracting a*b-c into
// FMA, which the strict-IEEE x86 oracle never does.
__device__ bool solve3(double A[3][3], double B[3]) {
    int ipiv[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        double big = 0.0;
        int irow = 0, icol = 0;
        for (int j = 0; j < 3; j++)
            if (ipiv[j] != 1)
                for (int k = 0; k < 3; k++)
                    if (ipiv[k] == 0)
                        if (fabs(A[j][k]) >= big) {
                            big = fabs(A[j][k]);
                            irow = j;
                            icol = k;
                        }
        ++ipiv[icol];
        if (irow != icol) {
            for (int l = 0; l < 3; l++) {
                double t = A[irow][l];
                A[irow][l] = A[icol][l];
                A[icol][l] = t;
            }
            double t = B[irow];
            B[irow] = B[icol];
            B[icol] = t;
        }
        if (fabs(A[icol][icol]) < 1e-20) return false;
        const double pivinv = __ddiv_rn(1.0, A[icol][icol]);
        A[icol][icol] = 1.0;
        for (int l = 0; l < 3; l++) A[icol][l] = __dmul_rn(A[icol][l], pivinv);
        B[icol] = __dmul_rn(B[icol], pivinv);
        for (int ll = 0; ll < 3; ll++)
            if (ll != icol) {
                const double dum = A[ll][icol];
                A[ll][icol] = 0.0;
                for (int l = 0; l < 3; l++) A[ll][l] = __dsub_rn(A[ll][l], __dmul_rn(A[icol][l], dum));
                B[ll] = __dsub_rn(B[ll], __dmul_rn(B[icol], dum));
            }
    }
    return true;
}

__device__ void fit_plane(const int32_t *__restrict__ support, int c1, int c2, int c3, bool right, float out[3]) {
    const int c[3] = {c1, c2, c3};
    double A[3][3], B[3];
    for (int r = 0; r < 3; r++) {
        const int u = support[3 * c[r]], v = support[3 * c[r] + 1], d = support[3 * c[r] + 2];
        A[r][0] = (double)(right ? u - d : u);
        A[r][1] = (double)v;
        A[r][2] = 1.0;
        B[r] = (double)d;
    }
    if (solve3(A, B)) {
        out[0] = __double2float_rn(B[0]);
        out[1] = __double2float_rn(B[1]);
        out[2] = __double2float_rn(B[2]);
    } else {
        out[0] = out[1] = out[2] = 0.f;
    }
}

// grid: (ceil(maxT/128), 2 sides, nf).  Side s consumes triangle list s (left / right triangulation).
__global__ void __launch_bounds__(128) k_planes(const int32_t *__restrict__ support_all, const int32_t *__restrict__ tri1_all,
                                               const int32_t *__restrict__ tri2_all, const int32_t *__restrict__ ntri_all,
                                               const int32_t *__restrict__ trioff_all, float *__restrict__ planes1_all,
                                               float *__restrict__ planes2_all, PlaneRec *__restrict__ rec1_all, PlaneRec *__restrict__ rec2_all,
                                               int maxS, int maxT) {
    const int f = blockIdx.z, side = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ntri_all[2 * f + side];
    if (i >= n) return;
    const int32_t *support = support_all + (size_t)f * maxS * 3;
    const int32_t *tri = (side ? tri2_all : tri1_all) + ((size_t)trioff_all[f] + i) * 3;  // packed lists, see stage_host()
    float t1[3], t2[3];
    fit_plane(support, tri[0], tri[1], tri[2], false, t1);
    fit_plane(support, tri[0], tri[1], tri[2], true, t2);
    float *pl = (side ? planes2_all : planes1_all);
    if (pl) {
        float *o = pl + ((size_t)f * maxT + i) * 6;
        o[0] = t1[0];
        o[1] = t1[1];
        o[2] = t1[2];
        o[3] = t2[0];
        o[4] = t2[1];
        o[5] = t2[2];
    }
    // elas.cpp:842-852,910: matching the left image uses t1* and checks t2a; the right image the other way round
    PlaneRec r;
    const float *own = side ? t2 : t1;
    const float other_a = side ? t1[0] : t2[0];
    r.a = own[0];
    r.b = own[1];
    r.c = own[2];
    r.valid = (fabs((double)own[0]) < 0.7 && fabs((double)other_a) < 0.7) ? 1 : 0;
    (side ? rec2_all : rec1_all)[(size_t)f * maxT + i] = r;
}

// ---- grid ---------------------------------------------------------------------------------------
// Bitmask form of the reference's per-cell candidate lists: bit d of cell (x,y) is set iff d is in the list.
// Lists are ascending, so walking the set bits low to high reproduces the reference's evaluation order.
__device__ __forceinline__ int floor_div(int a, int b) {
    int q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}

// grid: (ceil(maxS/128), 2, nf)
__global__ void __launch_bounds__(128) k_grid_scatter(const int32_t *__restrict__ support_all, const int32_t *__restrict__ nsupport_all,
                                                     uint32_t *__restrict__ tmp_all, int maxS, int gw, int gh, int gwords, int grid_size,
                                                     int disp_max) {
    const int f = blockIdx.z, side = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsupport_all[f]) return;
    const int32_t *sp = support_all + ((size_t)f * maxS + i) * 3;
    const int u = sp[0], v = sp[1], d = sp[2];
    // elas.cpp:599-602: left x = u / grid_size (integer), right x = floor((u-d)/grid_size), y = floor(v/grid_size)
    const int x = side ? floor_div(u - d, grid_size) : u / grid_size;
    const int y = floor_div(v, grid_size);
    if (x < 0 || x >= gw || y < 0 || y >= gh) return;
    uint32_t *cell = tmp_all + (((size_t)f * 2 + side) * gw * gh + (size_t)y * gw + x) * gwords;
    const int lo = max(d - 1, 0), hi = min(d + 1, disp_max);
    for (int dd = lo; dd <= hi; dd++) atomicOr(cell + (dd >> 5), 1u << (dd & 31));
}

// 3x3 OR over the FLAT cell index (elas.cpp:612-628: the pointer walk wraps across grid-row ends and never
// writes the first gw+1 / last gw+1 cells).  grid: (ceil(cells*gwords/256), 2, nf)
__global__ void __launch_bounds__(256) k_grid_diffuse(const uint32_t *__restrict__ tmp_all, uint32_t *__restrict__ grid1_all,
                                                     uint32_t *__restrict__ grid2_all, int gw, int gh, int gwords) {
    const int f = blockIdx.z, side = blockIdx.y;
    const int cells = gw * gh;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells * gwords) return;
    const int c = i / gwords, w = i - c * gwords;
    const uint32_t *in = tmp_all + ((size_t)f * 2 + side) * cells * gwords;
    uint32_t r = 0;
    if (c >= gw + 1 && c <= cells - gw - 2) {
        const int offs[9] = {-gw - 1, -gw, -gw + 1, -1, 0, 1, gw - 1, gw, gw + 1};
#pragma unroll
        for (int k = 0; k < 9; k++) r |= in[(size_t)(c + offs[k]) * gwords + w];
    }
    (side ? grid2_all : grid1_all)[(size_t)f * cells * gwords + i] = r;
}

// Expansion into the reference's int32 list layout [count, d0, d1, ...] with stride disp_max+2 (parity taps only).
__global__ void k_grid_expand(const uint32_t *__restrict__ grid, int32_t *__restrict__ out, int cells, int gwords, int disp_max) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    int32_t *o = out + (size_t)c * (disp_max + 2);
    int n = 0;
    for (int d = 0; d <= disp_max; d++)
        if (grid[(size_t)c * gwords + (d >> 5)] >> (d & 31) & 1u) o[1 + n++] = d;
    o[0] = n;
    for (int k = n + 1; k < disp_max + 2; k++) o[k] = 0;
}

// ---- raster -------------------------------------------------------------------------------------
// The reference visits triangles in list order and lets later triangles overwrite earlier ones; findMatch's
// early exits depend on the pixel only, so the surviving value is the one computed with the LAST triangle that
// covers the pixel: owner(u,v) = max triangle index covering (u,v).  One warp scan-converts one triangle with
// the reference's float arithmetic (separate mul/add, truncating conversions) and publishes its index with
// atomicMax.  grid: (ceil(maxT/4), 2 sides, nf), 4 warps per CTA.
__device__ __forceinline__ int f2i_trunc_x86(float x) {
    // cvttss2si: out-of-range and NaN give INT_MIN
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(x);
}
__device__ __forceinline__ int f2u_as_int_x86(float x) {
    // (int32_t)(uint32_t)x as gcc emits it on x86-64: 64-bit cvttss2si, low 32 bits
    if (!(x > -9223373136366403584.0f && x < 9223372036854775808.0f)) return 0;
    return (int)(unsigned)(unsigned long long)__float2ll_rz(x);
}

__global__ void __launch_bounds__(128) k_raster(const int32_t *__restrict__ support_all, const int32_t *__restrict__ tri1_all,
                                               const int32_t *__restrict__ tri2_all, const int32_t *__restrict__ ntri_all,
                                               const int32_t *__restrict__ trioff_all, int32_t *__restrict__ owner1_all,
                                               int32_t *__restrict__ owner2_all, int W, int H, int maxS, int row0, int row1, int sub, int Dw, int Dh, int tag) {
    const int f = blockIdx.z, side = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= ntri_all[2 * f + side]) return;
    const int32_t *support = support_all + (size_t)f * maxS * 3;
    const int32_t *tri = (side ? tri2_all : tri1_all) + ((size_t)trioff_all[f] + i) * 3;
    SVB_GUARD_ASSERT(tri[0] >= 0 && tri[0] < maxS && tri[1] >= 0 && tri[1] < maxS && tri[2] >= 0 && tri[2] < maxS);
    int32_t *owner = (side ? owner2_all : owner1_all) + (size_t)f * Dw * Dh;  // owner map has the disparity map's size

    float tu[3], tv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int c = tri[k];
        const int u = support[3 * c], v = support[3 * c + 1], d = support[3 * c + 2];
        tu[k] = (float)(side ? u - d : u);
        tv[k] = (float)v;
    }
    // elas.cpp:873-884: exchange sort by u, pairs (1,0) (2,0) (2,1)
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < j; k++)
            if (tu[k] > tu[j]) {
                float t = tu[j];
                tu[j] = tu[k];
                tu[k] = t;
                t = tv[j];
                tv[j] = tv[k];
                tv[k] = t;
            }
    const float A_u = tu[0], A_v = tv[0], B_u = tu[1], B_v = tv[1], C_u = tu[2], C_v = tv[2];
    const int iA = f2i_trunc_x86(A_u), iB = f2i_trunc_x86(B_u), iC = f2i_trunc_x86(C_u);
    float AB_a = 0.f, AC_a = 0.f, BC_a = 0.f;
    if (iA != iB) AB_a = __fdiv_rn(__fsub_rn(A_v, B_v), __fsub_rn(A_u, B_u));
    if (iA != iC) AC_a = __fdiv_rn(__fsub_rn(A_v, C_v), __fsub_rn(A_u, C_u));
    if (iB != iC) BC_a = __fdiv_rn(__fsub_rn(B_v, C_v), __fsub_rn(B_u, C_u));
    const float AB_b = __fsub_rn(A_v, __fmul_rn(AB_a, A_u));
    const float AC_b = __fsub_rn(A_v, __fmul_rn(AC_a, A_u));
    const float BC_b = __fsub_rn(B_v, __fmul_rn(BC_a, B_u));

    // part 0: A->B against AC with line AB; part 1: B->C against AC with line BC (elas.cpp:913-941)
#pragma unroll
    for (int part = 0; part < 2; part++) {
        const int i0 = part ? iB : iA, i1 = part ? iC : iB;
        if (i0 == i1) continue;
        const float la = part ? BC_a : AB_a, lb = part ? BC_b : AB_b;
        const int u_lo = max(i0, 0), u_hi = min(i1, W);
        // subsampling (elas.cpp:915-934): only even columns and even rows are matched; pixel (u, v) lives at (u/2, v/2)
        const int ustep = sub ? 2 : 1;
        for (int u = ((u_lo + ustep - 1) & ~(ustep - 1)) + lane * ustep; u < u_hi; u += 32 * ustep) {
            const float fu = (float)u;
            const int v_1 = f2u_as_int_x86(__fadd_rn(__fmul_rn(AC_a, fu), AC_b));
            const int v_2 = f2u_as_int_x86(__fadd_rn(__fmul_rn(la, fu), lb));
            // rows are clipped to [row0, row1): the whole image, or this device's band in the row-band split
            const int v_lo = max(max(min(v_1, v_2), 0), row0), v_hi = min(min(max(v_1, v_2), H), row1);
            if (!sub) {
                for (int v = v_lo; v < v_hi; v++) {
                    SVB_GUARD_ASSERT(u >= 0 && u < W && v >= 0 && v < H);
                    atomicMax(owner + (size_t)v * W + u, tag | i);
                }
            } else if ((u >> 1) < Dw) {
                for (int v = (v_lo + 1) & ~1; v < v_hi; v += 2)
                    if ((v >> 1) < Dh) {
                        SVB_GUARD_ASSERT(u >= 0 && v >= 0);
                        atomicMax(owner + (size_t)(v >> 1) * Dw + (u >> 1), tag | i);
                    }
            }
        }
    }
}

}  // namespace

int launch_planes(const Dims &d, const int32_t *support, const int32_t *tri1, const int32_t *tri2, const int32_t *ntri, const int32_t *trioff,
                  float *planes_ref1, float *planes_ref2, PlaneRec *rec1, PlaneRec *rec2, int nf, int max_tri, cudaStream_t s) {
    if (nf <= 0 || max_tri <= 0) return SVB_OK;
    if (max_tri > d.maxT) max_tri = d.maxT;
    dim3 grid((max_tri + 127) / 128, 2, nf);
    k_planes<<<grid, 128, 0, s>>>(support, tri1, tri2, ntri, trioff, planes_ref1, planes_ref2, rec1, rec2, d.maxS, d.maxT);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_grid(const Dims &d, const svb_params &p, const int32_t *support, const int32_t *nsupport, uint32_t *tmp, uint32_t *grid1,
                uint32_t *grid2, int nf, int max_support, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    if (max_support > d.maxS) max_support = d.maxS;
    const size_t words = (size_t)nf * 2 * d.gw * d.gh * d.gwords;
    cudaError_t e = cudaMemsetAsync(tmp, 0, words * sizeof(uint32_t), s);
    if (e != cudaSuccess) {
        set_error("cudaMemsetAsync(grid tmp): %s", cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    if (max_support > 0) {
        dim3 grid((max_support + 127) / 128, 2, nf);
        k_grid_scatter<<<grid, 128, 0, s>>>(support, nsupport, tmp, d.maxS, d.gw, d.gh, d.gwords, p.grid_size, p.disp_max);
        SVB_LAUNCH_CHECK();
    }
    {
        dim3 grid((d.gw * d.gh * d.gwords + 255) / 256, 2, nf);
        k_grid_diffuse<<<grid, 256, 0, s>>>(tmp, grid1, grid2, d.gw, d.gh, d.gwords);
        SVB_LAUNCH_CHECK();
    }
    return SVB_OK;
}

int launch_grid_expand(const Dims &d, const svb_params &p, const uint32_t *grid, int32_t *grid_ref, cudaStream_t s) {
    const int cells = d.gw * d.gh;
    k_grid_expand<<<(cells + 127) / 128, 128, 0, s>>>(grid, grid_ref, cells, d.gwords, p.disp_max);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_raster(const Dims &d, const int32_t *support, const int32_t *tri1, const int32_t *tri2, const int32_t *ntri, const int32_t *trioff,
                  int32_t *owner1, int32_t *owner2, int nf, int max_tri, cudaStream_t s, int gen) {
    return launch_raster_rows(d, support, tri1, tri2, ntri, trioff, owner1, owner2, nf, max_tri, 0, d.H, s, gen);
}

int launch_raster_rows(const Dims &d, const int32_t *support, const int32_t *tri1, const int32_t *tri2, const int32_t *ntri, const int32_t *trioff,
                       int32_t *owner1, int32_t *owner2, int nf, int max_tri, int row0, int row1, cudaStream_t s, int gen) {
    if (nf <= 0 || row1 <= row0) return SVB_OK;
    if (max_tri > d.maxT) max_tri = d.maxT;
    cudaError_t e = cudaSuccess;
    // gen > 0: the entries carry the generation in bits 24..30 (owner_tagged / owner_untag, svb_internal.h): whatever an earlier
    // generation left in the map is smaller than any entry of this one and reads as "no triangle", so the map is not cleared
    if (gen > 0) {
        if (gen > OWNER_GEN_MAX || d.maxT > OWNER_INDEX_MASK) {
            set_error("launch_raster: generation %d / %d triangles do not fit the owner encoding", gen, d.maxT);
            return SVB_ERR_ARG;
        }
    } else if (row0 == 0 && row1 == d.H) {
        e = cudaMemsetAsync(owner1, 0xFF, (size_t)nf * d.DN * sizeof(int32_t), s);
        if (e == cudaSuccess) e = cudaMemsetAsync(owner2, 0xFF, (size_t)nf * d.DN * sizeof(int32_t), s);
    } else if (d.sub) {
        set_error("row-band raster with subsampling is not supported");
        return SVB_ERR_UNSUPPORTED;
    } else {
        for (int f = 0; f < nf && e == cudaSuccess; f++) {
            const size_t off = (size_t)f * d.N + (size_t)row0 * d.W, cnt = (size_t)(row1 - row0) * d.W * sizeof(int32_t);
            e = cudaMemsetAsync(owner1 + off, 0xFF, cnt, s);
            if (e == cudaSuccess) e = cudaMemsetAsync(owner2 + off, 0xFF, cnt, s);
        }
    }
    if (e != cudaSuccess) {
        set_error("cudaMemsetAsync(owner): %s", cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    if (max_tri <= 0) return SVB_OK;
    dim3 grid((max_tri + 3) / 4, 2, nf);
    k_raster<<<grid, 128, 0, s>>>(support, tri1, tri2, ntri, trioff, owner1, owner2, d.W, d.H, d.maxS, row0, row1, d.sub, d.Dw, d.Dh, gen << OWNER_GEN_SHIFT);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
