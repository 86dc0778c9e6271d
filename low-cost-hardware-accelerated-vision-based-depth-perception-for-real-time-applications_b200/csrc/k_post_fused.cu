// Fused tail of the post-processing chain: adaptive mean (horizontal, vertical) -> median (horizontal, vertical) -> final float map
// -> u8 conversion -> reprojection to 3-D, one kernel, one pass over HBM.
//
// Replaces, for full-resolution maps, the sequence k_mean_h, k_mean_v, k_median_h, k_median_v (k_post.cu), the device-to-device
// copy of the final map and k_reproject (k_reproject.cu), i.e. Elas::adaptiveMean + Elas::median
// (src/serial_includes/elas/elas.cpp:1297-1559) followed by generateDisparityMap's convertTo and projectParallel
// (src/parallel_includes/main/stereo_vision.cu:324,188-212).  Those five kernels moved 4 B in and 4 B out per pixel EACH, through
// two scratch maps; this one reads the gap-interpolated map once (4 B per pixel, halo re-reads come out of L2) and writes the final
// map (4 B), the u8 map (1 B) and the point cloud (24 B) once: 33 B per pixel of algorithmic HBM traffic instead of 65.
//
// Geometry.  A CTA produces PF_TW x TH output pixels.  The four separable passes widen the footprint by 4+3 = 7 pixels on the low
// side and 3+3 = 6 on the high side of either axis; the staged region is the output tile plus a symmetric halo of 8, so that every
// group of four centres (the adaptive mean's summation order depends on the coordinate mod 4, post_device.cuh) starts at a multiple
// of 4 in image AND in tile coordinates and all shared-memory vector accesses are 16-byte aligned:
//     region row r  <->  image row y0 - 8 + r,  r in [0, TH + 16)        region column j  <->  image column x0 - 8 + j,  j in [0, 128)
// Two region-sized float buffers ping-pong:   A = Dg (gap output)   --mean_h-->   B = T   --mean_v (in place on A)-->   A = Dm
//                                             --median_h-->   B = Mh   --median_v + A-->   registers -> global.
// Horizontal passes: a thread owns four consecutive centres of one row (three LDS.128, one STS.128, conflict free);
// vertical passes: a thread owns four consecutive rows of one column (lanes = neighbouring columns, conflict free).
// The per-pixel arithmetic is the SAME device code the stage-by-stage kernels call, and border / out-of-image rules are restated
// from those kernels one by one (comments below), so fused and unfused paths agree bit for bit (tests/test_gpu_parity.py).
#include <stdlib.h>

#include "post_device.cuh"
#include "svb_internal.h"

namespace svb {

namespace {

constexpr int PF_TW = 112;            // output columns per CTA
constexpr int PF_SW = PF_TW + 16;     // region width = shared row stride (floats): 128
constexpr int PF_THREADS = 2 * PF_TW; // 224: in the last pass a thread owns one output column and half of the row groups

struct PostFusedArgs {
    const float *Din;   // [nimg][H][W] gap-interpolated maps
    float *Dout;        // [nimg][H][W] final maps, may be null
    uint8_t *dmap;      // [nimg][H][W] u8 maps, may be null
    double *points;     // [nimg][H][W][3], may be null
    int W, H;
    int has_mean, has_median;
    int float_disp;     // SVB_OUT_POINTS_FLOATDISP: the float disparity itself enters Q (invalid pixels as 0), no 4x u8 quantisation
    Calib cal;
};

template <int MODE, int TH>
__global__ void __launch_bounds__(PF_THREADS, TH == 32 ? 4 : 2) k_post_fused(const PostFusedArgs a) {
    constexpr int RH = TH + 16;  // region rows
    extern __shared__ __align__(16) float s_pf[];
    float *A = s_pf, *B = s_pf + RH * PF_SW;
    const int tid = threadIdx.x;
    const int W = a.W, H = a.H;
    const int x0 = blockIdx.x * PF_TW, y0 = blockIdx.y * TH;
    const size_t img = (size_t)blockIdx.z * (unsigned)(W * H);
    const float *Din = a.Din + img;

    // ---- phase 0: region -> A.  Outside the image: -10 when the mean runs first (k_mean_h / k_mean_v treat every tap outside the
    // image as -10), 0 when the median is the first filter (k_median_h: columns outside the image read as 0)
    const float oob = a.has_mean ? -10.f : 0.f;
    {
        // a warp stages whole region rows: lane l owns region columns 4l .. 4l+3 (four scalar loads -- image rows are only 4-byte
        // aligned -- and one 16-byte shared store); row and column range tests are hoisted out of the element loop
        const int lane = tid & 31, wid = tid >> 5;
        const int u = x0 - 8 + 4 * lane;
        const bool c0 = u >= 0 && u < W, c1 = u + 1 >= 0 && u + 1 < W, c2 = u + 2 >= 0 && u + 2 < W, c3 = u + 3 >= 0 && u + 3 < W;
        for (int r = wid; r < RH; r += PF_THREADS / 32) {
            const int v = y0 - 8 + r;
            float4 q = make_float4(oob, oob, oob, oob);
            if (v >= 0 && v < H) {
                const float *src = Din + (unsigned)(v * W) + u;
                if (c0) q.x = __ldg(src);
                if (c1) q.y = __ldg(src + 1);
                if (c2) q.z = __ldg(src + 2);
                if (c3) q.w = __ldg(src + 3);
            }
            *reinterpret_cast<float4 *>(A + r * PF_SW + 4 * lane) = q;
        }
    }
    __syncthreads();

    if (a.has_mean) {
        // ---- phase 1: horizontal mean, A -> B (= D_tmp of the reference).  Restates k_mean_h: taps are D_copy (invalid -> -10),
        // the output starts as (D < 0 ? -10 : 0) and is overwritten where the reference writes D_tmp: rows [3, H-3), centres [4, W-4]
        for (int i = tid; i < RH * 32; i += PF_THREADS) {
            // a warp takes a compact block of 8 groups (32 pixels) x 4 rows rather than 32 groups of one row: the smooth-window fast
            // path of mean8x4 is decided per warp, and a 32 x 4 block is smooth far more often than a 128 x 1 run
            const int blk = i >> 5, lane_in = i & 31;
            const int r = 4 * (blk >> 2) + (lane_in >> 3), g = 8 * (blk & 3) + (lane_in & 7);
            if (g == 0 || g == 31) continue;  // centres j = 4g .. 4g+3 need taps j-4 .. j+6
            const float *row = A + r * PF_SW + 4 * g;
            SVB_GUARD_ASSERT(row - 4 >= A && row + 8 <= A + RH * PF_SW);
            const float4 lo = *reinterpret_cast<const float4 *>(row - 4), mid = *reinterpret_cast<const float4 *>(row),
                         hi = *reinterpret_cast<const float4 *>(row + 4);
            // D_copy (elas.cpp:1313-1316: invalid -> -10) is the identity here: every negative value of a map that went through the
            // L/R check, the speckle removal and the gap pass is exactly -10 (k_lr_check / k_ccl_prune write nothing else)
            const float x[11] = {lo.x, lo.y, lo.z, lo.w, mid.x, mid.y, mid.z, mid.w, hi.x, hi.y, hi.z};
            const int v = y0 - 8 + r, c0 = x0 - 8 + 4 * g;
            float out[4];
#pragma unroll
            for (int j = 0; j < 4; j++) out[j] = x[j + 4] < 0.f ? -10.f : 0.f;
            // (rows and centre groups outside the filtered range -- tile halos beyond the image, the ragged last tile column -- skip the
            // arithmetic altogether)
            if (v >= 3 && v < H - 3 && c0 + 3 >= 4 && c0 <= W - 4) {
                const bool write[4] = {c0 + 0 >= 4 && c0 + 0 <= W - 4, c0 + 1 >= 4 && c0 + 1 <= W - 4, c0 + 2 >= 4 && c0 + 2 <= W - 4,
                                       c0 + 3 >= 4 && c0 + 3 <= W - 4};
                mean8x4<MODE>(x, write, out);
            }
            *reinterpret_cast<float4 *>(B + r * PF_SW + 4 * g) = make_float4(out[0], out[1], out[2], out[3]);
        }
        __syncthreads();
        // ---- phase 2: vertical mean on B, written in place into A (a pixel the reference does not write keeps its gap-pass value).
        // Restates k_mean_v: columns [3, W-3), centres (rows) [4, H-4], taps outside the image are -10 (B holds -10 there: the region's
        // rows outside the image were loaded as -10, phase 1 leaves them -10)
        for (int i = tid; i < (RH / 4) * PF_SW; i += PF_THREADS) {
            const int j = i & (PF_SW - 1), rg = i >> 7;
            if (rg == 0 || rg == RH / 4 - 1 || j < 4 || j >= PF_SW - 4) continue;
            const int u = x0 - 8 + j, r0 = 4 * rg, v0 = y0 - 8 + r0;
            if (u < 3 || u >= W - 3) continue;
            float x[11];
            SVB_GUARD_ASSERT(r0 - 4 >= 0 && r0 + 6 < RH);
#pragma unroll
            for (int k = 0; k < 11; k++) x[k] = B[(r0 - 4 + k) * PF_SW + j];
            if (v0 + 3 >= 4 && v0 <= H - 4) {
                const bool write[4] = {v0 + 0 >= 4 && v0 + 0 <= H - 4, v0 + 1 >= 4 && v0 + 1 <= H - 4, v0 + 2 >= 4 && v0 + 2 <= H - 4,
                                       v0 + 3 >= 4 && v0 + 3 <= H - 4};
                float res[4] = {A[(r0 + 0) * PF_SW + j], A[(r0 + 1) * PF_SW + j], A[(r0 + 2) * PF_SW + j], A[(r0 + 3) * PF_SW + j]};
                mean8x4<MODE>(x, write, res);
#pragma unroll
                for (int k = 0; k < 4; k++) A[(r0 + k) * PF_SW + j] = res[k];
            }
        }
        __syncthreads();
        // the median reads columns outside the image as 0, the mean read them as -10: only tiles on the left / right image border
        if (a.has_median && (x0 < 8 || x0 + PF_TW + 8 > W)) {
            for (int i = tid; i < RH * PF_SW; i += PF_THREADS) {
                const int u = x0 - 8 + (i & (PF_SW - 1));
                if (u < 0 || u >= W) A[i] = 0.f;
            }
            __syncthreads();
        }
    }

    if (a.has_median) {
        // ---- phase 3: horizontal median, A -> B (= D_temp of the reference, calloc'ed: 0 wherever it is not written).
        // Restates k_median_h: rows [3, H-3), columns [3, W-3): valid pixels get the median of 7 (invalid neighbours take part with their
        // value), invalid ones are copied
        for (int i = tid; i < RH * 32; i += PF_THREADS) {
            const int r = i >> 5, g = i & 31;
            if (g < 2 || g > 29 || r < 4 || r >= RH - 4) continue;  // output columns only, rows y0-4 .. y0+TH+3
            const int v = y0 - 8 + r, c0 = x0 - 8 + 4 * g;
            float out[4] = {0.f, 0.f, 0.f, 0.f};
            if (v >= 3 && v < H - 3) {
                const float *row = A + r * PF_SW + 4 * g;
                const float4 lo = *reinterpret_cast<const float4 *>(row - 4), mid = *reinterpret_cast<const float4 *>(row),
                             hi = *reinterpret_cast<const float4 *>(row + 4);
                const float x[10] = {lo.y, lo.z, lo.w, mid.x, mid.y, mid.z, mid.w, hi.x, hi.y, hi.z};  // columns c0-3 .. c0+6
                float m[4];
                median7x4(x, m);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int u = c0 + j;
                    if (u >= 3 && u < W - 3) {
                        const float own = x[j + 3];
                        out[j] = own >= 0.f ? m[j] : own;
                    }
                }
            }
            *reinterpret_cast<float4 *>(B + r * PF_SW + 4 * g) = make_float4(out[0], out[1], out[2], out[3]);
        }
        __syncthreads();
    }

    // ---- phase 4: vertical median (k_median_v: rows [3, H-3), columns [3, W-3), only where D >= 0; rows of D_temp outside the image
    // read as 0 = what phase 3 wrote for them), then the outputs
    const int j = 8 + (tid % PF_TW), half = tid / PF_TW;
    const int u = x0 - 8 + j;
    if (u >= W) return;
    float *Dout = a.Dout ? a.Dout + img : nullptr;
    uint8_t *dmap = a.dmap ? a.dmap + img : nullptr;
    double *points = a.points ? a.points + img * 3 : nullptr;
    const double fu = (double)u;
    const bool col_in = u >= 3 && u < W - 3;
    for (int rg = half; rg < TH / 4; rg += 2) {
        const int r0 = 8 + 4 * rg, v0 = y0 + 4 * rg;
        if (v0 >= H) break;
        float res[4];
#pragma unroll
        for (int k = 0; k < 4; k++) res[k] = A[(r0 + k) * PF_SW + j];
        if (a.has_median && col_in) {
            float x[10];
            SVB_GUARD_ASSERT(r0 - 3 >= 4 && r0 + 6 < RH - 4);
#pragma unroll
            for (int k = 0; k < 10; k++) x[k] = B[(r0 - 3 + k) * PF_SW + j];
            float m[4];
            median7x4(x, m);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int v = v0 + k;
                if (v >= 3 && v < H - 3 && res[k] >= 0.f) res[k] = m[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = v0 + k;
            if (v >= H) break;
            const unsigned p = (unsigned)(v * W + u);
            if (Dout) Dout[p] = res[k];
            if (dmap || points) {
                const int q = rp_quantise(res[k]);
                if (dmap) dmap[p] = (uint8_t)q;
                if (points) {
                    RpPixel px;
                    const double fy = (double)v;
#pragma unroll
                    for (int t = 0; t < 4; t++) px.base[t] = __fma_rn(fu, a.cal.Q[4 * t + 0], __dmul_rn(fy, a.cal.Q[4 * t + 1]));  // = rp_pixel_xy
                    double out[3];
                    // float_disp: the filtered disparity itself (invalid = negative -> 0, which projects to w = 0 like the u8 path's 0)
                    rp_point_d(a.cal, px, a.float_disp ? (double)fmaxf(res[k], 0.f) : (double)q, out);
                    double *dst = points + (size_t)p * 3;
                    __stcs(dst + 0, out[0]);
                    __stcs(dst + 1, out[1]);
                    __stcs(dst + 2, out[2]);
                }
            }
        }
    }
}

template <int MODE, int TH>
int launch_tile(const PostFusedArgs &a, int nimg, cudaStream_t s) {
    constexpr size_t smem = (size_t)2 * (TH + 16) * PF_SW * sizeof(float);
    static bool configured[64] = {};  // per device: opt in to > 48 KB of dynamic shared memory once
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_post_fused<MODE, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_post_fused): %s", cudaGetErrorString(e));
            return SVB_ERR_CUDA;
        }
        configured[dev] = true;
    }
    dim3 grid((a.W + PF_TW - 1) / PF_TW, (a.H + TH - 1) / TH, nimg);
    k_post_fused<MODE, TH><<<grid, PF_THREADS, smem, s>>>(a);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace

// Full-resolution maps only (the half-resolution chain of param.subsampling keeps the stage-by-stage kernels).
int launch_post_fused(const Dims &d, const svb_params &p, int mean_mode, const Calib &cal, const float *Din, float *Dout, uint8_t *dmap,
                      double *points, int float_disp, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    if (d.sub) {
        set_error("launch_post_fused: full-resolution maps only");
        return SVB_ERR_UNSUPPORTED;
    }
    PostFusedArgs a;
    a.Din = Din;
    a.Dout = Dout;
    a.dmap = dmap;
    a.points = points;
    a.W = d.W;
    a.H = d.H;
    a.has_mean = p.filter_adaptive_mean ? 1 : 0;
    a.has_median = p.filter_median ? 1 : 0;
    a.float_disp = float_disp;
    a.cal = cal;
    // 32-row tiles: four CTAs per SM.  64-row tiles (SVB_PF_TH=64) execute 10 % fewer instructions (halo 80/64 instead of 48/32) but only
    // two CTAs fit an SM, and the kernel is bound by issue latency, not by the instruction count: 11.3 vs 9.3 us per frame measured.
    static const int th_env = getenv("SVB_PF_TH") ? atoi(getenv("SVB_PF_TH")) : 0;
    const bool tall = th_env == 64;
    if (mean_mode == SVB_MEAN_TRUE_ABS) return tall ? launch_tile<1, 64>(a, nimg, s) : launch_tile<1, 32>(a, nimg, s);
    return tall ? launch_tile<0, 64>(a, nimg, s) : launch_tile<0, 32>(a, nimg, s);
}

}  // namespace svb
