// Left/right consistency check, gap interpolation, adaptive mean and median filters.
//
// Replaces Elas::leftRightConsistencyCheck, gapInterpolation, adaptiveMean (full-resolution branch) and median
// (src/serial_includes/elas/elas.cpp:946-1011, 1126-1295, 1297-1494, 1496-1559).  All of these are a few bytes
// of HBM traffic per pixel; the kernels are laid out so that every global access is coalesced along image rows.
#include "post_device.cuh"
#include "svb_internal.h"

namespace svb {

namespace {

// ---- L/R check ------------------------------------------------------------------------------------
// grid: (ceil(W/256), H, nf).  Out of place: the reference works on copies of both maps (elas.cpp:956-959).
constexpr int LR_ROWS = 8;  // rows per thread: fewer, longer CTAs and 8 independent loads in flight per thread

// grid: (ceil(W/256), ceil(rows/LR_ROWS), nf)
__global__ void __launch_bounds__(256) k_lr_check(const float *__restrict__ D1in, const float *__restrict__ D2in, float *__restrict__ D1out,
                                                 float *__restrict__ D2out, int W, int H, float lr_threshold, int row0, int row1, int half) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= W) return;
    const float fw = (float)W;
    const int v_end = min(row0 + (int)(blockIdx.y + 1) * LR_ROWS, row1);
#pragma unroll 4
    for (int v = row0 + blockIdx.y * LR_ROWS; v < v_end; v++) {
        const size_t base = (size_t)blockIdx.z * (unsigned)(H * W) + (unsigned)(v * W);
        const float d1 = D1in[base + u];
        const float d2 = D2in[base + u];
        float o1 = -10.f, o2 = -10.f;
        // subsampling (elas.cpp:972-975): the maps are half size, the disparities are not: warp by d/2
        const float uw1 = __fsub_rn((float)u, half ? __fdiv_rn(d1, 2.f) : d1);
        if (d1 >= 0.f && uw1 >= 0.f && uw1 < fw) {
            const float other = D2in[base + (int)uw1];
            o1 = (fabsf(__fsub_rn(other, d1)) > lr_threshold) ? -10.f : d1;
        }
        const float uw2 = __fadd_rn((float)u, half ? __fdiv_rn(d2, 2.f) : d2);
        if (d2 >= 0.f && uw2 >= 0.f && uw2 < fw) {
            const float other = D1in[base + (int)uw2];
            o2 = (fabsf(__fsub_rn(other, d2)) > lr_threshold) ? -10.f : d2;
        }
        D1out[base + u] = o1;
        if (D2out) D2out[base + u] = o2;
    }
}

// ---- gap interpolation, row pass --------------------------------------------------------------------
// For an invalid pixel only the nearest valid pixel on each side matters: with p = previous valid column and
// n = next valid column (values dp, dn, untouched by the pass),
//   p and n exist, n-p-1 <= gap : fill with (dp+dn)/2 if |dp-dn| < 3 else min(dp,dn)       (elas.cpp:1156-1176)
//   only n exists (n = first valid), add_corners, n-u <= gap : fill with dn                 (elas.cpp:1191-1201)
//   only p exists (p = last valid),  add_corners, u-p <= gap : fill with dp                 (elas.cpp:1204-1214)
// One warp owns one row: a forward sweep records p per column in shared memory, a backward sweep carries n.
constexpr int GAP_WARPS = 4;

__device__ __forceinline__ float gap_fill_value(float d1, float d2) {
    if (fabsf(__fsub_rn(d1, d2)) < 3.0f) return __fdiv_rn(__fadd_rn(d1, d2), 2.0f);
    return d2 < d1 ? d2 : d1;  // std::min(d1, d2)
}

// One line (a row in global memory, or a column of a shared-memory strip) handled by one warp: `line[i * stride]`,
// i = 0 .. n-1; prev = n ints of scratch.
__device__ __forceinline__ void gap_line(float *line, int stride, int n, int *prev, int lane, int gap_width, int add_corners) {
    int carry = -1;
    const int chunks = (n + 31) / 32;
    for (int k = 0; k < chunks; k++) {
        const int u = k * 32 + lane;
        const bool valid = (u < n) && (line[u * stride] >= 0.f);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, valid);
        const unsigned below = bal & ((1u << lane) - 1u);
        if (u < n) prev[u] = below ? (k * 32 + 31 - __clz(below)) : carry;
        if (bal) carry = k * 32 + 31 - __clz(bal);
    }
    __syncwarp();
    int ncarry = -1;  // next valid position beyond the current chunk
    for (int k = chunks - 1; k >= 0; k--) {
        const int u = k * 32 + lane;
        const float val = (u < n) ? line[u * stride] : -1.f;
        const bool valid = (u < n) && (val >= 0.f);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, valid);
        const unsigned above = (lane == 31) ? 0u : (bal & ~((2u << lane) - 1u));
        const int nx = above ? (k * 32 + __ffs(above) - 1) : ncarry;
        if (u < n && !valid) {
            const int p = prev[u];
            SVB_GUARD_ASSERT(p >= -1 && p < u && nx < n && (nx < 0 || nx > u));
            float fill = 0.f;
            bool do_fill = false;
            if (p >= 0 && nx >= 0) {
                if (nx - p - 1 <= gap_width) {
                    SVB_GUARD_ASSERT(p >= 0 && p < n && nx >= 0 && nx < n);
                    fill = gap_fill_value(line[p * stride], line[nx * stride]);
                    do_fill = true;
                }
            } else if (add_corners && p < 0 && nx >= 0) {
                if (nx - u <= gap_width) {
                    fill = line[nx * stride];
                    do_fill = true;
                }
            } else if (add_corners && p >= 0 && nx < 0) {
                if (u - p <= gap_width) {
                    fill = line[p * stride];
                    do_fill = true;
                }
            }
            // values at p and nx are valid pixels, which this pass never modifies: reading them while other
            // lanes write invalid positions is race free
            if (do_fill) line[u * stride] = fill;
        }
        if (bal) ncarry = k * 32 + __ffs(bal) - 1;
    }
}

// grid: (ceil(H/warps), nimg); dynamic smem: warps * Wpad int32
__global__ void __launch_bounds__(GAP_WARPS * 32) k_gap_rows(float *__restrict__ D_all, int W, int H, int gap_width, int add_corners, int Wpad) {
    extern __shared__ int s_prev[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int v = blockIdx.x * warps + wid;
    if (v >= H) return;
    gap_line(D_all + ((size_t)blockIdx.y * H + v) * W, 1, W, s_prev + wid * Wpad, lane, gap_width, add_corners);
}

// Column pass with the same warp-parallel line algorithm: a CTA stages a strip of GC_COLS columns in shared memory
// (row stride GC_COLS + 1, so that a warp reading 32 consecutive ROWS of one column hits 32 different banks), each of
// its warps handles GC_COLS / GC_WARPS columns, and the strip is written back.  Every invalid pixel only depends on
// the nearest valid pixels above and below in the state BEFORE the pass (fills never create a boundary for another
// gap, elas.cpp:1220-1293), so the sequential walk and this formulation agree exactly.
constexpr int GC_WARPS = 8;

// GC_COLS columns per CTA: 16, or 8 for very tall frames so that the strip still fits in shared memory
template <int GC_COLS>
__global__ void __launch_bounds__(GC_WARPS * 32) k_gap_cols_strip(float *__restrict__ D_all, int W, int H, int gap_width, int add_corners,
                                                                  int Hpad) {
    extern __shared__ float s_gc[];  // [H][GC_COLS + 1] floats, then GC_WARPS * Hpad ints
    float *strip = s_gc;
    int *prev_all = reinterpret_cast<int *>(s_gc + (size_t)H * (GC_COLS + 1));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int u0 = blockIdx.x * GC_COLS;
    float *Dimg = D_all + (size_t)blockIdx.y * W * H;
    for (int i = threadIdx.x; i < H * GC_COLS; i += GC_WARPS * 32) {
        const int v = i / GC_COLS, c = i - v * GC_COLS;
        strip[v * (GC_COLS + 1) + c] = (u0 + c < W) ? Dimg[(size_t)v * W + u0 + c] : -10.f;
    }
    __syncthreads();
    for (int c = wid; c < GC_COLS; c += GC_WARPS)
        if (u0 + c < W) gap_line(strip + c, GC_COLS + 1, H, prev_all + wid * Hpad, lane, gap_width, add_corners);
    __syncthreads();
    for (int i = threadIdx.x; i < H * GC_COLS; i += GC_WARPS * 32) {
        const int v = i / GC_COLS, c = i - v * GC_COLS;
        if (u0 + c < W) Dimg[(size_t)v * W + u0 + c] = strip[v * (GC_COLS + 1) + c];
    }
}

// ---- gap interpolation, column pass -------------------------------------------------------------------
// One thread walks one column top to bottom exactly like elas.cpp:1220-1293; neighbouring threads own
// neighbouring columns, so every step of the walk is a coalesced row access.
__global__ void __launch_bounds__(128) k_gap_cols(float *__restrict__ D_all, int W, int H, int gap_width, int add_corners) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= W) return;
    float *D = D_all + (size_t)blockIdx.y * W * H + u;
    int count = 0;
    int first_valid = -1, last_valid = -1;
    for (int v = 0; v < H; v++) {
        const float val = D[(size_t)v * W];
        if (val >= 0.f) {
            if (count >= 1 && count <= gap_width) {
                const int v_first = v - count, v_last = v - 1;
                if (v_first > 0 && v_last < H - 1) {
                    const float d1 = D[(size_t)(v_first - 1) * W];
                    const float fill = gap_fill_value(d1, val);
                    for (int vc = v_first; vc <= v_last; vc++) D[(size_t)vc * W] = fill;
                }
            }
            count = 0;
            if (first_valid < 0) first_valid = v;
            last_valid = v;
        } else {
            count++;
        }
    }
    if (add_corners && first_valid >= 0) {
        // the first / last valid pixel of the column is the same before and after the interior fill
        const float top = D[(size_t)first_valid * W];
        for (int v2 = max(first_valid - gap_width, 0); v2 < first_valid; v2++) D[(size_t)v2 * W] = top;
        const float bot = D[(size_t)last_valid * W];
        for (int v2 = last_valid + 1; v2 <= min(last_valid + gap_width, H - 1); v2++) D[(size_t)v2 * W] = bot;
    }
}

// ---- adaptive mean ----------------------------------------------------------------------------------
// 8-tap horizontal then 8-tap vertical weighted mean (elas.cpp:1401-1485).  Tap coordinates of a centre c are
// c-4 .. c+3.  The reference keeps the window in a ring buffer indexed by (coordinate mod 8) and sums the SSE
// lanes as ((s0+s1)+s2)+s3 with s_k = term(slot k) + term(slot k+4): taps whose coordinates are congruent mod 4
// are added first, then the four pair sums in the order of (coordinate mod 4).  A thread produces FOUR consecutive
// centres c0 .. c0+3 with c0 a multiple of 4, so that every tap's residue is known at compile time (no dynamic
// indexing) and the 11 loaded values are shared by the four windows.  Both passes are integer/FP32-ALU bound, not
// memory bound (ncu: ALU pipe 73-81 %), which is why instruction count per pixel is what is optimised here.
// mode 0: weight = max(0, 4 - float_and(x - xc, 0x4F000000))   (the serial reference's bit-mask "abs")
// mode 1: weight = max(0, 4 - |x - xc|)                         (the parallel reference)
// Horizontal pass: tmp = (D < 0 ? -10 : 0) overwritten by the filtered value where the reference writes D_tmp.
// (D_tmp is malloc'ed and only partly written in the reference; unwritten valid pixels are DEFINED as 0,
// SURVEY.md finding 5.)  grid: (ceil(ceil(W/4)/128), H, nimg); a thread owns columns c0 .. c0+3
template <int MODE>
__global__ void __launch_bounds__(128) k_mean_h(const float *__restrict__ D_all, float *__restrict__ tmp_all, int W, int H) {
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c0 >= W) return;
    const int v = blockIdx.y;
    const size_t base = (size_t)blockIdx.z * (unsigned)(H * W) + (unsigned)(v * W);
    const float *row = D_all + base;
    float x[11];
#pragma unroll
    for (int k = 0; k < 11; k++) {
        const int cc = c0 - 4 + k;
        const float val = (cc >= 0 && cc < W) ? row[cc] : -10.f;
        x[k] = val < 0.f ? -10.f : val;  // D_copy (elas.cpp:1313-1316)
    }
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = x[j + 4] < 0.f ? -10.f : 0.f;
    if (v >= 3 && v < H - 3) {
        float r;
        if (c0 + 0 >= 4 && c0 + 0 <= W - 4 && mean8<MODE, 0>(x, &r)) out[0] = r;
        if (c0 + 1 >= 4 && c0 + 1 <= W - 4 && mean8<MODE, 1>(x, &r)) out[1] = r;
        if (c0 + 2 >= 4 && c0 + 2 <= W - 4 && mean8<MODE, 2>(x, &r)) out[2] = r;
        if (c0 + 3 >= 4 && c0 + 3 <= W - 4 && mean8<MODE, 3>(x, &r)) out[3] = r;
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (c0 + j < W) tmp_all[base + c0 + j] = out[j];
}

// Vertical pass on tmp, writing D in place.  grid: (ceil(W/128), ceil(H/4), nimg); a thread owns rows r0 .. r0+3 of a column
template <int MODE>
__global__ void __launch_bounds__(128) k_mean_v(const float *__restrict__ tmp_all, float *__restrict__ D_all, int W, int H) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < 3 || u >= W - 3) return;
    const int r0 = blockIdx.y * 4;
    const size_t img = (size_t)blockIdx.z * (unsigned)(H * W);
    float x[11];
#pragma unroll
    for (int k = 0; k < 11; k++) {
        const int rr = r0 - 4 + k;
        x[k] = (rr >= 0 && rr < H) ? tmp_all[img + (unsigned)(rr * W + u)] : -10.f;
    }
    float r;
    if (r0 + 0 >= 4 && r0 + 0 <= H - 4 && mean8<MODE, 0>(x, &r)) D_all[img + (unsigned)((r0 + 0) * W + u)] = r;
    if (r0 + 1 >= 4 && r0 + 1 <= H - 4 && mean8<MODE, 1>(x, &r)) D_all[img + (unsigned)((r0 + 1) * W + u)] = r;
    if (r0 + 2 >= 4 && r0 + 2 <= H - 4 && mean8<MODE, 2>(x, &r)) D_all[img + (unsigned)((r0 + 2) * W + u)] = r;
    if (r0 + 3 >= 4 && r0 + 3 <= H - 4 && mean8<MODE, 3>(x, &r)) D_all[img + (unsigned)((r0 + 3) * W + u)] = r;
}

// Half-resolution variant (subsampling, elas.cpp:1334-1400): 4 taps c-2 .. c+1, ring slots = coordinate mod 4, the four
// terms are added in slot order ((t0 + t1) + t2) + t3.  x[0..6] = values at coordinates c0-2 .. c0+4, c0 % 4 == 0.
template <int MODE, int J>
__device__ __forceinline__ bool mean4(const float (&x)[7], float *out) {
    const float xc = x[J + 2];
    float wsl[4], fsl[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        // window element i sits at coordinate c0 - 2 + J + i, i.e. slot (J + i + 2) mod 4
        const int slot = (J + i + 2) & 3;
        const float w = mean_weight<MODE>(x[J + i], xc);
        wsl[slot] = w;
        fsl[slot] = __fmul_rn(x[J + i], w);
    }
    const float weight_sum = __fadd_rn(__fadd_rn(__fadd_rn(wsl[0], wsl[1]), wsl[2]), wsl[3]);
    const float factor_sum = __fadd_rn(__fadd_rn(__fadd_rn(fsl[0], fsl[1]), fsl[2]), fsl[3]);
    if (weight_sum > 0.f) {
        const float d = __fdiv_rn(factor_sum, weight_sum);
        if (d >= 0.f) {
            *out = d;
            return true;
        }
    }
    return false;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_mean4_h(const float *__restrict__ D_all, float *__restrict__ tmp_all, int W, int H) {
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c0 >= W) return;
    const int v = blockIdx.y;
    const size_t base = (size_t)blockIdx.z * (unsigned)(H * W) + (unsigned)(v * W);
    const float *row = D_all + base;
    float x[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const int cc = c0 - 2 + k;
        const float val = (cc >= 0 && cc < W) ? row[cc] : -10.f;
        x[k] = val < 0.f ? -10.f : val;
    }
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; j++) out[j] = x[j + 2] < 0.f ? -10.f : 0.f;
    if (v >= 3 && v < H - 3) {  // centre c = u - 1 for u = 3 .. W-1
        float r;
        if (c0 + 0 >= 2 && c0 + 0 <= W - 2 && mean4<MODE, 0>(x, &r)) out[0] = r;
        if (c0 + 1 >= 2 && c0 + 1 <= W - 2 && mean4<MODE, 1>(x, &r)) out[1] = r;
        if (c0 + 2 >= 2 && c0 + 2 <= W - 2 && mean4<MODE, 2>(x, &r)) out[2] = r;
        if (c0 + 3 >= 2 && c0 + 3 <= W - 2 && mean4<MODE, 3>(x, &r)) out[3] = r;
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (c0 + j < W) tmp_all[base + c0 + j] = out[j];
}

template <int MODE>
__global__ void __launch_bounds__(128) k_mean4_v(const float *__restrict__ tmp_all, float *__restrict__ D_all, int W, int H) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < 3 || u >= W - 3) return;
    const int r0 = blockIdx.y * 4;
    const size_t img = (size_t)blockIdx.z * (unsigned)(H * W);
    float x[7];
#pragma unroll
    for (int k = 0; k < 7; k++) {
        const int rr = r0 - 2 + k;
        x[k] = (rr >= 0 && rr < H) ? tmp_all[img + (unsigned)(rr * W + u)] : -10.f;
    }
    float r;
    if (r0 + 0 >= 2 && r0 + 0 <= H - 2 && mean4<MODE, 0>(x, &r)) D_all[img + (unsigned)((r0 + 0) * W + u)] = r;
    if (r0 + 1 >= 2 && r0 + 1 <= H - 2 && mean4<MODE, 1>(x, &r)) D_all[img + (unsigned)((r0 + 1) * W + u)] = r;
    if (r0 + 2 >= 2 && r0 + 2 <= H - 2 && mean4<MODE, 2>(x, &r)) D_all[img + (unsigned)((r0 + 2) * W + u)] = r;
    if (r0 + 3 >= 2 && r0 + 3 <= H - 2 && mean4<MODE, 3>(x, &r)) D_all[img + (unsigned)((r0 + 3) * W + u)] = r;
}

// ---- median -------------------------------------------------------------------------------------------
// Median of 7 by a 13-exchange selection network (the reference sorts with an insertion sort, elas.cpp:1519-1528;
// the median is a selection, so any correct method gives the same value; the inputs are never NaN).
// D_temp is calloc'ed (elas.cpp:1506): 0 outside [3,W-3)x[3,H-3).  A thread owns 4 consecutive columns.
// grid: (ceil(ceil(W/4)/128), H, nimg)
__global__ void __launch_bounds__(128) k_median_h(const float *__restrict__ D_all, float *__restrict__ tmp_all, int W, int H) {
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c0 >= W) return;
    const int v = blockIdx.y;
    const size_t base = (size_t)blockIdx.z * (unsigned)(H * W) + (unsigned)(v * W);
    float out[4] = {0.f, 0.f, 0.f, 0.f};
    if (v >= 3 && v < H - 3) {
        float x[10];
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const int cc = c0 - 3 + k;
            x[k] = (cc >= 0 && cc < W) ? D_all[base + cc] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int u = c0 + j;
            if (u >= 3 && u < W - 3) {
                const float own = x[j + 3];
                out[j] = own >= 0.f ? median7(x[j], x[j + 1], x[j + 2], x[j + 3], x[j + 4], x[j + 5], x[j + 6]) : own;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (c0 + j < W) tmp_all[base + c0 + j] = out[j];
}

// grid: (ceil(W/128), ceil(H/4), nimg); a thread owns rows r0 .. r0+3 of a column
__global__ void __launch_bounds__(128) k_median_v(const float *__restrict__ tmp_all, float *__restrict__ D_all, int W, int H) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < 3 || u >= W - 3) return;
    const int r0 = blockIdx.y * 4;
    const size_t img = (size_t)blockIdx.z * (unsigned)(H * W);
    float x[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const int rr = r0 - 3 + k;
        x[k] = (rr >= 0 && rr < H) ? tmp_all[img + (unsigned)(rr * W + u)] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int v = r0 + j;
        if (v >= 3 && v < H - 3) {
            const size_t idx = img + (unsigned)(v * W + u);
            if (D_all[idx] >= 0.f) D_all[idx] = median7(x[j], x[j + 1], x[j + 2], x[j + 3], x[j + 4], x[j + 5], x[j + 6]);
        }
    }
}

}  // namespace

int launch_lr_check(const Dims &d, const svb_params &p, const float *D1in, const float *D2in, float *D1out, float *D2out, int nf,
                    cudaStream_t s) {
    return launch_lr_check_rows(d, p, D1in, D2in, D1out, D2out, nf, 0, d.H, s);
}

int launch_lr_check_rows(const Dims &d, const svb_params &p, const float *D1in, const float *D2in, float *D1out, float *D2out, int nf, int row0,
                         int row1, cudaStream_t s) {
    if (nf <= 0 || row1 <= row0) return SVB_OK;
    if (d.sub) {
        row0 = 0;
        row1 = d.Dh;
    }
    dim3 grid((d.Dw + 255) / 256, (row1 - row0 + LR_ROWS - 1) / LR_ROWS, nf);
    k_lr_check<<<grid, 256, 0, s>>>(D1in, D2in, D1out, D2out, d.Dw, d.Dh, (float)p.lr_threshold, row0, row1, d.sub);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

namespace {

int gap_width_of(const Dims &d, const svb_params &p) { return d.sub ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width; }  // elas.cpp:1131-1135

int launch_gap_rows(const Dims &d, const svb_params &p, float *D, int nimg, cudaStream_t s) {
    const int W = d.Dw, H = d.Dh;
    const int Wpad = (W + 31) & ~31;
    const int warps = GAP_WARPS;  // 4 warps x 8192 px x 4 B = 128 KB at the largest supported width
    const size_t smem = (size_t)warps * Wpad * sizeof(int);
    static size_t configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(k_gap_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_gap_rows, %zu): %s", smem, cudaGetErrorString(e));
            return SVB_ERR_CUDA;
        }
        configured[dev] = smem;
    }
    dim3 grid((H + warps - 1) / warps, nimg);
    k_gap_rows<<<grid, warps * 32, smem, s>>>(D, W, H, gap_width_of(d, p), p.add_corners, Wpad);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_gap_cols(const Dims &d, const svb_params &p, float *D, int nimg, cudaStream_t s) {
    const int W = d.Dw, H = d.Dh;
    const int gap_width = gap_width_of(d, p);
    const int Hpad = (H + 31) & ~31;
    int GC_COLS = 16;
    size_t smem = (size_t)H * (GC_COLS + 1) * sizeof(float) + (size_t)GC_WARPS * Hpad * sizeof(int);
    if (smem > 200 * 1024) {
        GC_COLS = 8;
        smem = (size_t)H * (GC_COLS + 1) * sizeof(float) + (size_t)GC_WARPS * Hpad * sizeof(int);
    }
    if (smem <= 200 * 1024) {
        static size_t configured[64][2] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        const int which = GC_COLS == 16 ? 0 : 1;
        if (smem > 48 * 1024 && dev >= 0 && dev < 64 && configured[dev][which] < smem) {
            cudaError_t e = GC_COLS == 16 ? cudaFuncSetAttribute(k_gap_cols_strip<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                          : cudaFuncSetAttribute(k_gap_cols_strip<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(k_gap_cols_strip): %s", cudaGetErrorString(e));
                return SVB_ERR_CUDA;
            }
            configured[dev][which] = smem;
        }
        dim3 grid((W + GC_COLS - 1) / GC_COLS, nimg);
        if (GC_COLS == 16)
            k_gap_cols_strip<16><<<grid, GC_WARPS * 32, smem, s>>>(D, W, H, gap_width, p.add_corners, Hpad);
        else
            k_gap_cols_strip<8><<<grid, GC_WARPS * 32, smem, s>>>(D, W, H, gap_width, p.add_corners, Hpad);
        SVB_LAUNCH_CHECK();
        return SVB_OK;
    }
    dim3 grid((W + 127) / 128, nimg);
    k_gap_cols<<<grid, 128, 0, s>>>(D, W, H, gap_width, p.add_corners);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace

int launch_gap(const Dims &d, const svb_params &p, float *D, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    SVB_TRY(launch_gap_rows(d, p, D, nimg, s));
    return launch_gap_cols(d, p, D, nimg, s);
}

int launch_adaptive_mean(const Dims &d, int mean_mode, float *D, float *tmp, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    const int W = d.Dw, H = d.Dh;
    const dim3 gh(((W + 3) / 4 + 127) / 128, H, nimg);
    const dim3 gv((W + 127) / 128, (H + 3) / 4, nimg);
    const bool true_abs = mean_mode == SVB_MEAN_TRUE_ABS;
    if (d.sub) {
        if (true_abs) {
            k_mean4_h<1><<<gh, 128, 0, s>>>(D, tmp, W, H);
            SVB_LAUNCH_CHECK();
            k_mean4_v<1><<<gv, 128, 0, s>>>(tmp, D, W, H);
        } else {
            k_mean4_h<0><<<gh, 128, 0, s>>>(D, tmp, W, H);
            SVB_LAUNCH_CHECK();
            k_mean4_v<0><<<gv, 128, 0, s>>>(tmp, D, W, H);
        }
        SVB_LAUNCH_CHECK();
        return SVB_OK;
    }
    if (true_abs) {
        k_mean_h<1><<<gh, 128, 0, s>>>(D, tmp, W, H);
        SVB_LAUNCH_CHECK();
        k_mean_v<1><<<gv, 128, 0, s>>>(tmp, D, W, H);
        SVB_LAUNCH_CHECK();
    } else {
        k_mean_h<0><<<gh, 128, 0, s>>>(D, tmp, W, H);
        SVB_LAUNCH_CHECK();
        k_mean_v<0><<<gv, 128, 0, s>>>(tmp, D, W, H);
        SVB_LAUNCH_CHECK();
    }
    return SVB_OK;
}

int launch_median(const Dims &d, float *D, float *tmp, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    const int W = d.Dw, H = d.Dh;
    const dim3 gh(((W + 3) / 4 + 127) / 128, H, nimg);
    const dim3 gv((W + 127) / 128, (H + 3) / 4, nimg);
    k_median_h<<<gh, 128, 0, s>>>(D, tmp, W, H);
    SVB_LAUNCH_CHECK();
    k_median_v<<<gv, 128, 0, s>>>(tmp, D, W, H);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
