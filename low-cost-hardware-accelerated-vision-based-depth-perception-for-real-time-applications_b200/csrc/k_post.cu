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

// ---- gap interpolation ------------------------------------------------------------------------------------
// For an invalid pixel only the nearest valid pixel on each side of its line (row, then column) matters: with p = previous valid
// position and n = next valid position in the state BEFORE the pass (values dp, dn; fills never create a boundary for another gap),
//   p and n exist, n-p-1 <= gap : fill with (dp+dn)/2 if |dp-dn| < 3 else min(dp,dn)       (elas.cpp:1156-1176, 1220-1293)
//   only n exists (n = first valid), add_corners, n-u <= gap : fill with dn                 (elas.cpp:1191-1201)
//   only p exists (p = last valid),  add_corners, u-p <= gap : fill with dp                 (elas.cpp:1204-1214)
// Both passes therefore work on VALIDITY BITS: one word per 32 positions of a line, the last valid position before each word and the
// first one after it.  The map is read once (row pass, all loads independent); only words with a gap are looked at again, and only
// the gap's two end values are loaded.  The row pass leaves the validity bits of its RESULT (one word per 32 columns of a row) in a
// scratch array, from which the column pass builds its column words by 32 x 32 bit transposes instead of reading the map again.
#ifndef SVB_GAP_WARPS
#define SVB_GAP_WARPS 8
#endif
constexpr int GAP_WARPS = SVB_GAP_WARPS;

__device__ __forceinline__ float gap_fill_value(float d1, float d2) {
    if (fabsf(__fsub_rn(d1, d2)) < 3.0f) return __fdiv_rn(__fadd_rn(d1, d2), 2.0f);
    return d2 < d1 ? d2 : d1;  // std::min(d1, d2)
}

// Row pass, one warp per row.  Dynamic smem: GAP_WARPS * 3 * Cpad words (validity words, carry, ncarry per 32-column chunk).
// grid: (ceil(H / GAP_WARPS), nimg)
// labels_all != nullptr: the pruning step of the speckle removal happens here, on the fly (k_ccl_prune's rule: a valid pixel whose
// component has fewer than min_size pixels becomes -10; most pixels carry CCL_KEPT and need no look-up).
__global__ void __launch_bounds__(GAP_WARPS * 32) k_gap_rows(float *__restrict__ D_all, uint32_t *__restrict__ bits_all, int W, int H, int Cw,
                                                             int Cpad, int gap_width, int add_corners, const int32_t *__restrict__ labels_all,
                                                             const int32_t *__restrict__ sizes_all, int min_size) {
    extern __shared__ int s_gap[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int v = blockIdx.x * GAP_WARPS + wid;
    if (v >= H) return;  // warp-uniform; the kernel has no CTA-wide barrier
    float *line = D_all + ((size_t)blockIdx.y * H + v) * W;
    const int32_t *labels = labels_all ? labels_all + (size_t)blockIdx.y * H * W : nullptr;  // the image's label map (indices are image-relative)
    const int32_t *sizes = labels_all ? sizes_all + (size_t)blockIdx.y * H * W : nullptr;
    uint32_t *words = reinterpret_cast<uint32_t *>(s_gap) + (size_t)wid * 3 * Cpad;
    int *carry = s_gap + (size_t)wid * 3 * Cpad + Cpad, *ncarry = carry + Cpad;
    // 1: validity words (four independent loads in flight per lane)
    for (int k0 = 0; k0 < Cw; k0 += 4) {
        float val[4];
        int lab[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int u = (k0 + i) * 32 + lane;
            val[i] = u < W ? line[u] : -1.f;
            lab[i] = (labels && u < W) ? labels[(size_t)v * W + u] : CCL_KEPT;
        }
        if (labels) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                // lab is the pixel's tile-local root (k_ccl_totals left every such root pointing straight at its component's root)
                if (val[i] >= 0.f && !(lab[i] & CCL_KEPT) && sizes[labels[lab[i]]] < min_size) {
                    val[i] = -10.f;
                    line[(k0 + i) * 32 + lane] = -10.f;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, val[i] >= 0.f);
            if (lane == 0 && k0 + i < Cw) words[k0 + i] = bal;
        }
    }
    __syncwarp();
    // 2: last valid column before / first valid column after every chunk: lane j scans chunks j*per .. j*per+per-1, the lanes are
    // chained by a prefix maximum / suffix minimum
    const int per = (Cw + 31) >> 5;
    const int k_lo = min(lane * per, Cw), k_hi = min(k_lo + per, Cw);
    int last = -1, first = 0x7FFFFFFF;
    for (int k = k_lo; k < k_hi; k++) {
        const uint32_t w = words[k];
        if (w) {
            last = k * 32 + 31 - __clz(w);
            if (first == 0x7FFFFFFF) first = k * 32 + __ffs(w) - 1;
        }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int a = __shfl_up_sync(0xFFFFFFFFu, last, off), b = __shfl_down_sync(0xFFFFFFFFu, first, off);
        if (lane >= off) last = max(last, a);
        if (lane + off < 32) first = min(first, b);
    }
    int run = __shfl_up_sync(0xFFFFFFFFu, last, 1), nrun = __shfl_down_sync(0xFFFFFFFFu, first, 1);
    if (lane == 0) run = -1;
    if (lane == 31) nrun = 0x7FFFFFFF;
    for (int k = k_lo; k < k_hi; k++) {
        carry[k] = run;
        const uint32_t w = words[k];
        if (w) run = k * 32 + 31 - __clz(w);
    }
    for (int k = k_hi - 1; k >= k_lo; k--) {
        ncarry[k] = nrun == 0x7FFFFFFF ? -1 : nrun;
        const uint32_t w = words[k];
        if (w) nrun = k * 32 + __ffs(w) - 1;
    }
    __syncwarp();
    // 3: chunks with a gap
    for (int k = 0; k < Cw; k++) {
        const uint32_t w = words[k];
        const int u = k * 32 + lane;
        const uint32_t in_row = (k * 32 + 32 <= W) ? 0xFFFFFFFFu : (0xFFFFFFFFu >> (k * 32 + 32 - W));
        if (w == in_row) continue;  // warp-uniform
        bool do_fill = false;
        if (!((w >> lane) & 1u) && u < W) {
            const uint32_t below = w & ((1u << lane) - 1u), above = lane == 31 ? 0u : (w & ~((2u << lane) - 1u));
            const int p = below ? (k * 32 + 31 - __clz(below)) : carry[k];
            const int nx = above ? (k * 32 + __ffs(above) - 1) : ncarry[k];
            SVB_GUARD_ASSERT(p >= -1 && p < u && nx < W && (nx < 0 || nx > u));
            float fill = 0.f;
            // the values at p and nx are valid pixels, which this pass never modifies: reading them while other lanes write
            // invalid positions is race free
            if (p >= 0 && nx >= 0) {
                if (nx - p - 1 <= gap_width) {
                    fill = gap_fill_value(line[p], line[nx]);
                    do_fill = true;
                }
            } else if (add_corners && p < 0 && nx >= 0) {
                if (nx - u <= gap_width) {
                    fill = line[nx];
                    do_fill = true;
                }
            } else if (add_corners && p >= 0 && nx < 0) {
                if (u - p <= gap_width) {
                    fill = line[p];
                    do_fill = true;
                }
            }
            if (do_fill) line[u] = fill;
        }
        // every fill value is >= 0 (a valid disparity, a mean or a minimum of two): the filled pixels are valid for the column pass
        const unsigned filled = __ballot_sync(0xFFFFFFFFu, do_fill);
        if (lane == 0) words[k] = w | filled;
    }
    __syncwarp();
    uint32_t *out = bits_all + ((size_t)blockIdx.y * H + v) * Cw;
    for (int k = lane; k < Cw; k += 32) out[k] = words[k];
}

// Column pass: a CTA owns 32 neighbouring columns (lane = column) over all rows.  Dynamic smem: 3 * HW * 32 words (column validity
// words [k][lane], carry, ncarry), HW = ceil(H / 32).
// grid: (Cw, nimg)
__global__ void __launch_bounds__(GAP_WARPS * 32) k_gap_cols(float *__restrict__ D_all, const uint32_t *__restrict__ bits_all, int W, int H, int Cw,
                                                             int HW, int gap_width, int add_corners) {
    extern __shared__ int s_gap[];
    uint32_t *colw = reinterpret_cast<uint32_t *>(s_gap);
    int *carry = s_gap + (size_t)HW * 32, *ncarry = carry + (size_t)HW * 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int u = blockIdx.x * 32 + lane;
    const uint32_t *bits = bits_all + (size_t)blockIdx.y * H * Cw + blockIdx.x;
    // 1: lane r reads the row word of row 32 k + r; 32 ballots transpose the 32 x 32 bit block into one word per column
    for (int k = wid; k < HW; k += GAP_WARPS) {
        const int v = k * 32 + lane;
        const uint32_t rw = v < H ? bits[(size_t)v * Cw] : 0u;
        uint32_t word = 0u;
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const unsigned b = __ballot_sync(0xFFFFFFFFu, (rw >> c) & 1u);
            if (lane == c) word = b;
        }
        colw[k * 32 + lane] = word;
    }
    __syncthreads();
    // 2: per column, the last valid row before / the first valid row after every word (warp 0 downwards, warp 1 upwards)
    if (wid == 0) {
        int run = -1;
        for (int k = 0; k < HW; k++) {
            carry[k * 32 + lane] = run;
            const uint32_t w = colw[k * 32 + lane];
            if (w) run = k * 32 + 31 - __clz(w);
        }
    } else if (wid == 1) {
        int nrun = -1;
        for (int k = HW - 1; k >= 0; k--) {
            ncarry[k * 32 + lane] = nrun;
            const uint32_t w = colw[k * 32 + lane];
            if (w) nrun = k * 32 + __ffs(w) - 1;
        }
    }
    __syncthreads();
    if (u >= W) return;
    // 3: every lane walks the gaps of its column word by word; a maximal run of invalid rows inside a word has its end points either
    // in the word (the neighbouring bits) or in carry / ncarry
    float *D = D_all + (size_t)blockIdx.y * W * H + u;
    for (int k = wid; k < HW; k += GAP_WARPS) {
        const uint32_t inv = ~colw[k * 32 + lane];  // rows >= H are invalid: a run that reaches them has no next valid row in the word
        const int rows = min(32, H - k * 32);
        uint32_t todo = rows == 32 ? inv : (inv & ((1u << rows) - 1u));
        while (todo) {
            const int r0 = __ffs(todo) - 1;
            const uint32_t t = ~(inv >> r0);  // first zero of the shifted word = end of the run
            const int len = t ? __ffs(t) - 1 : 32 - r0;
            const int r1 = min(r0 + len, rows);  // rows r0 .. r1-1 are this word's part of the gap
            todo = r0 + len >= 32 ? 0u : (todo & (0xFFFFFFFFu << (r0 + len)));
            const int p = r0 > 0 ? k * 32 + r0 - 1 : carry[k * 32 + lane];
            const int nx = r0 + len < 32 ? k * 32 + r0 + len : ncarry[k * 32 + lane];
            SVB_GUARD_ASSERT(p >= -1 && p < k * 32 + r0 && nx < H && (nx < 0 || nx >= k * 32 + r1));
            if (p >= 0 && nx >= 0) {
                if (nx - p - 1 <= gap_width) {
                    const float fill = gap_fill_value(D[(size_t)p * W], D[(size_t)nx * W]);
                    for (int r = r0; r < r1; r++) D[(size_t)(k * 32 + r) * W] = fill;
                }
            } else if (add_corners && p < 0 && nx >= 0) {
                const float fill = D[(size_t)nx * W];
                for (int r = max(r0, nx - gap_width - k * 32); r < r1; r++) D[(size_t)(k * 32 + r) * W] = fill;
            } else if (add_corners && p >= 0 && nx < 0) {
                const float fill = D[(size_t)p * W];
                for (int r = r0; r < min(r1, p + gap_width + 1 - k * 32); r++) D[(size_t)(k * 32 + r) * W] = fill;
            }
        }
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

// opt-in for more than 48 KB of dynamic shared memory, once per device and kernel
template <typename K>
int gap_smem_optin(K kernel, size_t smem, size_t (&configured)[64]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(gap kernel, %zu): %s", smem, cudaGetErrorString(e));
            return SVB_ERR_CUDA;
        }
        configured[dev] = smem;
    }
    return SVB_OK;
}

}  // namespace

size_t gap_scratch_words(const Dims &d, int nimg) { return (size_t)nimg * d.Dh * ((d.Dw + 31) / 32); }

// scratch: gap_scratch_words(d, nimg) words (the validity bits the row pass hands to the column pass)
int launch_gap(const Dims &d, const svb_params &p, float *D, uint32_t *scratch, const int32_t *labels, const int32_t *sizes, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    const int W = d.Dw, H = d.Dh, Cw = (W + 31) / 32, HW = (H + 31) / 32;
    const int Cpad = (Cw + 3) & ~3;
    {
        const size_t smem = (size_t)GAP_WARPS * 3 * Cpad * sizeof(int);  // 24 KB at the largest supported width
        static size_t configured[64] = {};
        SVB_TRY(gap_smem_optin(k_gap_rows, smem, configured));
        dim3 grid((H + GAP_WARPS - 1) / GAP_WARPS, nimg);
        k_gap_rows<<<grid, GAP_WARPS * 32, smem, s>>>(D, scratch, W, H, Cw, Cpad, gap_width_of(d, p), p.add_corners, labels, sizes, ccl_min_size(d, p));
        SVB_LAUNCH_CHECK();
    }
    {
        const size_t smem = (size_t)3 * HW * 32 * sizeof(int);  // 96 KB at the largest supported height
        static size_t configured[64] = {};
        SVB_TRY(gap_smem_optin(k_gap_cols, smem, configured));
        dim3 grid(Cw, nimg);
        k_gap_cols<<<grid, GAP_WARPS * 32, smem, s>>>(D, scratch, W, H, Cw, HW, gap_width_of(d, p), p.add_corners);
        SVB_LAUNCH_CHECK();
    }
    return SVB_OK;
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
