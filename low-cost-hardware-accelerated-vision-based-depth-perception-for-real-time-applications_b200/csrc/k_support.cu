// Support-point matching on the candidate lattice, the three order-dependent lattice filters, ordered
// compaction into the support list and the optional corner points.
//
// Replaces Elas::computeMatchingDisparity / computeSupportMatches / removeInconsistentSupportPoints /
// removeRedundantSupportPoints / addCornerSupportPoints (src/serial_includes/elas/elas.cpp:152-440).
#include <stdlib.h>

#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ uint32_t sad16(const uint4 &a, const uint4 &b) {
    // 4 x VABSDIFF4.U8.ACC on sm_100a
    return __vsadu4(a.x, b.x) + __vsadu4(a.y, b.y) + __vsadu4(a.z, b.z) + __vsadu4(a.w, b.w);
}

__device__ __forceinline__ uint32_t texture16(const uint4 &a) {
    const uint4 k = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    return sad16(a, k);  // sum |byte - 128|   (elas.cpp:296-298)
}

// ------------------------------------------------------------------------------------------------
// Matching (elas.cpp:266-371 for every lattice candidate, forward then backward, elas.cpp:394-411).
//
// energy(u, d) = sum over the 4 anchors (u-+2, v-+2) of SAD16(own anchor, other image at the anchor shifted by d)
// depends on the candidate column u and on the MATCHED column x = u -+ d.  The kernel walks over x:
// a warp owns a patch of 8 consecutive candidates x 4 consecutive lattice rows, one candidate per lane, whose
// four anchors (64 B) and running best / second-best keys live in the lane's registers.  In step x every lane
// loads the four "other image" descriptors of column x of ITS lattice row -- 4 distinct addresses per load
// instruction, i.e. 4 L1 wavefronts instead of the 32 a candidate-major walk needs -- and evaluates the one
// hypothesis d = u - x (or x - u).  Nothing is reduced across lanes.  The integer pipe is the limiter (ncu: ALU
// pipe > 80 % busy), so the loop body is kept to 16 chained VABSDIFF4.ACC plus 5 other ALU instructions; the
// patch shape keeps 256 / (256 + 35) = 88 % of the evaluated hypotheses inside their candidates' ranges.
//
// best = smallest energy, lowest d on ties; second = second smallest energy of the multiset (elas.cpp:352-360);
// both follow from keeping the two smallest keys (E << 16 | d), which are distinct per candidate.
constexpr int SM_WARPS = 8;   // warps per CTA: 8 neighbouring patches of the same 4 lattice rows (shared L1 footprint)
constexpr int PATCH_U = 8;    // candidates per patch row
constexpr int PATCH_V = 4;    // lattice rows per patch

__device__ __forceinline__ unsigned sad16_acc(const uint4 &a, const uint4 &b, unsigned acc) {
    // one dependent chain of VABSDIFF4.U8.ACC (no separate adds)
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.x), "r"(b.x));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.y), "r"(b.y));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.z), "r"(b.z));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.w), "r"(b.w));
    return acc;
}

// One pass over the patch: every lane holds one candidate (u, v) of image `own`; returns its disparity or -1.
template <bool right_image, bool COUNT>
__device__ __forceinline__ int match_pass(const uint4 *__restrict__ own, const uint4 *__restrict__ oth, int u, int v, bool has, int W, int H,
                                          int disp_min, int disp_max, int support_texture, float support_threshold, unsigned &n_hyp,
                                          const uint4 *oth_lo, const uint4 *oth_hi) {
    (void)oth_lo;
    (void)oth_hi;
    // candidate validity and disparity range (elas.cpp:279,296-300,318-327)
    bool ok = has && u >= 5 && u <= W - 6 && v >= 5 && v <= H - 6;
    if (ok) ok = (int)texture16(__ldg(own + (size_t)v * W + u)) >= support_texture;
    const int dmin = max(disp_min, 0);
    const int dmax = right_image ? min(disp_max, W - u - 5) : min(disp_max, u - 5);
    ok = ok && (dmax - dmin >= 10);
    if (COUNT) n_hyp += ok ? (unsigned)(dmax - dmin + 1) : 0u;  // elas.cpp:330: every d of the range is evaluated
    if (!__any_sync(0xFFFFFFFFu, ok)) return -1;
    // matched columns x = u - d (left candidate) or u + d (right candidate)
    const int xlo = ok ? (right_image ? u + dmin : u - dmax) : 0x7FFFFFFF;
    const int xhi = ok ? (right_image ? u + dmax : u - dmin) : (int)0x80000000;
    const int vv = ok ? v : 2;  // lanes without a candidate read (and discard) a row that certainly exists
    const uint4 *ot = oth + (size_t)(vv - 2) * W;
    const uint4 *ob = oth + (size_t)(vv + 2) * W;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    if (ok) {
        a0 = __ldg(own + (size_t)(v - 2) * W + u - 2);
        a1 = __ldg(own + (size_t)(v - 2) * W + u + 2);
        a2 = __ldg(own + (size_t)(v + 2) * W + u - 2);
        a3 = __ldg(own + (size_t)(v + 2) * W + u + 2);
    }
    const int wxlo = __reduce_min_sync(0xFFFFFFFFu, xlo);  // >= 5 - 0 ... every x in [wxlo, wxhi] has x-2 >= 0 and x+2 < W
    const int wxhi = __reduce_max_sync(0xFFFFFFFFu, xhi);
    // d = sgn * (u - x); valid iff 0 <= d - dmin <= dmax - dmin.  Lanes without a candidate compute garbage that is
    // discarded below, so `ok` is not part of the loop.
    constexpr int sgn = right_image ? -1 : 1;
    unsigned drel = (unsigned)(sgn * (u - wxlo) - dmin);   // d - dmin at x = wxlo; moves by -sgn per step
    constexpr unsigned dstep = (unsigned)(-sgn);
    const unsigned span = (unsigned)(dmax - dmin);
    unsigned best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;     // (E << 16) | (d - dmin)
    // The descriptor of column x+2 is needed again four steps later as column (x+4)-2: a ring of four descriptors per
    // row keeps it in registers, so a step issues two loads (2 x 4 L1 wavefronts) instead of four.  The trip count is
    // rounded up to a multiple of 4; the extra steps evaluate out-of-range hypotheses (rejected by drel > span) on
    // columns that still lie inside the frame (row v+2 <= H-4, so a few descriptors past its end are the next row).
    const uint4 *pt = ot + wxlo, *pb = ob + wxlo;
    SVB_GUARD_ASSERT(pt - 2 >= oth_lo && pb + 2 <= oth_hi);
    uint4 t0 = __ldg(pt - 2), t1 = __ldg(pt - 1), t2 = __ldg(pt), t3 = __ldg(pt + 1);
    uint4 b0 = __ldg(pb - 2), b1 = __ldg(pb - 1), b2 = __ldg(pb), b3 = __ldg(pb + 1);
#define SVB_MATCH_STEP(TK, BK, OFF)                                                        \
    {                                                                                      \
        SVB_GUARD_ASSERT(pt + 2 + OFF >= oth_lo && pb + 2 + OFF + 1 <= oth_hi);            \
        const uint4 tn = __ldg(pt + 2 + OFF), bn = __ldg(pb + 2 + OFF);                    \
        unsigned e = sad16_acc(a0, TK, 0u);                                                \
        e = sad16_acc(a1, tn, e);                                                          \
        e = sad16_acc(a2, BK, e);                                                          \
        e = sad16_acc(a3, bn, e);                                                          \
        const unsigned key = e * 65536u + drel;                                            \
        if (drel <= span) { /* three predicated min/max instead of a select + three */    \
            second = min(second, max(best, key));                                          \
            best = min(best, key);                                                         \
        }                                                                                  \
        drel += dstep;                                                                     \
        TK = tn;                                                                           \
        BK = bn;                                                                           \
    }
#pragma unroll 2
    for (int n = (wxhi - wxlo + 4) >> 2; n > 0; n--) {
        SVB_MATCH_STEP(t0, b0, 0)
        SVB_MATCH_STEP(t1, b1, 1)
        SVB_MATCH_STEP(t2, b2, 2)
        SVB_MATCH_STEP(t3, b3, 3)
        pt += 4;
        pb += 4;
    }
#undef SVB_MATCH_STEP
    if (!ok) return -1;
    // at least 11 hypotheses were evaluated, so both minima exist (min_1_d >= 0 && min_2_d >= 0)
    const int min1_e = (int)(best >> 16), min1_d = (int)(best & 0xFFFFu) + dmin, min2_e = (int)(second >> 16);
    if ((float)min1_e < __fmul_rn(support_threshold, (float)min2_e)) return min1_d;  // elas.cpp:364
    return -1;
}

// grid: (ceil(#patches / SM_WARPS), nf); patches are numbered along the lattice row first
// COUNT: also add up the evaluated hypotheses (svb_set_eval_counting); the timed kernel is the one without
template <bool COUNT>
__global__ void __launch_bounds__(SM_WARPS * 32) k_support_match(const uint8_t *__restrict__ desc1, const uint8_t *__restrict__ desc2,
                                                                 int16_t *__restrict__ dcan_raw, int W, int H, int cw, int ch, int step,
                                                                 int disp_min, int disp_max, int support_texture, float support_threshold,
                                                                 int lr_threshold, int vc0, int vc1, unsigned long long *evals, size_t pad, int nf) {
    // lattice rows vc0 .. vc1-1 (1 .. ch-1 for a whole frame; a sub-range in the row-band split)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pu = (cw - 1 + PATCH_U - 1) / PATCH_U, pv = (vc1 - vc0 + PATCH_V - 1) / PATCH_V;
    const int patch = blockIdx.x * SM_WARPS + wid;
    if (patch >= pu * pv) return;  // warp-uniform
    const int prow = patch / pu, pcol = patch - prow * pu;
    const int f = blockIdx.y;
    const int uc = 1 + pcol * PATCH_U + (lane & (PATCH_U - 1));
    const int vc = vc0 + prow * PATCH_V + (lane / PATCH_U);
    const bool has = uc < cw && vc < vc1;
    const int u = uc * step, v = vc * step;
    const size_t fo = (size_t)f * W * H;
    const uint4 *d1 = reinterpret_cast<const uint4 *>(desc1) + fo;
    const uint4 *d2 = reinterpret_cast<const uint4 *>(desc2) + fo;

    // forward: candidate in the left image, search the right image (elas.cpp:403)
    unsigned n_hyp = 0u;
    // arenas including their guard bands (range asserts of the guard build only)
    const uint4 *lo1 = reinterpret_cast<const uint4 *>(desc1 - pad), *hi1 = reinterpret_cast<const uint4 *>(desc1 + (size_t)nf * W * H * 16 + pad);
    const uint4 *lo2 = reinterpret_cast<const uint4 *>(desc2 - pad), *hi2 = reinterpret_cast<const uint4 *>(desc2 + (size_t)nf * W * H * 16 + pad);
    const int d = match_pass<false, COUNT>(d1, d2, u, v, has, W, H, disp_min, disp_max, support_texture, support_threshold, n_hyp, lo2, hi2);
    // backward: the match (u-d, v) as a candidate of the right image, search the left image (elas.cpp:406)
    const int dback = match_pass<true, COUNT>(d2, d1, u - d, v, has && d >= 0, W, H, disp_min, disp_max, support_texture, support_threshold, n_hyp, lo1,
                                              hi1);
    if (COUNT) {
        const unsigned tot = __reduce_add_sync(0xFFFFFFFFu, n_hyp);
        if (lane == 0) atomicAdd(evals, (unsigned long long)tot);
    }
    int result = -1;
    if (d >= 0 && dback >= 0 && abs(d - dback) <= lr_threshold) result = d;  // elas.cpp:404-409
    if (has) dcan_raw[(size_t)f * cw * ch + (size_t)vc * cw + uc] = (int16_t)result;
}

// ------------------------------------------------------------------------------------------------
// Matching, one CTA per lattice row (the form every frame up to 3845 pixels wide uses; the patch kernel above remains for wider ones).
//
// The two "other image" rows a lattice row reads (v - 2 and v + 2) are staged in shared memory once, and every lane walks ITS OWN
// matched columns: in step k all lanes of the warp evaluate the SAME disparity d_k (so the disparity is a scalar of the key, and
// a lane's range is the warp's range wherever no border cuts it short) on columns x = u -+ d_k that differ per lane.  Reading
// different columns per lane costs nothing in shared memory -- the 8 lanes of a quarter warp sit 5 columns = 80 bytes apart, which
// spreads a 16-byte access over all 32 banks -- whereas the broadcast form above has to walk the UNION of the lanes' column ranges
// (88 % of its hypotheses are in range) and to test every hypothesis against its lane's range.  Loop body: 2 LDS.128, 16 chained
// VABSDIFF4.ACC, 3 min/max, 2 IMAD (FMA pipe).  Forward and backward pass are two launches of the same kernel with the images'
// roles exchanged; the forward result waits in dcan_raw.
#ifndef SVB_MR_UNROLL_PRAGMA
#define SVB_MR_UNROLL_PRAGMA "unroll 2"  // groups of four steps per loop trip (1 and 4 measured: see DESIGN.md)
#endif
#ifndef SVB_MR_MAX_THREADS
#define SVB_MR_MAX_THREADS 768
#endif
constexpr int MR_MAX_THREADS = SVB_MR_MAX_THREADS;  // candidates per lattice row (one CTA); 80 registers x 768 threads fit the register file

__device__ __forceinline__ unsigned imad_fma_pipe(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// RIGHT = false: candidates (uc * step, v) of image `own` = left, matched columns x = u - d in the right image, result -> dcan_raw
// RIGHT = true : candidates (u - d_forward, v) of image `own` = right, matched columns x = u' + d in the left image; dcan_raw keeps
//                d_forward where the backward match agrees within lr_threshold, else -1 (elas.cpp:394-411)
// grid: (vc1 - vc0, nf); dynamic smem: 2 * row_cols uint4
template <bool RIGHT, bool COUNT>
__global__ void __launch_bounds__(MR_MAX_THREADS) k_support_match_row(const uint8_t *__restrict__ desc_own, const uint8_t *__restrict__ desc_oth,
                                                                       int16_t *__restrict__ dcan_raw, int W, int H, int cw, int ch, int step,
                                                                       int disp_min, int disp_max, int support_texture, float support_threshold,
                                                                       int lr_threshold, int vc0, int padl, int row_cols, unsigned long long *evals) {
    // [2][row_cols]: rows v - 2 and v + 2 of the other image, column x at index padl + x.  padl = 32 * step + 8 columns left of column 0:
    // the lanes beside the one with the largest range start up to 31 * step + 2 columns before column 5 (their hypotheses there are discarded)
    extern __shared__ __align__(128) uint4 s_match_rows[];
    const int lane = threadIdx.x & 31;
    const int vc = vc0 + blockIdx.x, f = blockIdx.y;
    const int v = vc * step;
    int uc = 1 + threadIdx.x;
    bool has = uc < cw;
    const size_t row_cell = (size_t)f * cw * ch + (size_t)vc * cw;
    if (v < 5 || v > H - 6) {  // no candidate of this row is inside the frame (elas.cpp:279)
        if (!RIGHT && has) dcan_raw[row_cell + uc] = -1;
        return;
    }
    // Backward pass: only the candidates the forward pass matched take part (5 of 6, fewer on real frames).  They are packed into the
    // first lanes of the CTA -- neither the shared-memory walk nor its bank alignment cares which candidate a lane holds -- and the
    // warps left without one retire.
    __shared__ int s_count[MR_MAX_THREADS / 32];
    __shared__ int s_packed[MR_MAX_THREADS];  // (uc << 16) | d_forward
    int d_fwd = -1;
    if (RIGHT) {
        d_fwd = has ? dcan_raw[row_cell + uc] : -1;
        const unsigned matched = __ballot_sync(0xFFFFFFFFu, d_fwd >= 0);
        if (lane == 0) s_count[threadIdx.x >> 5] = __popc(matched);
        __syncthreads();
        int before = __popc(matched & ((1u << lane) - 1u));
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) before += s_count[w];
        if (d_fwd >= 0) s_packed[before] = (uc << 16) | d_fwd;
    }
    const size_t fo = (size_t)f * W * H;
    const uint4 *own = reinterpret_cast<const uint4 *>(desc_own) + fo;
    const uint4 *oth = reinterpret_cast<const uint4 *>(desc_oth) + fo;
    for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) {
        const int r = i >= W ? 1 : 0, x = i - r * W;
        s_match_rows[r * row_cols + padl + x] = __ldg(oth + (size_t)(v + (r ? 2 : -2)) * W + x);
    }
    __syncthreads();

    // the candidate's column in its own image
    int u;
    bool ok;
    if (RIGHT) {
        int total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) total += s_count[w];
        if ((int)(threadIdx.x & ~31u) >= total) return;  // whole warps; no barrier follows
        has = (int)threadIdx.x < total;
        const int e = has ? s_packed[threadIdx.x] : 0;
        uc = has ? e >> 16 : 1;
        d_fwd = has ? (e & 0xFFFF) : -1;
        u = uc * step - max(d_fwd, 0);
        ok = has;
    } else {
        u = min(uc, cw - 1) * step;
        ok = has;
    }
    const size_t cell = row_cell + uc;
    // candidate validity and disparity range (elas.cpp:279,296-300,318-327)
    ok = ok && u >= 5 && u <= W - 6;
    if (ok) ok = (int)texture16(__ldg(own + (size_t)v * W + u)) >= support_texture;
    const int dmin = max(disp_min, 0);
    const int dmax = RIGHT ? min(disp_max, W - u - 5) : min(disp_max, u - 5);
    ok = ok && (dmax - dmin >= 10);
    unsigned n_hyp = ok ? (unsigned)(dmax - dmin + 1) : 0u;  // elas.cpp:330: every d of the range is evaluated
    int result = -1;
    const int d_hi = __reduce_max_sync(0xFFFFFFFFu, ok ? dmax : -1);  // the warp walks d = dmin .. d_hi
    if (d_hi >= 0) {
        uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
        if (ok) {
            a0 = __ldg(own + (size_t)(v - 2) * W + u - 2);
            a1 = __ldg(own + (size_t)(v - 2) * W + u + 2);
            a2 = __ldg(own + (size_t)(v + 2) * W + u - 2);
            a3 = __ldg(own + (size_t)(v + 2) * W + u + 2);
        }
        // Columns ascend in both passes.  The left candidates of a warp sit `step` columns apart and walk d = d_hi down to dmin in
        // lock step: with step = 5 the 8 lanes of a quarter warp read 16-byte columns 80 bytes apart -- all 32 banks, no conflict.
        // The right candidates u - d_forward sit anywhere; lane j of a quarter warp therefore starts up to 7 columns early, on a column
        // whose index is j modulo 8 (again one bank group per lane), and carries its own disparity counter.
        const int shift = RIGHT ? ((padl + u + dmin - lane) & 7) : 0;
        const int x0 = RIGHT ? u + dmin - shift : u - d_hi;
        const int n = RIGHT ? __reduce_max_sync(0xFFFFFFFFu, ok ? dmax - dmin + 1 + shift : 0) : d_hi - dmin + 1;
        const uint4 *pt = s_match_rows + padl + x0, *pb = pt + row_cols;
        SVB_GUARD_ASSERT(padl + x0 - 2 >= 0 && padl + x0 + n + 2 <= row_cols);
        unsigned drel = (unsigned)(RIGHT ? -shift : d_hi - dmin);  // d - dmin of the step; a hypothesis counts iff drel <= span
        const unsigned span = (unsigned)(dmax - dmin);
        // constants the compiler cannot see through (blockDim.z is 1): key = e * 65536 + drel stays an IMAD on the FMA pipe, which is
        // idle here, instead of becoming a shift-add on the ALU pipe, which is the limiter
        const unsigned one = blockDim.z, shift16 = 65536u * blockDim.z;
        const unsigned dstep = RIGHT ? 1u : 0xFFFFFFFFu;
        unsigned best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;  // (E << 16) | (d - dmin)
        // ring of four descriptors per row: the descriptor of column x + 2 is needed again four steps later as column (x + 4) - 2
        uint4 t0 = pt[-2], t1 = pt[-1], t2 = pt[0], t3 = pt[1];
        uint4 b0 = pb[-2], b1 = pb[-1], b2 = pb[0], b3 = pb[1];
#define SVB_ROW_STEP(TK, BK, OFF, PRED)                                \
    {                                                                  \
        const uint4 tn = pt[2 + OFF], bn = pb[2 + OFF];                \
        unsigned e = sad16_acc(a0, TK, 0u);                            \
        e = sad16_acc(a1, tn, e);                                      \
        e = sad16_acc(a2, BK, e);                                      \
        e = sad16_acc(a3, bn, e);                                      \
        const unsigned key = imad_fma_pipe(e, shift16, drel);          \
        if (!(PRED) || drel <= span) {                                 \
            second = min(second, max(best, key));                      \
            best = min(best, key);                                     \
        }                                                              \
        drel = imad_fma_pipe(dstep, one, drel);                        \
        TK = tn;                                                       \
        BK = bn;                                                       \
    }
#define SVB_ROW_WALK(PRED)                                 \
    {                                                      \
        _Pragma(SVB_MR_UNROLL_PRAGMA) for (int g = n >> 2; g > 0; g--) { \
            SVB_ROW_STEP(t0, b0, 0, PRED)                  \
            SVB_ROW_STEP(t1, b1, 1, PRED)                  \
            SVB_ROW_STEP(t2, b2, 2, PRED)                  \
            SVB_ROW_STEP(t3, b3, 3, PRED)                  \
            pt += 4;                                       \
            pb += 4;                                       \
        }                                                  \
        const int rem = n & 3;                             \
        if (rem > 0) SVB_ROW_STEP(t0, b0, 0, PRED)         \
        if (rem > 1) SVB_ROW_STEP(t1, b1, 1, PRED)         \
        if (rem > 2) SVB_ROW_STEP(t2, b2, 2, PRED)         \
    }
        // lanes without a candidate compute garbage that is discarded; where every candidate of the warp has the warp's range
        // (no border cuts a lane short) the hypotheses need no test at all
        if (!RIGHT && __all_sync(0xFFFFFFFFu, !ok || dmax == d_hi))
            SVB_ROW_WALK(false)
        else
            SVB_ROW_WALK(true)
#undef SVB_ROW_WALK
#undef SVB_ROW_STEP
        if (ok) {
            // at least 11 hypotheses were evaluated, so both minima exist (min_1_d >= 0 && min_2_d >= 0)
            const int min1_e = (int)(best >> 16), min1_d = (int)(best & 0xFFFFu) + dmin, min2_e = (int)(second >> 16);
            if ((float)min1_e < __fmul_rn(support_threshold, (float)min2_e)) result = min1_d;  // elas.cpp:364
        }
    }
    if (COUNT) {
        const unsigned tot = __reduce_add_sync(0xFFFFFFFFu, n_hyp);
        if (lane == 0 && tot) atomicAdd(evals, (unsigned long long)tot);
    }
    if (RIGHT) result = (d_fwd >= 0 && result >= 0 && abs(d_fwd - result) <= lr_threshold) ? d_fwd : -1;  // elas.cpp:404-409
    if (has) dcan_raw[cell] = (int16_t)result;
}

// Row 0 / column 0 of D_can keep the calloc zero (elas.cpp:387): a valid disparity-0 neighbour for the filters.
__global__ void k_dcan_border(int16_t *__restrict__ dcan_raw, int cw, int ch) {
    int16_t *d = dcan_raw + (size_t)blockIdx.y * cw * ch;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cw) d[i] = 0;
    if (i < ch) d[(size_t)i * cw] = 0;
}

// ------------------------------------------------------------------------------------------------
// The three in-place lattice filters with the reference's sequential (column-major) semantics, then the compaction.
// k_incon_first_sweep (all cells of all frames at once) does the first and most expensive sweep of the first filter;
// one CTA per frame (k_support_filter) continues from its flags.
//
// removeInconsistentSupportPoints (elas.cpp:152-176) is a sweep in which a cell sees the FINAL state of
// the cells that precede it in column-major order and the ORIGINAL state of those that follow.  Let R* be
// the set it removes.  R* is the unique fixpoint of
//     R = { c valid : #{ w in window(c) : valid(w), |d_c-d_w|<=thr, not (w in R and w precedes c) } < min_support }
// (unique because membership of c only depends on cells preceding c).  Starting from R = {} and marking cells
// whenever their count with the CURRENT R drops below the bound never marks a cell outside R* (counts only
// shrink as R grows towards R*), and a sweep without change means R is a fixpoint, hence R = R*.  Every sweep
// is embarrassingly parallel; racing reads of `removed` are benign because marks are monotone.
//
// removeRedundantSupportPoints (elas.cpp:178-233): the vertical pass only looks along a column and the
// horizontal pass only along a row, so columns (rows) are independent and one thread walks each in order.
// ------------------------------------------------------------------------------------------------
constexpr int SF_THREADS = 1024;
constexpr int SF_REMOVED = 0x4000;  // flag bit of a lattice cell (disparities are < 4096)
constexpr int SF_MARK_A = 0x2000, SF_MARK_B = 0x1000;  // "re-examine in this / the next sweep" (alternating roles)
constexpr int SF_VALUE = 0x0FFF;

__device__ __forceinline__ bool precedes_colmajor(int u2, int v2, int u, int v) { return u2 < u || (u2 == u && v2 < v); }

// One line (column or row) of removeRedundantSupportPoints with redun_max_dist 5, redun_threshold 1
// (elas.cpp:419-420): sequential along the line, looking at the UPDATED state behind and the original state ahead.
// The 11-cell neighbourhood slides through registers, so a step costs one load.
__device__ __forceinline__ void redundant_line(int16_t *work, int base, int stride, int len) {
    int win[11];
#pragma unroll
    for (int k = 0; k < 11; k++) {
        const int pos = k - 5;
        win[k] = (pos >= 0 && pos < len) ? (int)work[base + pos * stride] : -1;
    }
    for (int i = 0; i < len; i++) {
        const int dc = win[5];
        if (dc >= 0) {
            bool behind = false, ahead = false;
#pragma unroll
            for (int k = 0; k < 5; k++) behind |= (win[k] >= 0 && abs(dc - win[k]) <= 1);
#pragma unroll
            for (int k = 6; k < 11; k++) ahead |= (win[k] >= 0 && abs(dc - win[k]) <= 1);
            if (behind && ahead) {
                win[5] = -1;
                work[base + i * stride] = (int16_t)-1;
            }
        }
#pragma unroll
        for (int k = 0; k < 10; k++) win[k] = win[k + 1];
        const int nxt = i + 6;
        win[10] = nxt < len ? (int)work[base + nxt * stride] : -1;
    }
}

// OR / AND-NOT of flag bits into one 16-bit lattice cell through a 32-bit atomic on the word that holds it (shared or
// global memory; the other half of the word belongs to a neighbouring cell and is left alone).
__device__ __forceinline__ void cell_set(int16_t *cell, int bits) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(cell);
    atomicOr(reinterpret_cast<unsigned *>(a & ~(uintptr_t)3), (unsigned)bits << ((a & 2) * 8));
}
__device__ __forceinline__ void cell_clear(int16_t *cell, int bits) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(cell);
    atomicAnd(reinterpret_cast<unsigned *>(a & ~(uintptr_t)3), ~((unsigned)bits << ((a & 2) * 8)));
}

// First sweep of removeInconsistentSupportPoints for every cell of every frame at once (the fixpoint iteration of
// k_support_filter starts from R = {}: this IS its first sweep, and by far the most expensive one -- every valid cell
// scans its window until it has found its supporters).  dcan = raw, plus SF_REMOVED where the count falls short.
// grid: (ceil(cells / 256), nf)
__global__ void __launch_bounds__(256) k_incon_first_sweep(const int16_t *__restrict__ dcan_raw_all, int16_t *__restrict__ dcan_all, int cw, int ch,
                                                          int incon_window, int incon_threshold, int incon_min_support) {
    const int cells = cw * ch;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    const int16_t *raw = dcan_raw_all + (size_t)blockIdx.y * (unsigned)cells;
    int e = raw[i];
    if (e >= 0) {
        const int v = i / cw, u = i - v * cw;
        const int u_lo = max(u - incon_window, 0), u_hi = min(u + incon_window, cw - 1);
        const int v_lo = max(v - incon_window, 0), v_hi = min(v + incon_window, ch - 1);
        int support_cnt = 0;
        for (int v2 = v_lo; v2 <= v_hi && support_cnt < incon_min_support; v2++)
            for (int u2 = u_lo; u2 <= u_hi; u2++) {
                const int e2 = raw[v2 * cw + u2];
                support_cnt += (e2 >= 0 && abs(e - e2) <= incon_threshold) ? 1 : 0;
            }
        if (support_cnt < incon_min_support) e |= SF_REMOVED;
    }
    dcan_all[(size_t)blockIdx.y * (unsigned)cells + i] = (int16_t)e;
}

// SMEM = true: the lattice lives in shared memory (2 bytes per cell; every frame size up to about 1080p at step 5);
// SMEM = false: it is worked on in place in the global dcan array (4K frames).
// phases: which parts of the filter chain this launch runs.  Small lattices do everything in one launch; large ones (4K frames) run
// the sweeps and the redundant-point passes as multi-CTA kernels (below) and use this kernel for what is left:
//   SF_PH_MARK | SF_PH_SWEEP  inconsistent points to the fixpoint (SF_PH_MARK: first mark the windows of the cells the first sweep removed;
//                             without it the marks of bit `first_bit` are already in place -- the multi-CTA sweeps were cut off before
//                             the fixpoint, `unconverged` != 0)
//   SF_PH_FINAL               flags -> final lattice values;   SF_PH_REDUNDANT  vertical, then horizontal redundant-point pass
//   SF_PH_COMPACT             ordered compaction, corner points, host copy
constexpr int SF_PH_MARK = 1, SF_PH_SWEEP = 2, SF_PH_FINAL = 4, SF_PH_REDUNDANT = 8, SF_PH_COMPACT = 16, SF_PH_ALL = 31;

template <bool SMEM>
__global__ void __launch_bounds__(SF_THREADS) k_support_filter(int16_t *__restrict__ dcan_all,
                                                               int32_t *__restrict__ support_all, int32_t *__restrict__ nsupport_all,
                                                               int32_t *__restrict__ h_support_all, int32_t *__restrict__ h_nsupport_all, int W,
                                                               int H, int cw, int ch, int step, int incon_window, int incon_threshold,
                                                               int incon_min_support, int add_corners, int maxS, int phases, int first_bit,
                                                               const int32_t *__restrict__ unconverged, int unconverged_stride) {
    extern __shared__ int16_t s_lattice[];
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    const int cells = cw * ch;
    int16_t *dcan = dcan_all + (size_t)f * cells;
    int16_t *work = SMEM ? s_lattice : dcan;
    int32_t *support = support_all + (size_t)f * maxS * 3;

    __shared__ int s_changed;
    __shared__ int s_warp_tot[SF_THREADS / 32];
    __shared__ int s_base;
    __shared__ unsigned long long s_best[4];
    // the last multi-CTA sweep removed nothing: the fixpoint is reached, nothing to do
    if (phases == SF_PH_SWEEP && unconverged && unconverged[f * unconverged_stride] == 0) return;

    // dcan already holds the result of the first sweep (k_incon_first_sweep): raw values, SF_REMOVED where it struck
    if (SMEM) {
        for (int i = tid; i < cells; i += SF_THREADS) work[i] = dcan[i];
        __syncthreads();
    }

    // ---- inconsistent points: parallel sweeps to the fixpoint; a removed cell keeps its value and gets SF_REMOVED ----
    // A cell's count can change only when a cell of its window is removed, so a removal marks its window for the NEXT
    // sweep (two alternating mark bits in the cell itself, set with 32-bit atomics on the word that holds the cell).
    // The removals of the first sweep mark their windows here.
    // in global memory the flag atomics act in L2: the cells are read around L1 for as long as flags are in play
    auto cell = [&](int idx) -> int { return SMEM ? (int)work[idx] : (int)__ldcg(work + idx); };
    int cur_bit = first_bit, next_bit = first_bit == SF_MARK_A ? SF_MARK_B : SF_MARK_A;
    if (phases & SF_PH_MARK)
        for (int i = tid; i < cells; i += SF_THREADS) {
            const int e = cell(i);
            if (e < 0 || !(e & SF_REMOVED)) continue;
            const int v = i / cw, u = i - v * cw;
            for (int v2 = max(v - incon_window, 0); v2 <= min(v + incon_window, ch - 1); v2++)
                for (int u2 = max(u - incon_window, 0); u2 <= min(u + incon_window, cw - 1); u2++)
                    if (cell(v2 * cw + u2) >= 0) cell_set(work + v2 * cw + u2, cur_bit);
        }
    __syncthreads();
    while (phases & SF_PH_SWEEP) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < cells; i += SF_THREADS) {
            const int e = cell(i);
            if (e < 0 || (e & SF_REMOVED) || !(e & cur_bit)) continue;
            cell_clear(work + i, cur_bit);
            const int dc = e & SF_VALUE;
            const int v = i / cw, u = i - v * cw;
            const int u_lo = max(u - incon_window, 0), u_hi = min(u + incon_window, cw - 1);
            const int v_lo = max(v - incon_window, 0), v_hi = min(v + incon_window, ch - 1);
            int support_cnt = 0;
            for (int v2 = v_lo; v2 <= v_hi && support_cnt < incon_min_support; v2++)
                for (int u2 = u_lo; u2 <= u_hi; u2++) {
                    const int e2 = cell(v2 * cw + u2);
                    if (e2 < 0) continue;
                    if (abs(dc - (e2 & SF_VALUE)) > incon_threshold) continue;
                    if ((e2 & SF_REMOVED) && precedes_colmajor(u2, v2, u, v)) continue;
                    support_cnt++;
                }
            if (support_cnt < incon_min_support) {
                s_changed = 1;
                cell_set(work + i, SF_REMOVED);
                for (int v2 = v_lo; v2 <= v_hi; v2++)
                    for (int u2 = u_lo; u2 <= u_hi; u2++)
                        if (cell(v2 * cw + u2) >= 0) cell_set(work + v2 * cw + u2, next_bit);
            }
        }
        __syncthreads();
        const int again = s_changed;
        __syncthreads();
        if (!again) break;
        const int t = cur_bit;
        cur_bit = next_bit;
        next_bit = t;
    }
    if (phases & SF_PH_FINAL)
        for (int i = tid; i < cells; i += SF_THREADS) {
            const int e = cell(i);
            if (e >= 0) work[i] = (e & SF_REMOVED) ? (int16_t)-1 : (int16_t)(e & SF_VALUE);
        }
    __syncthreads();

    // ---- redundant points: vertical pass (columns independent), then horizontal pass (rows independent) ----
    if (phases & SF_PH_REDUNDANT) {
        for (int u = tid; u < cw; u += SF_THREADS) redundant_line(work, u, cw, ch);
        __syncthreads();
        for (int v = tid; v < ch; v += SF_THREADS) redundant_line(work, v * cw, 1, cw);
        __syncthreads();
    }
    if (SMEM)
        for (int i = tid; i < cells; i += SF_THREADS) dcan[i] = work[i];
    if (!(phases & SF_PH_COMPACT)) return;

    // ---- ordered compaction: u_can outer, v_can inner, both from 1 (elas.cpp:424-428) ----
    // Column-wise: a thread counts the surviving cells of its lattice column, one block-wide exclusive scan over the columns gives
    // every column its place in the list, and the thread writes its column's points in v order.  (A scan over the flattened cells
    // needs cells / 1024 block-wide steps: 324 of them at 4K.)
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5;
    for (int u0 = 1; u0 < cw; u0 += SF_THREADS) {
        const int uc = u0 + tid;
        int cnt = 0;
        if (uc < cw)
            for (int vc = 1; vc < ch; vc++) cnt += work[vc * cw + uc] >= 0 ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp_tot[wid] = incl;
        __syncthreads();
        int off = s_base + incl - cnt;
        for (int w = 0; w < wid; w++) off += s_warp_tot[w];
        if (uc < cw && cnt > 0)
            for (int vc = 1; vc < ch; vc++) {
                const int dd = work[vc * cw + uc];
                if (dd >= 0) {
                    support[3 * off + 0] = uc * step;
                    support[3 * off + 1] = vc * step;
                    support[3 * off + 2] = dd;
                    off++;
                }
            }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < SF_THREADS / 32; w++) t += s_warp_tot[w];
            s_base += t;
        }
        __syncthreads();
    }
    int n = s_base;

    // ---- corner points (elas.cpp:235-264) ----
    if (add_corners) {
        const int bu[4] = {0, 0, W - 1, W - 1};
        const int bv[4] = {0, H - 1, 0, H - 1};
        if (tid < 4) s_best[tid] = 0xFFFFFFFFFFFFFFFFull;
        __syncthreads();
        unsigned long long loc[4] = {~0ull, ~0ull, ~0ull, ~0ull};
        for (int j = tid; j < n; j += SF_THREADS) {
            const int su = support[3 * j], sv = support[3 * j + 1];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int du = bu[k] - su, dv = bv[k] - sv;
                const unsigned long long dist = (unsigned long long)(du * du + dv * dv);
                if (dist < 10000000ull) {  // best_dist starts at 10000000 with a strict compare
                    const unsigned long long key = (dist << 32) | (unsigned)j;  // first in list order wins ties
                    if (key < loc[k]) loc[k] = key;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (loc[k] != ~0ull) atomicMin(&s_best[k], loc[k]);
        __syncthreads();
        if (tid == 0) {
            int bd[4];
            for (int k = 0; k < 4; k++) bd[k] = (s_best[k] == ~0ull) ? 0 : support[3 * (int)(s_best[k] & 0xFFFFFFFFull) + 2];
            for (int k = 0; k < 4; k++) {
                support[3 * (n + k) + 0] = bu[k];
                support[3 * (n + k) + 1] = bv[k];
                support[3 * (n + k) + 2] = bd[k];
            }
            support[3 * (n + 4) + 0] = bu[2] + bd[2];
            support[3 * (n + 4) + 1] = bv[2];
            support[3 * (n + 4) + 2] = bd[2];
            support[3 * (n + 5) + 0] = bu[3] + bd[3];
            support[3 * (n + 5) + 1] = bv[3];
            support[3 * (n + 5) + 2] = bd[3];
        }
        n += 6;
    }
    if (tid == 0) nsupport_all[f] = n;
    // The host stage (Delaunay) needs the list: it is written straight into mapped pinned host memory, so only the
    // n points that exist cross PCIe and no device-to-host copy has to be queued behind this kernel.
    if (h_support_all) {
        __syncthreads();
        int32_t *hs = h_support_all + (size_t)f * maxS * 3;
        for (int i = tid; i < 3 * n; i += SF_THREADS) hs[i] = support[i];
        if (tid == 0) h_nsupport_all[f] = n;
    }
}


// ---- multi-CTA form of the lattice filters for lattices that do not fit one CTA's shared memory (4K frames: 768 x 432 cells) ----
// The same fixpoint iteration as in k_support_filter, one kernel launch per sweep (the launch boundary is the grid-wide barrier), on
// the lattice in global memory; flags are read around L1 (__ldcg) because the flag atomics act in L2.
// grid: (ceil(cells / 256), nf)
__global__ void __launch_bounds__(256) k_incon_mark(int16_t *__restrict__ dcan_all, int cw, int ch, int incon_window) {
    const int cells = cw * ch;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    int16_t *work = dcan_all + (size_t)blockIdx.y * (unsigned)cells;
    const int e = __ldcg(work + i);
    if (e < 0 || !(e & SF_REMOVED)) return;
    const int v = i / cw, u = i - v * cw;
    for (int v2 = max(v - incon_window, 0); v2 <= min(v + incon_window, ch - 1); v2++)
        for (int u2 = max(u - incon_window, 0); u2 <= min(u + incon_window, cw - 1); u2++)
            if (__ldcg(work + v2 * cw + u2) >= 0) cell_set(work + v2 * cw + u2, SF_MARK_A);
}

// One sweep: cells marked `cur_bit` are re-examined; changed[f * stride + it] = 1 if the sweep removed something (then the next one
// has work).  A sweep whose predecessor changed nothing returns at once.
__global__ void __launch_bounds__(256) k_incon_sweep(int16_t *__restrict__ dcan_all, int32_t *__restrict__ changed, int stride, int it, int cw, int ch,
                                                    int incon_window, int incon_threshold, int incon_min_support) {
    const int f = blockIdx.y;
    if (it > 0 && changed[f * stride + it - 1] == 0) return;
    const int cells = cw * ch;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells) return;
    int16_t *work = dcan_all + (size_t)f * (unsigned)cells;
    const int cur_bit = (it & 1) ? SF_MARK_B : SF_MARK_A, next_bit = (it & 1) ? SF_MARK_A : SF_MARK_B;
    const int e = __ldcg(work + i);
    if (e < 0 || (e & SF_REMOVED) || !(e & cur_bit)) return;
    cell_clear(work + i, cur_bit);
    const int dc = e & SF_VALUE;
    const int v = i / cw, u = i - v * cw;
    const int u_lo = max(u - incon_window, 0), u_hi = min(u + incon_window, cw - 1);
    const int v_lo = max(v - incon_window, 0), v_hi = min(v + incon_window, ch - 1);
    int support_cnt = 0;
    for (int v2 = v_lo; v2 <= v_hi && support_cnt < incon_min_support; v2++)
        for (int u2 = u_lo; u2 <= u_hi; u2++) {
            const int e2 = __ldcg(work + v2 * cw + u2);
            if (e2 < 0) continue;
            if (abs(dc - (e2 & SF_VALUE)) > incon_threshold) continue;
            if ((e2 & SF_REMOVED) && precedes_colmajor(u2, v2, u, v)) continue;
            support_cnt++;
        }
    if (support_cnt < incon_min_support) {
        changed[f * stride + it] = 1;
        cell_set(work + i, SF_REMOVED);
        for (int v2 = v_lo; v2 <= v_hi; v2++)
            for (int u2 = u_lo; u2 <= u_hi; u2++)
                if (__ldcg(work + v2 * cw + u2) >= 0) cell_set(work + v2 * cw + u2, next_bit);
    }
}

// Final lattice values + vertical redundant-point pass: a CTA stages RV_COLS lattice columns in shared memory (a walk along a column
// in global memory pays an L2 round trip per step: 432 dependent loads per column at 4K), one thread per column.
// grid: (ceil(cw / RV_COLS), nf), RV_COLS threads; dynamic smem: ch * RV_COLS int16
constexpr int RV_COLS = 32;
__global__ void __launch_bounds__(RV_COLS) k_lattice_vertical(int16_t *__restrict__ dcan_all, int cw, int ch) {
    extern __shared__ int16_t s_strip[];
    int16_t *dcan = dcan_all + (size_t)blockIdx.y * (unsigned)(cw * ch);
    const int u = blockIdx.x * RV_COLS + threadIdx.x;
    const bool in = u < cw;
    for (int v = 0; v < ch; v++) {
        int e = in ? (int)dcan[v * cw + u] : -1;
        if (e >= 0) e = (e & SF_REMOVED) ? -1 : (e & SF_VALUE);  // the inconsistent-point filter's verdict
        s_strip[v * RV_COLS + threadIdx.x] = (int16_t)e;
    }
    if (!in) return;
    redundant_line(s_strip, threadIdx.x, RV_COLS, ch);
    for (int v = 0; v < ch; v++) dcan[v * cw + u] = s_strip[v * RV_COLS + threadIdx.x];
}

// Horizontal redundant-point pass: RH_ROWS lattice rows per CTA in shared memory, one thread per row; the row stride is an odd number
// of 32-bit words so that the 32 threads, which all stand at the same column, hit different banks.
// grid: (ceil(ch / RH_ROWS), nf), RH_ROWS threads; dynamic smem: RH_ROWS * stride int16
constexpr int RH_ROWS = 16;
__global__ void __launch_bounds__(RH_ROWS) k_lattice_horizontal(int16_t *__restrict__ dcan_all, int cw, int ch, int stride) {
    extern __shared__ int16_t s_rows[];
    int16_t *dcan = dcan_all + (size_t)blockIdx.y * (unsigned)(cw * ch);
    const int v0 = blockIdx.x * RH_ROWS;
    const int rows = min(RH_ROWS, ch - v0);
    for (int r = 0; r < rows; r++)
        for (int u = threadIdx.x; u < cw; u += RH_ROWS) s_rows[r * stride + u] = dcan[(v0 + r) * cw + u];
    __syncthreads();
    if ((int)threadIdx.x < rows) redundant_line(s_rows, threadIdx.x * stride, 1, cw);
    __syncthreads();
    for (int r = 0; r < rows; r++)
        for (int u = threadIdx.x; u < cw; u += RH_ROWS) dcan[(v0 + r) * cw + u] = s_rows[r * stride + u];
}

}  // namespace

int launch_dcan_border(const Dims &d, int16_t *dcan_raw, int nf, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    int m = d.cw > d.ch ? d.cw : d.ch;
    dim3 grid((m + 127) / 128, nf);
    k_dcan_border<<<grid, 128, 0, s>>>(dcan_raw, d.cw, d.ch);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_support_match_rows(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int nf, int vc0,
                              int vc1, cudaStream_t s) {
    if (nf <= 0 || d.cw < 2 || vc1 <= vc0) return SVB_OK;
    // one CTA per lattice row wherever a row's candidates fit one CTA and its two descriptor rows fit shared memory (SVB_MATCH_ROWS=0:
    // the patch kernel everywhere)
    const char *rows_env = getenv("SVB_MATCH_ROWS");  // read per launch: the determinism stress test switches it between contexts
    const bool rows_off = rows_env && atoi(rows_env) == 0;
    const int padl = 32 * d.step + 8;
    const int row_cols = (padl + d.W + max(p.disp_max, 0) + 24 + 7) & ~7;  // a multiple of 8 columns: both rows start on the same bank
    const size_t smem = (size_t)2 * row_cols * sizeof(uint4);
    const int threads = ((d.cw - 1 + 31) / 32) * 32;
    if (!rows_off && threads <= MR_MAX_THREADS && smem <= 200 * 1024) {
        static size_t configured[64][4] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        const dim3 grid(vc1 - vc0, nf);
        for (int pass = 0; pass < 2; pass++) {
            const int which = pass * 2 + (d.evals ? 1 : 0);
            const void *fn = which == 0   ? (const void *)k_support_match_row<false, false>
                             : which == 1 ? (const void *)k_support_match_row<false, true>
                             : which == 2 ? (const void *)k_support_match_row<true, false>
                                          : (const void *)k_support_match_row<true, true>;
            if (smem > 48 * 1024 && dev >= 0 && dev < 64 && configured[dev][which] < smem) {
                cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) {
                    set_error("cudaFuncSetAttribute(k_support_match_row, %zu): %s", smem, cudaGetErrorString(e));
                    return SVB_ERR_CUDA;
                }
                configured[dev][which] = smem;
            }
            const uint8_t *own = pass ? desc2 : desc1, *oth = pass ? desc1 : desc2;
            if (which == 0)
                k_support_match_row<false, false><<<grid, threads, smem, s>>>(own, oth, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                                              p.support_texture, p.support_threshold, p.lr_threshold, vc0, padl, row_cols, nullptr);
            else if (which == 1)
                k_support_match_row<false, true><<<grid, threads, smem, s>>>(own, oth, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                                             p.support_texture, p.support_threshold, p.lr_threshold, vc0, padl, row_cols, d.evals);
            else if (which == 2)
                k_support_match_row<true, false><<<grid, threads, smem, s>>>(own, oth, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                                             p.support_texture, p.support_threshold, p.lr_threshold, vc0, padl, row_cols, nullptr);
            else
                k_support_match_row<true, true><<<grid, threads, smem, s>>>(own, oth, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                                            p.support_texture, p.support_threshold, p.lr_threshold, vc0, padl, row_cols, d.evals);
            SVB_LAUNCH_CHECK();
        }
        return SVB_OK;
    }
    const int patches = ((d.cw - 1 + PATCH_U - 1) / PATCH_U) * ((vc1 - vc0 + PATCH_V - 1) / PATCH_V);
    dim3 grid((patches + SM_WARPS - 1) / SM_WARPS, nf);
    if (d.evals)
        k_support_match<true><<<grid, SM_WARPS * 32, 0, s>>>(desc1, desc2, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                             p.support_texture, p.support_threshold, p.lr_threshold, vc0, vc1, d.evals, d.desc_pad, nf);
    else
        k_support_match<false><<<grid, SM_WARPS * 32, 0, s>>>(desc1, desc2, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                              p.support_texture, p.support_threshold, p.lr_threshold, vc0, vc1, nullptr, d.desc_pad, nf);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_support_match(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int nf,
                         cudaStream_t s) {
    SVB_TRY(launch_dcan_border(d, dcan_raw, nf, s));
    return launch_support_match_rows(d, p, desc1, desc2, dcan_raw, nf, 1, d.ch, s);
}

int launch_support_filter(const Dims &d, const svb_params &p, const int16_t *dcan_raw, int16_t *dcan, int32_t *support, int32_t *nsupport,
                          int32_t *h_support, int32_t *h_nsupport, int32_t *changed, int nf, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    if (p.disp_max > SF_VALUE) {
        set_error("support filter: disp_max %d too large for the lattice cell encoding", p.disp_max);
        return SVB_ERR_UNSUPPORTED;
    }
    // first sweep of the inconsistent-point filter for all cells of all frames at once (dcan = raw + SF_REMOVED flags); the
    // one-CTA-per-frame kernel below continues from there
    k_incon_first_sweep<<<dim3((d.cw * d.ch + 255) / 256, nf), 256, 0, s>>>(dcan_raw, dcan, d.cw, d.ch, p.incon_window_size, p.incon_threshold,
                                                                             p.incon_min_support);
    SVB_LAUNCH_CHECK();
    const size_t smem = ((size_t)d.cw * d.ch * sizeof(int16_t) + 3) & ~(size_t)3;  // whole 32-bit words: the sweep marks cells with word atomics
    // SVB_SF_MULTI=1 sends every lattice down the multi-CTA path, SVB_SF_SWEEPS=n (1 .. 8) cuts its sweeps short so that the one-CTA
    // finisher has work: what the tests use to exercise both on small frames
    static const bool force_multi = getenv("SVB_SF_MULTI") && atoi(getenv("SVB_SF_MULTI")) != 0;
    if (smem <= 200 * 1024 && !force_multi) {
        static bool configured[64] = {};  // per device: opt in to large dynamic shared memory once, not on every launch
        int dev = 0;
        cudaGetDevice(&dev);
        if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !configured[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_support_filter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(k_support_filter): %s", cudaGetErrorString(e));
                return SVB_ERR_CUDA;
            }
            configured[dev] = true;
        }
        k_support_filter<true><<<nf, SF_THREADS, smem, s>>>(dcan, support, nsupport, h_support, h_nsupport, d.W, d.H, d.cw, d.ch, d.step, p.incon_window_size,
                                                            p.incon_threshold, p.incon_min_support, p.add_corners, d.maxS, SF_PH_ALL, SF_MARK_A, nullptr, 0);
        SVB_LAUNCH_CHECK();
        return SVB_OK;
    }
    // ---- large lattices (4K): sweeps and redundant-point passes as multi-CTA kernels, k_support_filter for what is left ----
    if (!changed) {
        set_error("support filter: a lattice of %d x %d cells needs the sweep scratch buffer", d.cw, d.ch);
        return SVB_ERR_ARG;
    }
    const int cells = d.cw * d.ch;
    const dim3 cgrid((cells + 255) / 256, nf);
    constexpr int STRIDE = 8;  // per frame: one "this sweep removed something" flag per sweep
    static_assert(STRIDE <= SUPPORT_FILTER_SCRATCH_INTS, "scratch size");
    static const int sweeps_env = getenv("SVB_SF_SWEEPS") ? atoi(getenv("SVB_SF_SWEEPS")) : STRIDE;
    const int SWEEPS = sweeps_env < 1 ? 1 : (sweeps_env > STRIDE ? STRIDE : sweeps_env);
    SVB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int32_t) * STRIDE * (size_t)nf, s));
    k_incon_mark<<<cgrid, 256, 0, s>>>(dcan, d.cw, d.ch, p.incon_window_size);
    SVB_LAUNCH_CHECK();
    for (int it = 0; it < SWEEPS; it++) {
        k_incon_sweep<<<cgrid, 256, 0, s>>>(dcan, changed, STRIDE, it, d.cw, d.ch, p.incon_window_size, p.incon_threshold, p.incon_min_support);
        SVB_LAUNCH_CHECK();
    }
    // not at the fixpoint after SWEEPS sweeps (the last one still removed cells; marks of bit SWEEPS & 1 are pending): one CTA per
    // frame finishes the iteration
    k_support_filter<false><<<nf, SF_THREADS, 0, s>>>(dcan, support, nsupport, h_support, h_nsupport, d.W, d.H, d.cw, d.ch, d.step, p.incon_window_size,
                                                      p.incon_threshold, p.incon_min_support, p.add_corners, d.maxS, SF_PH_SWEEP,
                                                      (SWEEPS & 1) ? SF_MARK_B : SF_MARK_A, changed + SWEEPS - 1, STRIDE);
    SVB_LAUNCH_CHECK();
    {
        const size_t sm = (size_t)d.ch * RV_COLS * sizeof(int16_t);
        if (sm > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k_lattice_vertical, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_lattice_vertical<<<dim3((d.cw + RV_COLS - 1) / RV_COLS, nf), RV_COLS, sm, s>>>(dcan, d.cw, d.ch);
        SVB_LAUNCH_CHECK();
    }
    {
        int stride = (d.cw + 1) & ~1;           // int16 units, even
        if (((stride / 2) & 1) == 0) stride += 2;  // an odd number of 32-bit words
        const size_t sm = (size_t)RH_ROWS * stride * sizeof(int16_t);
        if (sm > 48 * 1024) SVB_CUDA(cudaFuncSetAttribute(k_lattice_horizontal, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_lattice_horizontal<<<dim3((d.ch + RH_ROWS - 1) / RH_ROWS, nf), RH_ROWS, sm, s>>>(dcan, d.cw, d.ch, stride);
        SVB_LAUNCH_CHECK();
    }
    k_support_filter<false><<<nf, SF_THREADS, 0, s>>>(dcan, support, nsupport, h_support, h_nsupport, d.W, d.H, d.cw, d.ch, d.step, p.incon_window_size,
                                                      p.incon_threshold, p.incon_min_support, p.add_corners, d.maxS, SF_PH_COMPACT, SF_MARK_A, nullptr, 0);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
