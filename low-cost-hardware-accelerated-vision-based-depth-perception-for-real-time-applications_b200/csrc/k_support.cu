// Support-point matching on the candidate lattice, the three order-dependent lattice filters, ordered
// compaction into the support list and the optional corner points.
//
// Replaces Elas::computeMatchingDisparity / computeSupportMatches / removeInconsistentSupportPoints /
// removeRedundantSupportPoints / addCornerSupportPoints (src/serial_includes/elas/elas.cpp:152-440).
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

// One warp evaluates every disparity hypothesis of one candidate (elas.cpp:266-371).
// `own` is the descriptor image the candidate lives in, `oth` the image that is searched.
__device__ int match_candidate(const uint4 *__restrict__ own, const uint4 *__restrict__ oth, int u, int v, bool right_image, int W, int H,
                               int disp_min, int disp_max, int support_texture, float support_threshold, int lane) {
    // window_size 3 + step 2 (elas.cpp:279)
    if (!(u >= 5 && u <= W - 6 && v >= 5 && v <= H - 6)) return -1;
    const size_t rowc = (size_t)v * W;
    const uint4 c = __ldg(own + rowc + u);
    if ((int)texture16(c) < support_texture) return -1;

    int dmin = max(disp_min, 0);
    int dmax = right_image ? min(disp_max, W - u - 5) : min(disp_max, u - 5);
    if (dmax - dmin < 10) return -1;  // elas.cpp:326

    const size_t rowt = (size_t)(v - 2) * W;
    const size_t rowb = (size_t)(v + 2) * W;
    const uint4 a1 = __ldg(own + rowt + u - 2);
    const uint4 a2 = __ldg(own + rowt + u + 2);
    const uint4 a3 = __ldg(own + rowb + u - 2);
    const uint4 a4 = __ldg(own + rowb + u + 2);

    // per-lane running best (energy, lowest d) and second-smallest energy of the multiset (elas.cpp:352-360)
    uint32_t best = 0xFFFFFFFFu;  // (E << 16) | d
    uint32_t m1 = 32767u, m2 = 32767u;
    for (int dd = dmin + lane; dd <= dmax; dd += 32) {
        int uw = right_image ? u + dd : u - dd;
        uint32_t e = sad16(a1, __ldg(oth + rowt + uw - 2)) + sad16(a2, __ldg(oth + rowt + uw + 2)) + sad16(a3, __ldg(oth + rowb + uw - 2)) +
                     sad16(a4, __ldg(oth + rowb + uw + 2));
        uint32_t key = (e << 16) | (uint32_t)dd;
        best = min(best, key);
        if (e < m1) {
            m2 = m1;
            m1 = e;
        } else if (e < m2) {
            m2 = e;
        }
    }
    // warp argmin with lowest-d tie break; second smallest of the union of the per-lane pairs
    best = __reduce_min_sync(0xFFFFFFFFu, best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint32_t o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, o);
        uint32_t o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, o);
        uint32_t hi = max(m1, o1);
        m1 = min(m1, o1);
        m2 = min(hi, min(m2, o2));
    }
    const int min1_e = (int)(best >> 16);
    const int min1_d = (int)(best & 0xFFFFu);
    // at least 11 hypotheses were evaluated, so both minima exist (min_1_d >= 0 && min_2_d >= 0)
    if ((float)min1_e < __fmul_rn(support_threshold, (float)(int)m2)) return min1_d;
    return -1;
}

constexpr int SM_WARPS = 8;

__global__ void __launch_bounds__(SM_WARPS * 32) k_support_match(const uint8_t *__restrict__ desc1, const uint8_t *__restrict__ desc2,
                                                                 int16_t *__restrict__ dcan_raw, int W, int H, int cw, int ch, int step,
                                                                 int disp_min, int disp_max, int support_texture, float support_threshold,
                                                                 int lr_threshold) {
    const int lane = threadIdx.x & 31;
    const int cells = (cw - 1) * (ch - 1);
    const int cell = blockIdx.x * SM_WARPS + (threadIdx.x >> 5);
    if (cell >= cells) return;
    const int f = blockIdx.y;
    // warp order follows image rows so that neighbouring warps share descriptor rows in L1/L2
    const int vc = 1 + cell / (cw - 1);
    const int uc = 1 + cell - (vc - 1) * (cw - 1);
    const size_t fo = (size_t)f * W * H;
    const uint4 *d1 = reinterpret_cast<const uint4 *>(desc1) + fo;
    const uint4 *d2 = reinterpret_cast<const uint4 *>(desc2) + fo;
    const int u = uc * step, v = vc * step;

    int result = -1;
    int d = match_candidate(d1, d2, u, v, false, W, H, disp_min, disp_max, support_texture, support_threshold, lane);
    if (d >= 0) {
        int d2nd = match_candidate(d2, d1, u - d, v, true, W, H, disp_min, disp_max, support_texture, support_threshold, lane);
        if (d2nd >= 0 && abs(d - d2nd) <= lr_threshold) result = d;  // elas.cpp:404-409
    }
    if (lane == 0) dcan_raw[(size_t)f * cw * ch + (size_t)vc * cw + uc] = (int16_t)result;
}

// Row 0 / column 0 of D_can keep the calloc zero (elas.cpp:387): a valid disparity-0 neighbour for the filters.
__global__ void k_dcan_border(int16_t *__restrict__ dcan_raw, int cw, int ch) {
    int16_t *d = dcan_raw + (size_t)blockIdx.y * cw * ch;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cw) d[i] = 0;
    if (i < ch) d[(size_t)i * cw] = 0;
}

// ------------------------------------------------------------------------------------------------
// One CTA per frame runs the three in-place lattice filters with the reference's sequential
// (column-major) semantics, then compacts.
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

__device__ __forceinline__ bool precedes_colmajor(int u2, int v2, int u, int v) { return u2 < u || (u2 == u && v2 < v); }

__global__ void __launch_bounds__(SF_THREADS) k_support_filter(const int16_t *__restrict__ dcan_raw_all, int16_t *__restrict__ dcan_all,
                                                               uint8_t *__restrict__ removed_all, int32_t *__restrict__ support_all,
                                                               int32_t *__restrict__ nsupport_all, int W, int H, int cw, int ch, int step,
                                                               int incon_window, int incon_threshold, int incon_min_support, int add_corners,
                                                               int maxS) {
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    const int cells = cw * ch;
    const int16_t *raw = dcan_raw_all + (size_t)f * cells;
    int16_t *dcan = dcan_all + (size_t)f * cells;
    uint8_t *removed = removed_all + (size_t)f * cells;
    int32_t *support = support_all + (size_t)f * maxS * 3;

    __shared__ int s_changed;
    __shared__ int s_warp_tot[SF_THREADS / 32];
    __shared__ int s_base;
    __shared__ unsigned long long s_best[4];

    for (int i = tid; i < cells; i += SF_THREADS) removed[i] = 0;
    __syncthreads();

    // ---- inconsistent points: parallel sweeps to the fixpoint ----
    while (true) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < cells; i += SF_THREADS) {
            const int v = i / cw, u = i - v * cw;
            const int dc = raw[i];
            if (dc < 0 || removed[i]) continue;
            int support_cnt = 0;
            const int u_lo = max(u - incon_window, 0), u_hi = min(u + incon_window, cw - 1);
            const int v_lo = max(v - incon_window, 0), v_hi = min(v + incon_window, ch - 1);
            for (int v2 = v_lo; v2 <= v_hi; v2++)
                for (int u2 = u_lo; u2 <= u_hi; u2++) {
                    const int j = v2 * cw + u2;
                    const int d2 = raw[j];
                    if (d2 < 0 || abs(dc - d2) > incon_threshold) continue;
                    if (removed[j] && precedes_colmajor(u2, v2, u, v)) continue;
                    support_cnt++;
                }
            if (support_cnt < incon_min_support) {
                removed[i] = 1;
                s_changed = 1;
            }
        }
        __syncthreads();
        const int again = s_changed;
        __syncthreads();
        if (!again) break;
    }
    for (int i = tid; i < cells; i += SF_THREADS) dcan[i] = removed[i] ? (int16_t)-1 : raw[i];
    __syncthreads();

    // ---- redundant points, vertical pass (redun_max_dist 5, redun_threshold 1; elas.cpp:419) ----
    for (int u = tid; u < cw; u += SF_THREADS) {
        for (int v = 0; v < ch; v++) {
            const int dc = dcan[v * cw + u];
            if (dc < 0) continue;
            bool up = false, down = false;
            for (int j = 1; j <= 5 && v - j >= 0; j++) {
                const int d2 = dcan[(v - j) * cw + u];
                if (d2 >= 0 && abs(dc - d2) <= 1) {
                    up = true;
                    break;
                }
            }
            if (up)
                for (int j = 1; j <= 5 && v + j < ch; j++) {
                    const int d2 = dcan[(v + j) * cw + u];
                    if (d2 >= 0 && abs(dc - d2) <= 1) {
                        down = true;
                        break;
                    }
                }
            if (up && down) dcan[v * cw + u] = -1;
        }
    }
    __syncthreads();
    // ---- redundant points, horizontal pass (elas.cpp:420) ----
    for (int v = tid; v < ch; v += SF_THREADS) {
        for (int u = 0; u < cw; u++) {
            const int dc = dcan[v * cw + u];
            if (dc < 0) continue;
            bool left = false, right = false;
            for (int j = 1; j <= 5 && u - j >= 0; j++) {
                const int d2 = dcan[v * cw + u - j];
                if (d2 >= 0 && abs(dc - d2) <= 1) {
                    left = true;
                    break;
                }
            }
            if (left)
                for (int j = 1; j <= 5 && u + j < cw; j++) {
                    const int d2 = dcan[v * cw + u + j];
                    if (d2 >= 0 && abs(dc - d2) <= 1) {
                        right = true;
                        break;
                    }
                }
            if (left && right) dcan[v * cw + u] = -1;
        }
    }
    __syncthreads();

    // ---- ordered compaction: u_can outer, v_can inner, both from 1 (elas.cpp:424-428) ----
    const int inner = ch - 1;
    const int total = (cw - 1) * inner;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5;
    for (int start = 0; start < total; start += SF_THREADS) {
        const int i = start + tid;
        int uc = 0, vc = 0, dd = -1;
        if (i < total) {
            uc = 1 + i / inner;
            vc = 1 + i - (uc - 1) * inner;
            dd = dcan[vc * cw + uc];
        }
        const bool keep = dd >= 0;
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warp_tot[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; w++) off += s_warp_tot[w];
        if (keep) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            support[3 * pos + 0] = uc * step;
            support[3 * pos + 1] = vc * step;
            support[3 * pos + 2] = dd;
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
}

}  // namespace

int launch_support_match(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int nf,
                         cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    {
        int m = d.cw > d.ch ? d.cw : d.ch;
        dim3 grid((m + 127) / 128, nf);
        k_dcan_border<<<grid, 128, 0, s>>>(dcan_raw, d.cw, d.ch);
        SVB_LAUNCH_CHECK();
    }
    const int cells = (d.cw - 1) * (d.ch - 1);
    if (cells <= 0) return SVB_OK;
    dim3 grid((cells + SM_WARPS - 1) / SM_WARPS, nf);
    k_support_match<<<grid, SM_WARPS * 32, 0, s>>>(desc1, desc2, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                   p.support_texture, p.support_threshold, p.lr_threshold);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_support_filter(const Dims &d, const svb_params &p, const int16_t *dcan_raw, int16_t *dcan, uint8_t *scratch, int32_t *support,
                          int32_t *nsupport, int nf, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    k_support_filter<<<nf, SF_THREADS, 0, s>>>(dcan_raw, dcan, scratch, support, nsupport, d.W, d.H, d.cw, d.ch, d.step, p.incon_window_size,
                                               p.incon_threshold, p.incon_min_support, p.add_corners, d.maxS);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
