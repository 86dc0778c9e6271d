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

// ------------------------------------------------------------------------------------------------
// Matching (elas.cpp:266-371 for every lattice candidate, forward then backward, elas.cpp:394-411).
//
// energy(u, d) = sum over the 4 anchors (u-+2, v-+2) of SAD16(own anchor, other image at the anchor shifted by d)
// depends on the candidate column u and on the MATCHED column x = u -+ d.  The kernel is organised around x:
// a warp owns 32 consecutive candidates of one lattice row (lane l <-> candidate l: its anchors sit in shared
// memory, its running best / second best in the lane's registers) and walks over blocks of 32 consecutive matched
// columns.  For a block every lane loads the four "other image" descriptors of ITS column once (coalesced), then
// the warp loops over the candidates whose disparity range overlaps the block: the anchors come from shared
// memory as broadcasts (one wavefront per load instead of four), each lane evaluates one hypothesis, and two
// REDUX.MIN give the block's smallest and second smallest (energy, d) keys, which the owning lane merges.
// Per hypothesis this costs 16 VABSDIFF4 + 5 broadcast LDS.128 per 32 hypotheses, against 4 lane-distinct
// LDG.128 (16 L1 wavefronts per 32 hypotheses) of the candidate-major formulation.
//
// best = smallest energy, lowest d on ties; second = second smallest energy of the multiset (elas.cpp:352-360).
constexpr int SM_WARPS = 4;

struct __align__(16) CandMeta {
    int u, xlo, xhi, pad;
};

__device__ __forceinline__ int match_pass(const uint4 *__restrict__ own, const uint4 *__restrict__ oth, int u, bool has, int v, bool right_image,
                                          int W, int H, int disp_min, int disp_max, int support_texture, float support_threshold,
                                          uint4 (*s_anchor)[4], CandMeta *s_meta, int lane) {
    // candidate validity and disparity range (elas.cpp:279,296-300,318-327)
    bool ok = has && u >= 5 && u <= W - 6 && v >= 5 && v <= H - 6;
    if (ok) ok = (int)texture16(__ldg(own + (size_t)v * W + u)) >= support_texture;
    const int dmin = max(disp_min, 0);
    const int dmax = right_image ? min(disp_max, W - u - 5) : min(disp_max, u - 5);
    ok = ok && (dmax - dmin >= 10);
    // matched columns x = u - d (left candidate) or u + d (right candidate)
    const int xlo = ok ? (right_image ? u + dmin : u - dmax) : 0x7FFFFFFF;
    const int xhi = ok ? (right_image ? u + dmax : u - dmin) : (int)0x80000000;
    const unsigned okmask = __ballot_sync(0xFFFFFFFFu, ok);
    if (!okmask) return -1;
    const size_t rowt = (size_t)(v - 2) * W, rowb = (size_t)(v + 2) * W;  // v >= 5 whenever any lane is ok
    __syncwarp();
    if (ok) {
        s_anchor[lane][0] = __ldg(own + rowt + u - 2);
        s_anchor[lane][1] = __ldg(own + rowt + u + 2);
        s_anchor[lane][2] = __ldg(own + rowb + u - 2);
        s_anchor[lane][3] = __ldg(own + rowb + u + 2);
        CandMeta m;
        m.u = u;
        m.xlo = xlo;
        m.xhi = xhi;
        m.pad = 0;
        s_meta[lane] = m;
    }
    __syncwarp();
    const int wxlo = __reduce_min_sync(0xFFFFFFFFu, xlo);
    const int wxhi = __reduce_max_sync(0xFFFFFFFFu, xhi);
    unsigned best = 0xFFFFFFFFu, second = 0xFFFFFFFFu;  // (E << 16) | d
    for (int xb = wxlo; xb <= wxhi; xb += 32) {
        const int x = xb + lane;
        uint4 ot0 = make_uint4(0, 0, 0, 0), ot1 = ot0, ob0 = ot0, ob1 = ot0;
        if (x >= 2 && x + 2 < W) {
            ot0 = __ldg(oth + rowt + x - 2);
            ot1 = __ldg(oth + rowt + x + 2);
            ob0 = __ldg(oth + rowb + x - 2);
            ob1 = __ldg(oth + rowb + x + 2);
        }
        unsigned m = __ballot_sync(0xFFFFFFFFu, ok && xlo <= xb + 31 && xhi >= xb);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const CandMeta cm = s_meta[j];
            const uint4 a0 = s_anchor[j][0], a1 = s_anchor[j][1], a2 = s_anchor[j][2], a3 = s_anchor[j][3];
            const unsigned e = sad16(a0, ot0) + sad16(a1, ot1) + sad16(a2, ob0) + sad16(a3, ob1);
            const int d = right_image ? x - cm.u : cm.u - x;
            const unsigned key = (x >= cm.xlo && x <= cm.xhi) ? ((e << 16) | (unsigned)d) : 0xFFFFFFFFu;
            const unsigned k1 = __reduce_min_sync(0xFFFFFFFFu, key);
            const unsigned k2 = __reduce_min_sync(0xFFFFFFFFu, key == k1 ? 0xFFFFFFFFu : key);  // valid keys are distinct (distinct d)
            if (lane == j) {
                const unsigned hi = max(best, k1);
                best = min(best, k1);
                second = min(hi, min(second, k2));
            }
        }
    }
    if (!ok) return -1;
    // at least 11 hypotheses were evaluated, so both minima exist (min_1_d >= 0 && min_2_d >= 0)
    const int min1_e = (int)(best >> 16), min1_d = (int)(best & 0xFFFFu), min2_e = (int)(second >> 16);
    if ((float)min1_e < __fmul_rn(support_threshold, (float)min2_e)) return min1_d;  // elas.cpp:364
    return -1;
}

// grid: (ceil(rows * groups / SM_WARPS), nf); one warp = 32 consecutive candidates of one lattice row
__global__ void __launch_bounds__(SM_WARPS * 32) k_support_match(const uint8_t *__restrict__ desc1, const uint8_t *__restrict__ desc2,
                                                                 int16_t *__restrict__ dcan_raw, int W, int H, int cw, int ch, int step,
                                                                 int disp_min, int disp_max, int support_texture, float support_threshold,
                                                                 int lr_threshold) {
    __shared__ uint4 s_anchor[SM_WARPS][32][4];
    __shared__ CandMeta s_meta[SM_WARPS][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int groups = (cw - 1 + 31) / 32;
    const int wg = blockIdx.x * SM_WARPS + wid;
    const int row = wg / groups;
    if (row >= ch - 1) return;  // warp-uniform
    const int g = wg - row * groups;
    const int f = blockIdx.y;
    const int vc = 1 + row, uc = 1 + g * 32 + lane;
    const bool has = uc < cw;
    const int u = uc * step, v = vc * step;
    const size_t fo = (size_t)f * W * H;
    const uint4 *d1 = reinterpret_cast<const uint4 *>(desc1) + fo;
    const uint4 *d2 = reinterpret_cast<const uint4 *>(desc2) + fo;

    // forward: candidate in the left image, search the right image (elas.cpp:403)
    const int d = match_pass(d1, d2, u, has, v, false, W, H, disp_min, disp_max, support_texture, support_threshold, s_anchor[wid], s_meta[wid],
                             lane);
    // backward: the match (u-d, v) as a candidate of the right image, search the left image (elas.cpp:406)
    const int dback = match_pass(d2, d1, u - d, has && d >= 0, v, true, W, H, disp_min, disp_max, support_texture, support_threshold,
                                 s_anchor[wid], s_meta[wid], lane);
    int result = -1;
    if (d >= 0 && dback >= 0 && abs(d - dback) <= lr_threshold) result = d;  // elas.cpp:404-409
    if (has) dcan_raw[(size_t)f * cw * ch + (size_t)vc * cw + uc] = (int16_t)result;
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
constexpr int SF_REMOVED = 0x4000;  // flag bit of a lattice cell (disparities are < 4096)

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

// SMEM = true: the lattice lives in shared memory (2 bytes per cell; every frame size up to about 1080p at step 5);
// SMEM = false: it is worked on in place in the global dcan array (4K frames).
template <bool SMEM>
__global__ void __launch_bounds__(SF_THREADS) k_support_filter(const int16_t *__restrict__ dcan_raw_all, int16_t *__restrict__ dcan_all,
                                                               int32_t *__restrict__ support_all, int32_t *__restrict__ nsupport_all, int W, int H,
                                                               int cw, int ch, int step, int incon_window, int incon_threshold,
                                                               int incon_min_support, int add_corners, int maxS) {
    extern __shared__ int16_t s_lattice[];
    const int f = blockIdx.x;
    const int tid = threadIdx.x;
    const int cells = cw * ch;
    const int16_t *raw = dcan_raw_all + (size_t)f * cells;
    int16_t *dcan = dcan_all + (size_t)f * cells;
    int16_t *work = SMEM ? s_lattice : dcan;
    int32_t *support = support_all + (size_t)f * maxS * 3;

    __shared__ int s_changed;
    __shared__ int s_warp_tot[SF_THREADS / 32];
    __shared__ int s_base;
    __shared__ unsigned long long s_best[4];

    for (int i = tid; i < cells; i += SF_THREADS) work[i] = raw[i];
    __syncthreads();

    // ---- inconsistent points: parallel sweeps to the fixpoint; a removed cell keeps its value and gets SF_REMOVED ----
    while (true) {
        if (tid == 0) s_changed = 0;
        __syncthreads();
        for (int i = tid; i < cells; i += SF_THREADS) {
            const int e = work[i];
            if (e < 0 || (e & SF_REMOVED)) continue;
            const int v = i / cw, u = i - v * cw;
            const int u_lo = max(u - incon_window, 0), u_hi = min(u + incon_window, cw - 1);
            const int v_lo = max(v - incon_window, 0), v_hi = min(v + incon_window, ch - 1);
            int support_cnt = 0;
            for (int v2 = v_lo; v2 <= v_hi && support_cnt < incon_min_support; v2++)
                for (int u2 = u_lo; u2 <= u_hi; u2++) {
                    const int e2 = work[v2 * cw + u2];
                    if (e2 < 0) continue;
                    if (abs(e - (e2 & (SF_REMOVED - 1))) > incon_threshold) continue;
                    if ((e2 & SF_REMOVED) && precedes_colmajor(u2, v2, u, v)) continue;
                    support_cnt++;
                }
            if (support_cnt < incon_min_support) {
                work[i] = (int16_t)(e | SF_REMOVED);
                s_changed = 1;
            }
        }
        __syncthreads();
        const int again = s_changed;
        __syncthreads();
        if (!again) break;
    }
    for (int i = tid; i < cells; i += SF_THREADS) {
        const int e = work[i];
        if (e >= 0 && (e & SF_REMOVED)) work[i] = (int16_t)-1;
    }
    __syncthreads();

    // ---- redundant points: vertical pass (columns independent), then horizontal pass (rows independent) ----
    for (int u = tid; u < cw; u += SF_THREADS) redundant_line(work, u, cw, ch);
    __syncthreads();
    for (int v = tid; v < ch; v += SF_THREADS) redundant_line(work, v * cw, 1, cw);
    __syncthreads();
    if (SMEM)
        for (int i = tid; i < cells; i += SF_THREADS) dcan[i] = work[i];

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
            dd = work[vc * cw + uc];
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
    if (d.cw < 2 || d.ch < 2) return SVB_OK;
    const int warps = (d.ch - 1) * ((d.cw - 1 + 31) / 32);
    dim3 grid((warps + SM_WARPS - 1) / SM_WARPS, nf);
    k_support_match<<<grid, SM_WARPS * 32, 0, s>>>(desc1, desc2, dcan_raw, d.W, d.H, d.cw, d.ch, d.step, p.disp_min, p.disp_max,
                                                   p.support_texture, p.support_threshold, p.lr_threshold);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_support_filter(const Dims &d, const svb_params &p, const int16_t *dcan_raw, int16_t *dcan, uint8_t *scratch, int32_t *support,
                          int32_t *nsupport, int nf, cudaStream_t s) {
    (void)scratch;
    if (nf <= 0) return SVB_OK;
    if (p.disp_max >= SF_REMOVED) {
        set_error("support filter: disp_max %d too large for the lattice cell encoding", p.disp_max);
        return SVB_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)d.cw * d.ch * sizeof(int16_t);
    if (smem <= 200 * 1024) {
        if (smem > 48 * 1024) {  // opt in to large dynamic shared memory (per device, so not cached in a static)
            cudaError_t e = cudaFuncSetAttribute(k_support_filter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(k_support_filter): %s", cudaGetErrorString(e));
                return SVB_ERR_CUDA;
            }
        }
        k_support_filter<true><<<nf, SF_THREADS, smem, s>>>(dcan_raw, dcan, support, nsupport, d.W, d.H, d.cw, d.ch, d.step, p.incon_window_size,
                                                            p.incon_threshold, p.incon_min_support, p.add_corners, d.maxS);
    } else {
        k_support_filter<false><<<nf, SF_THREADS, 0, s>>>(dcan_raw, dcan, support, nsupport, d.W, d.H, d.cw, d.ch, d.step, p.incon_window_size,
                                                          p.incon_threshold, p.incon_min_support, p.add_corners, d.maxS);
    }
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
