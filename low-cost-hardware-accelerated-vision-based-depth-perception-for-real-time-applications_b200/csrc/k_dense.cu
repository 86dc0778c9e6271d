// Dense matching: per pixel MAP disparity over (grid candidates outside the plane band) then (the plane band
// with the Gaussian prior), integer SAD over 16-byte descriptors.
//
// Replaces Elas::findMatch + updatePosteriorMinimum (src/serial_includes/elas/elas.cpp:655-802) as driven by
// Elas::computeDisparity (elas.cpp:804-944); the triangle -> pixel assignment comes from the owner map
// (k_prior.cu).  Evaluation order and the strict "<" (first evaluated wins ties) are the reference's.
#include <stdlib.h>

#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ uint32_t sad16(const uint4 &a, const uint4 &b) {
    return __vsadu4(a.x, b.x) + __vsadu4(a.y, b.y) + __vsadu4(a.z, b.z) + __vsadu4(a.w, b.w);
}

__device__ __forceinline__ int f2i_trunc_x86(float x) {
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(x);
}

struct DenseArgs {
    const uint8_t *desc[2];
    const int32_t *owner[2];
    const PlaneRec *rec[2];
    const uint32_t *grid[2];
    float *D[2];
    int W, H, maxT, gw, gh, gwords, grid_size, disp_max, match_texture, plane_radius;
    int row0;  // first map row of the launch (0, or the start of this device's band)
    int Dw, DN, shift;  // disparity map width / size and log2 of the pixel step (1 with subsampling: map pixel (x,y) = image (2x,2y))
    unsigned grid_magic;  // ceil(2^32 / grid_size)
    int P[8];
    const uint8_t *desc_lo[2], *desc_hi[2];  // the descriptor arenas including their guard bands (range asserts of the guard build)
    unsigned bias;  // Dims::cost_bias: keeps SAD + P >= 0 in the unsigned key
    unsigned long long *evals;  // COUNT variant only: number of evaluated hypotheses (elas.cpp:759-793)
    int owner_gen;              // generation of the owner map's entries (owner_untag); 0 = plain indices
};

__device__ __forceinline__ unsigned sad16_acc(const uint4 &a, const uint4 &b, unsigned acc) {
    // one dependent chain of VABSDIFF4.U8.ACC seeded with `acc`
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.x), "r"(b.x));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.y), "r"(b.y));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.z), "r"(b.z));
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc) : "r"(a.w), "r"(b.w));
    return acc;
}

__device__ __forceinline__ unsigned imad_fma_pipe(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// 16 bytes of shared memory at a 32-bit shared-window address
__device__ __forceinline__ uint4 lds128(unsigned addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

// base + idx as one IMAD.WIDE (FMA pipe) instead of a 4-instruction 64-bit LEA sequence on the ALU pipe
// guard build: a descriptor load must stay inside the arena (frames + zero-filled guard bands)
#define SVB_GUARD_DESC(ptr, side) \
    SVB_GUARD_ASSERT((const uint8_t *)(ptr) >= a.desc_lo[side] && (const uint8_t *)(ptr) + 16 <= a.desc_hi[side])

__device__ __forceinline__ const uint4 *desc_at(const uint4 *base, int idx) {
    unsigned long long r;
    asm("mad.wide.s32 %0, %1, 16, %2;" : "=l"(r) : "r"(idx), "l"(base));
    return reinterpret_cast<const uint4 *>(r);
}

// mask of the bits lo..hi (inclusive, any ints) that fall into a 32-bit word covering values base..base+31
__device__ __forceinline__ uint32_t range_mask(int lo, int hi, int base) {
    lo = max(lo - base, 0);
    hi = min(hi - base, 31);
    if (lo > hi) return 0u;
    return (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
}

// (i) of dense_body: the warp-uniform walk over the union of the lanes' grid candidates; returns the packed minimum.
// CLIP = false: no lane of the warp has an admissible interval [dlo, dhi] that cuts into [0, disp_max]; a lane's candidates of word w
// are then its cell's bits minus the band [dmin, dmax], which is a window of at most 15 bits starting in word dmin >> 5
// (5 instructions per word instead of the two general range masks).
template <int SIDE, bool CLIP, bool COUNT>
__device__ __forceinline__ unsigned grid_phase(const DenseArgs &a, const uint32_t *cell, const uint4 &m4_first, bool active, int dmin, int dmax,
                                               int dlo, int dhi, const uint4 &c, const uint4 *po, unsigned &n_hyp) {
    unsigned key = 0xFFFFFFFFu;
    const unsigned width = dmax >= dmin ? (2u << (dmax - dmin)) - 1u : 0u;
    const unsigned band_lo = width << (dmin & 31), band_hi = __funnelshift_l(width, 0u, dmin & 31);
    const int wband = dmin >> 5;  // dmin >= 0 wherever the band is not empty
    // gwords is a multiple of 4 (make_dims): a cell is read as 16-byte vectors, and a group of four words (128
    // disparities) without any candidate in the whole warp is skipped with one vote
    for (int w4 = 0; w4 < a.gwords; w4 += 4) {
        uint4 m4 = w4 == 0 ? m4_first : __ldg(reinterpret_cast<const uint4 *>(cell + w4));
        if (!active) m4 = make_uint4(0, 0, 0, 0);
        if (!__any_sync(0xFFFFFFFFu, (m4.x | m4.y | m4.z | m4.w) != 0u)) continue;
        const int rel = wband - w4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int w = w4 + j;
            uint32_t mine = j == 0 ? m4.x : j == 1 ? m4.y : j == 2 ? m4.z : m4.w;
            if (CLIP) {
                if (mine) mine &= ~range_mask(dmin, dmax, w << 5) & range_mask(dlo, dhi, w << 5);
            } else {
                mine &= ~((rel == j ? band_lo : 0u) | (rel == j - 1 ? band_hi : 0u));
            }
            if (COUNT) n_hyp += __popc(mine);
            uint32_t uni = __reduce_or_sync(0xFFFFFFFFu, mine);
            while (uni) {
                // two candidates per trip: both loads are in flight before either SAD chain starts
                const uint32_t bit0 = uni & (0u - uni);  // lowest candidate of the union
                uni ^= bit0;
                const uint32_t bit1 = uni & (0u - uni);  // next one (0 if there is none: it then re-evaluates bit0's d, masked out)
                uni ^= bit1;
                const int d0 = (w << 5) + (31 - __clz(bit0));
                const int d1 = bit1 ? (w << 5) + (31 - __clz(bit1)) : d0;
                SVB_GUARD_DESC(desc_at(po, SIDE ? d0 : -d0), SIDE ^ 1);
                SVB_GUARD_DESC(desc_at(po, SIDE ? d1 : -d1), SIDE ^ 1);
                const uint4 o0 = __ldg(desc_at(po, SIDE ? d0 : -d0));
                const uint4 o1 = __ldg(desc_at(po, SIDE ? d1 : -d1));
                const unsigned cand0 = (sad16_acc(c, o0, a.bias) << 13) + (unsigned)d0;
                const unsigned cand1 = (sad16_acc(c, o1, a.bias) << 13) + (unsigned)d1;
                key = min(key, (mine & bit0) ? cand0 : 0xFFFFFFFFu);
                key = min(key, (mine & bit1) ? cand1 : 0xFFFFFFFFu);
            }
        }
    }
    return key;
}

// grid: (ceil(W/128), H, nf*2); blockIdx.z = 2*frame + side.  One thread = one pixel, one warp = 32 consecutive
// pixels of a row.  The candidate loops are WARP-UNIFORM: the warp walks the union of its lanes' grid-cell bit
// masks (REDUX.OR) in ascending d, so in every step all lanes look at the same disparity -- their loads hit 32
// consecutive descriptors (one coalesced 512-byte access) and there is no loop divergence; a lane takes part in a
// step iff d is in ITS cell's list, outside ITS plane band and its warped column is inside the image.  The band
// (d_plane - r .. d_plane + r) is walked by the offset k, uniform as well, so the prior P[|k|] is a scalar.
//
// The running minimum is one packed key  (SAD + prior + 16) << 13 | phase << 12 | d :
//   smaller cost wins; on equal cost the grid phase (0) beats the band phase (1) and, inside a phase, the smaller d
//   wins -- exactly "first evaluated wins" of the reference's strict `<` over its evaluation order
//   (elas.cpp:757-794: grid candidates ascending, then the band ascending).  min_val starts at 10000 > any cost.
// RADIUS > 0: plane radius known at compile time (the band loop is fully unrolled and all of its loads are issued up
// front); RADIUS == 0: generic radius from the arguments.
template <int SIDE, int RADIUS, bool COUNT>
__device__ __forceinline__ void dense_body(const DenseArgs &a) {
    unsigned n_hyp = 0u;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;  // map pixel
    const int y = a.row0 + blockIdx.y;
    const int u = x << a.shift, v = y << a.shift;         // image pixel (elas.cpp:707-711: d_addr = (u/2, v/2) when subsampling)
    const int f = blockIdx.z >> 1;
    const int W = a.W, H = a.H;
    const bool in = x < a.Dw;
    // per-frame bases are warp-uniform 64-bit values; everything per pixel is a 32-bit offset from them (one IMAD.WIDE each)
    // (unsigned x unsigned -> 64 bit is one IMAD.WIDE.U32; signed operands would drag sign extensions along)
    const unsigned uf = (unsigned)f;
    const size_t fN = (size_t)uf * (unsigned)(W * H), fDN = (size_t)uf * (unsigned)a.DN;
    const int pix = y * a.Dw + (in ? x : 0);
    float *D = a.D[SIDE] + fDN;

    const int row = max(min(v, H - 3), 2);  // elas.cpp:718
    const int rowW = row * W;
    const uint4 *own = reinterpret_cast<const uint4 *>(a.desc[SIDE]) + fN;
    // descriptor of the other image at this pixel's own column; hypothesis d reads po[-d] (left) / po[+d] (right).
    // Rows are contiguous and 2 <= row <= H-3, so po[+-d] stays inside the frame's descriptor image for every
    // d <= disp_max even where the warped column leaves the row: loads never need a guard, only the result does.
    const uint4 *po = desc_at(reinterpret_cast<const uint4 *>(a.desc[SIDE ^ 1]) + fN, rowW + u);

    // Three independent loads are issued up front (owner index, own descriptor, first 128 candidate bits of the grid
    // cell), so that the prologue waits for ONE memory round trip before the plane record instead of four chained ones.
    const int uc = in ? min(u, W - 1) : W - 1;
    const int gx = a.grid_size == 1 ? uc : (int)__umulhi((unsigned)uc, a.grid_magic);  // u / grid_size by reciprocal (exact for u < 2^16)
    const int gy = a.grid_size == 1 ? v : (int)__umulhi((unsigned)v, a.grid_magic);    // u, v >= 0: equals the float floor (elas.cpp:744-745)
    const uint32_t *cell = a.grid[SIDE] + (size_t)uf * (unsigned)(a.gw * a.gh * a.gwords) + (unsigned)((gy * a.gw + gx) * a.gwords);
    const int o = in ? owner_untag(__ldg(a.owner[SIDE] + fDN + (unsigned)pix), a.owner_gen) : -1;
    SVB_GUARD_DESC(desc_at(own, rowW + uc), SIDE);
    SVB_GUARD_ASSERT(!in || (pix >= 0 && pix < a.DN));
    const uint4 c = __ldg(desc_at(own, rowW + uc));
    uint4 m4_first = __ldg(reinterpret_cast<const uint4 *>(cell));
    const uint4 k128 = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    // elas.cpp:714 (column range) and :731-736 (texture)
    const bool active = in && o >= 0 && u >= 2 && u < W - 2 && (int)sad16(c, k128) >= a.match_texture;
    int d_plane = 0, dmin = 1, dmax = 0;  // empty band for inactive lanes
    unsigned prior_on = 0u;
    if (active) {
        const PlaneRec pr = (a.rec[SIDE] + (size_t)uf * (unsigned)a.maxT)[(unsigned)o];
        // elas.cpp:739: (a*u + b*v) + c in f32, separate roundings, truncation like cvttss2si
        const float fp = __fadd_rn(__fadd_rn(__fmul_rn(pr.a, (float)u), __fmul_rn(pr.b, (float)v)), pr.c);
        d_plane = f2i_trunc_x86(fp);
        dmin = max((int)((unsigned)d_plane - (unsigned)a.plane_radius), 0);
        dmax = min((int)((unsigned)d_plane + (unsigned)a.plane_radius), a.disp_max);
        prior_on = pr.valid ? 0xFFFFFFFFu : 0u;
    }
    // warped column u -+ d must lie in [2, W-2): an interval of admissible d
    const int dlo = SIDE ? 2 - u : u - (W - 3);
    const int dhi = SIDE ? (W - 3) - u : u - 2;

    // (i) grid candidates outside the band, ascending (elas.cpp:759-767 / 778-786).  Only warps next to the image border (some
    // lane's admissible interval [dlo, dhi] cuts into [0, disp_max], the range of the grid's bits: k_grid_scatter) need the clipping
    // masks; everywhere else a lane's candidates are its cell's bits minus the band, and the band is one window of at most 15 bits.
    const bool clip = __any_sync(0xFFFFFFFFu, active && (dlo > 0 || dhi < a.disp_max));
    unsigned key = clip ? grid_phase<SIDE, true, COUNT>(a, cell, m4_first, active, dmin, dmax, dlo, dhi, c, po, n_hyp)
                        : grid_phase<SIDE, false, COUNT>(a, cell, m4_first, active, dmin, dmax, dlo, dhi, c, po, n_hyp);
    // (ii) the plane band with the prior (elas.cpp:768-774 / 787-793)
    const int lo2 = max(dmin, dlo), hi2 = min(dmax, dhi);
    const unsigned span = hi2 >= lo2 ? (unsigned)(hi2 - lo2) : 0u;
    const int lo3 = hi2 >= lo2 ? lo2 : 0x40000000;  // empty interval: nothing passes the unsigned range test
    if (RADIUS > 0) {
        // The band's descriptors sit at consecutive addresses around pb = po -+ d_plane: loads with immediate offsets, all issued up
        // front and unconditionally.  For the ADDRESS the plane disparity is clamped to [-8, disp_max + 8]: a band with any admissible
        // hypothesis has d_plane within the radius of [0, disp_max], where the clamp is the identity, and every other band is discarded
        // by okb anyway -- but its loads now stay within disp_max + 11 descriptors of po, i.e. inside the arena's guard bands
        // (Dims::desc_pad), so they need neither a predicate nor zeroed destination registers.
        const int d_addr = min(max(d_plane, -8), a.disp_max + 8);
        const uint4 *pb = desc_at(po, SIDE ? d_addr : -d_addr);
        uint4 ob[2 * RADIUS + 1];
        bool okb[2 * RADIUS + 1];
#pragma unroll
        for (int k = -RADIUS; k <= RADIUS; k++) {
            const int d = (int)((unsigned)d_plane + (unsigned)k);
            okb[k + RADIUS] = (unsigned)(d - lo3) <= span;
            SVB_GUARD_DESC(pb + (SIDE ? k : -k), SIDE ^ 1);
            ob[k + RADIUS] = __ldg(pb + (SIDE ? k : -k));
            if (COUNT) n_hyp += okb[k + RADIUS] ? 1u : 0u;
        }
        unsigned seed[RADIUS + 1];  // bias + prior of |k| (the prior only where both planes are valid, elas.cpp:910)
#pragma unroll
        for (int k = 0; k <= RADIUS; k++) seed[k] = a.bias + ((unsigned)a.P[k] & prior_on);
#pragma unroll
        for (int k = -RADIUS; k <= RADIUS; k++) {
            const unsigned dk = (unsigned)d_plane + (unsigned)(0x1000 + k);  // phase bit + d
            const unsigned cand = (sad16_acc(c, ob[k + RADIUS], seed[k < 0 ? -k : k]) << 13) + dk;
            key = min(key, okb[k + RADIUS] ? cand : 0xFFFFFFFFu);
        }
    } else {
        const int r = a.plane_radius;
        for (int k = -r; k <= r; k++) {
            const int d = (int)((unsigned)d_plane + (unsigned)k);
            const bool ok = (unsigned)(d - lo3) <= span;
            if (COUNT) n_hyp += ok ? 1u : 0u;
            const int ds = ok ? d : 0;
            const unsigned seed = a.bias + ((unsigned)a.P[k < 0 ? -k : k] & prior_on);
            SVB_GUARD_DESC(desc_at(po, SIDE ? ds : -ds), SIDE ^ 1);
            const unsigned cost = sad16_acc(c, __ldg(desc_at(po, SIDE ? ds : -ds)), seed);
            const unsigned cand = (cost << 13) + (0x1000u + (unsigned)ds);
            key = min(key, ok ? cand : 0xFFFFFFFFu);
        }
    }
    if (in) {
        float out = -10.f;                                                    // elas.cpp:820-826: pixels nobody writes keep -10
        if (active) out = key != 0xFFFFFFFFu ? (float)(key & 0xFFFu) : -1.f;  // elas.cpp:797-800
        D[(unsigned)pix] = out;
    }
    if (COUNT) {
        const unsigned tot = __reduce_add_sync(0xFFFFFFFFu, n_hyp);
        if ((threadIdx.x & 31) == 0 && tot) atomicAdd(a.evals + 1, (unsigned long long)tot);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Row form (full-resolution maps, plane radius 2 or 3, disp_max <= 1023): one CTA owns one image row of one side.  The row of the
// other image's descriptors every hypothesis of the row reads, and the bit masks of the one grid row its pixels fall into, are staged
// in shared memory once; a thread then walks pixels x, x + 256, ... of the row.  Against one pixel per thread this
//  * replaces the per-pixel 64-bit address arithmetic over six arrays by row bases computed once per CTA,
//  * lets every lane walk ITS OWN candidates (shared memory does not need the warp to agree on the disparity): a trip of the candidate
//    loop serves the lane's next two bits, so the trip count is the largest candidate count of a lane, not the size of the union,
//  * visits only the mask words that have a bit in some lane's cell (a per-cell summary of non-zero words, one REDUX per pixel).
// Arithmetic, evaluation-order semantics and the packed key are those of dense_body.
#ifndef SVB_DR_THREADS
#define SVB_DR_THREADS 256
#endif
#ifndef SVB_DR_LEAN
#define SVB_DR_LEAN 0  // candidate walk on bit positions inside a word (see dense_row_body)
#endif
constexpr int DR_THREADS = SVB_DR_THREADS;

template <int SIDE, int RADIUS, bool COUNT>
__device__ __forceinline__ void dense_row_body(const DenseArgs &a, uint4 *s_oth, uint32_t *s_cell, uint32_t *s_nz) {
    const int W = a.W, H = a.H;
    const int v = a.row0 + blockIdx.x;
    const unsigned uf = blockIdx.y >> 1;
    const int row = max(min(v, H - 3), 2);  // elas.cpp:718
    const size_t row_desc = (size_t)uf * (unsigned)(W * H) + (size_t)(row * W);
    const uint4 *own = reinterpret_cast<const uint4 *>(a.desc[SIDE]) + row_desc;
    const uint4 *oth = reinterpret_cast<const uint4 *>(a.desc[SIDE ^ 1]) + row_desc;
    const int pad = a.disp_max + 16;  // columns either side of the row: a hypothesis (or a clamped band address) may leave the row, its result is discarded
    for (int x = threadIdx.x; x < W; x += DR_THREADS) {
        SVB_GUARD_DESC(oth + x, SIDE ^ 1);
        s_oth[pad + x] = __ldg(oth + x);
    }
    const int gy = a.grid_size == 1 ? v : (int)__umulhi((unsigned)v, a.grid_magic);  // v >= 0: equals the float floor (elas.cpp:744-745)
    const int gwords = a.gwords;
    const uint32_t *cells = a.grid[SIDE] + (size_t)uf * (unsigned)(a.gw * a.gh * gwords) + (size_t)(gy * a.gw * gwords);
    for (int i = threadIdx.x; i < a.gw * gwords; i += DR_THREADS) s_cell[i] = __ldg(cells + i);
    __syncthreads();
    for (int g = threadIdx.x; g < a.gw; g += DR_THREADS) {
        unsigned nz = 0u;
        for (int w = 0; w < gwords; w++) nz |= s_cell[g * gwords + w] ? (1u << w) : 0u;
        s_nz[g] = nz;
    }
    __syncthreads();
    const size_t row_map = (size_t)uf * (unsigned)a.DN + (size_t)(v * W);
    const int32_t *owner = a.owner[SIDE] + row_map;
    float *D = a.D[SIDE] + row_map;
    const PlaneRec *rec = a.rec[SIDE] + (size_t)uf * (unsigned)a.maxT;
    const uint4 k128 = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
    const float fv = (float)v;
    unsigned n_hyp = 0u;
    const int lane = threadIdx.x & 31;
#if SVB_DR_LEAN
    const unsigned s_oth32 = (unsigned)__cvta_generic_to_shared(s_oth);
    const unsigned bias = a.bias * blockDim.z;  // x 1: a register instead of a constant-bank load in front of every trip of the candidate loop
#endif
    for (int u = threadIdx.x; u - lane < W; u += DR_THREADS) {  // whole warps: the loops below vote
        const bool in = u < W;
        const int uc = in ? u : W - 1;
        const int gx = a.grid_size == 1 ? uc : (int)__umulhi((unsigned)uc, a.grid_magic);
        const int o = in ? owner_untag(__ldg(owner + u), a.owner_gen) : -1;
        const uint4 c = __ldg(own + uc);
        // elas.cpp:714 (column range) and :731-736 (texture)
        const bool active = in && o >= 0 && u >= 2 && u < W - 2 && (int)sad16(c, k128) >= a.match_texture;
        int d_plane = 0, dmin = 1, dmax = 0;  // empty band for inactive lanes
        unsigned prior_on = 0u;
        if (active) {
            const PlaneRec pr = rec[(unsigned)o];
            // elas.cpp:739: (a*u + b*v) + c in f32, separate roundings, truncation like cvttss2si
            const float fp = __fadd_rn(__fadd_rn(__fmul_rn(pr.a, (float)u), __fmul_rn(pr.b, fv)), pr.c);
            d_plane = f2i_trunc_x86(fp);
            dmin = max((int)((unsigned)d_plane - (unsigned)a.plane_radius), 0);
            dmax = min((int)((unsigned)d_plane + (unsigned)a.plane_radius), a.disp_max);
            prior_on = pr.valid ? 0xFFFFFFFFu : 0u;
        }
        // warped column u -+ d must lie in [2, W-2): an interval of admissible d
        const int dlo = SIDE ? 2 - u : u - (W - 3);
        const int dhi = SIDE ? (W - 3) - u : u - 2;
        const uint4 *po = s_oth + pad + uc;  // hypothesis d reads po[-d] (left) / po[+d] (right)
#if SVB_DR_LEAN
        const unsigned po32 = s_oth32 + (unsigned)((pad + uc) << 4);
#endif

        // (i) grid candidates outside the band (elas.cpp:759-767 / 778-786); see grid_phase for the two mask forms
        unsigned key = 0xFFFFFFFFu;
        const bool clip = __any_sync(0xFFFFFFFFu, active && (dlo > 0 || dhi < a.disp_max));
        const unsigned width = dmax >= dmin ? (2u << (dmax - dmin)) - 1u : 0u;
        const unsigned band_lo = width << (dmin & 31), band_hi = __funnelshift_l(width, 0u, dmin & 31);
        const int wband = dmin >> 5;
        const uint32_t *cw = s_cell + gx * gwords;
        unsigned words = __reduce_or_sync(0xFFFFFFFFu, active ? s_nz[gx] : 0u);
        while (words) {
            const int w = __ffs(words) - 1;
            words &= words - 1u;
            uint32_t mine = active ? cw[w] : 0u;
            if (clip) {
                if (mine) mine &= ~range_mask(dmin, dmax, w << 5) & range_mask(dlo, dhi, w << 5);
            } else {
                mine &= ~((w == wband ? band_lo : 0u) | (w == wband + 1 ? band_hi : 0u));
            }
            if (COUNT) n_hyp += __popc(mine);
#if SVB_DR_LEAN
            // Inside a word only the bit position p = d - 32 w is carried: the address is the word's base column -+ p and the key
            // (cost << 13) + p, one IMAD each (FMA pipe); the word's minimum gets its 32 w once (adding a constant below bit 13 to all
            // keys of a word keeps their order, d < 8192).  A lane that has no candidate left gets p = -1 from bfind: the column next
            // to the word's first one, inside the staged row's margins; its result is discarded by the predicate.  27 instead of 37
            // instructions per trip.
            const unsigned pw = po32 + (unsigned)(SIDE ? (w << 9) : -(w << 9));  // byte address of column uc -+ 32 w in shared memory
            unsigned kw = 0xFFFFFFFFu;
            while (__any_sync(0xFFFFFFFFu, mine != 0u)) {
                const uint32_t bit0 = mine & (0u - mine);
                mine ^= bit0;
                const uint32_t bit1 = mine & (0u - mine);
                mine ^= bit1;
                unsigned p0, p1;
                asm("bfind.u32 %0, %1;" : "=r"(p0) : "r"(bit0));
                asm("bfind.u32 %0, %1;" : "=r"(p1) : "r"(bit1));
                SVB_GUARD_ASSERT((w << 5) + (int)p0 >= -1 && (w << 5) + (int)p0 <= pad && (w << 5) + (int)p1 >= -1 && (w << 5) + (int)p1 <= pad);
                const uint4 o0 = lds128(imad_fma_pipe(p0, SIDE ? 16u : 0xFFFFFFF0u, pw));
                const uint4 o1 = lds128(imad_fma_pipe(p1, SIDE ? 16u : 0xFFFFFFF0u, pw));
                const unsigned cand0 = imad_fma_pipe(sad16_acc(c, o0, bias), 0x2000u, p0);
                const unsigned cand1 = imad_fma_pipe(sad16_acc(c, o1, bias), 0x2000u, p1);
                if (bit0) kw = min(kw, cand0);
                if (bit1) kw = min(kw, cand1);
            }
            if (kw != 0xFFFFFFFFu) key = min(key, kw + (unsigned)(w << 5));
#else
            while (__any_sync(0xFFFFFFFFu, mine != 0u)) {
                // the lane's next two candidates: both loads are in flight before either SAD chain starts (a lane that has none left
                // re-reads the word's first column and discards the result)
                const uint32_t bit0 = mine & (0u - mine);
                mine ^= bit0;
                const uint32_t bit1 = mine & (0u - mine);
                mine ^= bit1;
                const int d0 = (w << 5) + (bit0 ? 31 - __clz(bit0) : 0);
                const int d1 = (w << 5) + (bit1 ? 31 - __clz(bit1) : 0);
                SVB_GUARD_ASSERT(d0 >= 0 && d0 <= pad && d1 >= 0 && d1 <= pad);
                const uint4 o0 = po[SIDE ? d0 : -d0];
                const uint4 o1 = po[SIDE ? d1 : -d1];
                const unsigned cand0 = (sad16_acc(c, o0, a.bias) << 13) + (unsigned)d0;
                const unsigned cand1 = (sad16_acc(c, o1, a.bias) << 13) + (unsigned)d1;
                key = min(key, bit0 ? cand0 : 0xFFFFFFFFu);
                key = min(key, bit1 ? cand1 : 0xFFFFFFFFu);
            }
#endif
        }
        // (ii) the plane band with the prior (elas.cpp:768-774 / 787-793)
        const int lo2 = max(dmin, dlo), hi2 = min(dmax, dhi);
        const unsigned span = hi2 >= lo2 ? (unsigned)(hi2 - lo2) : 0u;
        const int lo3 = hi2 >= lo2 ? lo2 : 0x40000000;  // empty interval: nothing passes the unsigned range test
        {
            // for the ADDRESS the plane disparity is clamped to [-8, disp_max + 8] (see dense_body): inside the staged row's margins
            const int d_addr = min(max(d_plane, -8), a.disp_max + 8);
            const uint4 *pb = po + (SIDE ? d_addr : -d_addr);
            uint4 ob[2 * RADIUS + 1];
            bool okb[2 * RADIUS + 1];
#pragma unroll
            for (int k = -RADIUS; k <= RADIUS; k++) {
                const int d = (int)((unsigned)d_plane + (unsigned)k);
                okb[k + RADIUS] = (unsigned)(d - lo3) <= span;
                ob[k + RADIUS] = pb[SIDE ? k : -k];
                if (COUNT) n_hyp += okb[k + RADIUS] ? 1u : 0u;
            }
            unsigned seed[RADIUS + 1];  // bias + prior of |k| (the prior only where both planes are valid, elas.cpp:910)
#pragma unroll
            for (int k = 0; k <= RADIUS; k++) seed[k] = a.bias + ((unsigned)a.P[k] & prior_on);
#pragma unroll
            for (int k = -RADIUS; k <= RADIUS; k++) {
                const unsigned dk = (unsigned)d_plane + (unsigned)(0x1000 + k);  // phase bit + d
                const unsigned cand = (sad16_acc(c, ob[k + RADIUS], seed[k < 0 ? -k : k]) << 13) + dk;
                key = min(key, okb[k + RADIUS] ? cand : 0xFFFFFFFFu);
            }
        }
        if (in) {
            float out = -10.f;                                                    // elas.cpp:820-826: pixels nobody writes keep -10
            if (active) out = key != 0xFFFFFFFFu ? (float)(key & 0xFFFu) : -1.f;  // elas.cpp:797-800
            D[u] = out;
        }
    }
    if (COUNT) {
        const unsigned tot = __reduce_add_sync(0xFFFFFFFFu, n_hyp);
        if (lane == 0 && tot) atomicAdd(a.evals + 1, (unsigned long long)tot);
    }
}

// grid: (rows, nf * 2); dynamic smem: (W + 2 (disp_max + 16)) uint4, then gw * gwords + gw words
template <int RADIUS, bool COUNT>
__global__ void __launch_bounds__(DR_THREADS) k_dense_row(const DenseArgs a) {
    extern __shared__ __align__(16) uint4 s_dense_row[];
    uint4 *s_oth = s_dense_row;
    uint32_t *s_cell = reinterpret_cast<uint32_t *>(s_dense_row + a.W + 2 * (a.disp_max + 16));
    uint32_t *s_nz = s_cell + a.gw * a.gwords;
    if (blockIdx.y & 1)
        dense_row_body<1, RADIUS, COUNT>(a, s_oth, s_cell, s_nz);
    else
        dense_row_body<0, RADIUS, COUNT>(a, s_oth, s_cell, s_nz);
}

template <int RADIUS, bool COUNT>
__global__ void __launch_bounds__(128, 12) k_dense(const DenseArgs a) {
    if (blockIdx.z & 1)
        dense_body<1, RADIUS, COUNT>(a);
    else
        dense_body<0, RADIUS, COUNT>(a);
}

}  // namespace

int launch_dense(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, const int32_t *owner1, const int32_t *owner2,
                 const PlaneRec *rec1, const PlaneRec *rec2, const uint32_t *grid1, const uint32_t *grid2, float *D1, float *D2, int nf,
                 cudaStream_t s, int owner_gen) {
    return launch_dense_rows(d, p, desc1, desc2, owner1, owner2, rec1, rec2, grid1, grid2, D1, D2, nf, 0, d.H, s, owner_gen);
}

int launch_dense_rows(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, const int32_t *owner1, const int32_t *owner2,
                      const PlaneRec *rec1, const PlaneRec *rec2, const uint32_t *grid1, const uint32_t *grid2, float *D1, float *D2, int nf,
                      int row0, int row1, cudaStream_t s, int owner_gen) {
    if (nf <= 0 || row1 <= row0) return SVB_OK;
    DenseArgs a;
    a.owner_gen = owner_gen;
    a.desc[0] = desc1;
    a.desc[1] = desc2;
    a.owner[0] = owner1;
    a.owner[1] = owner2;
    a.rec[0] = rec1;
    a.rec[1] = rec2;
    a.grid[0] = grid1;
    a.grid[1] = grid2;
    a.D[0] = D1;
    a.D[1] = D2;
    a.W = d.W;
    a.H = d.H;
    a.maxT = d.maxT;
    a.gw = d.gw;
    a.gh = d.gh;
    a.gwords = d.gwords;
    a.grid_size = p.grid_size;
    a.grid_magic = p.grid_size > 1 ? (unsigned)((0x100000000ull + p.grid_size - 1) / p.grid_size) : 0u;
    a.disp_max = p.disp_max;
    a.match_texture = p.match_texture;
    a.plane_radius = d.plane_radius;
    for (int i = 0; i < 8; i++) a.P[i] = d.P[i];
    a.bias = (unsigned)d.cost_bias;
    for (int sd = 0; sd < 2; sd++) {
        a.desc_lo[sd] = a.desc[sd] - d.desc_pad;
        a.desc_hi[sd] = a.desc[sd] + (size_t)nf * d.N * 16 + d.desc_pad;
    }
    a.row0 = row0;
    a.Dw = d.Dw;
    a.DN = d.DN;
    a.shift = d.sub ? 1 : 0;
    if (d.sub && (row0 != 0 || row1 != d.H)) {
        set_error("row-band dense matching with subsampling is not supported");
        return SVB_ERR_UNSUPPORTED;
    }
    const int rows = d.sub ? d.Dh : row1 - row0;
    a.evals = d.evals;
    // the row form wherever it applies (SVB_DENSE_ROWS=0: one pixel per thread everywhere)
    const char *rows_env = getenv("SVB_DENSE_ROWS");  // read per launch: the determinism stress test switches it between contexts
    const bool rows_off = rows_env && atoi(rows_env) == 0;
    const size_t row_smem = ((size_t)d.W + 2 * (p.disp_max + 16)) * sizeof(uint4) + ((size_t)d.gw * d.gwords + d.gw) * sizeof(uint32_t);
    if (!rows_off && !d.sub && (d.plane_radius == 2 || d.plane_radius == 3) && d.gwords <= 32 && row_smem <= 55 * 1024) {  // at least four CTAs per SM: wider rows (4K) are faster one pixel per thread
        const int which = (d.plane_radius == 3 ? 2 : 0) + (d.evals ? 1 : 0);
        const void *fn = which == 0   ? (const void *)k_dense_row<2, false>
                         : which == 1 ? (const void *)k_dense_row<2, true>
                         : which == 2 ? (const void *)k_dense_row<3, false>
                                      : (const void *)k_dense_row<3, true>;
        static size_t configured[64][4] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (row_smem > 48 * 1024 && dev >= 0 && dev < 64 && configured[dev][which] < row_smem) {
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem);
            if (e != cudaSuccess) {
                set_error("cudaFuncSetAttribute(k_dense_row, %zu): %s", row_smem, cudaGetErrorString(e));
                return SVB_ERR_CUDA;
            }
            configured[dev][which] = row_smem;
        }
        const dim3 rgrid(rows, nf * 2);
        if (which == 0)
            k_dense_row<2, false><<<rgrid, DR_THREADS, row_smem, s>>>(a);
        else if (which == 1)
            k_dense_row<2, true><<<rgrid, DR_THREADS, row_smem, s>>>(a);
        else if (which == 2)
            k_dense_row<3, false><<<rgrid, DR_THREADS, row_smem, s>>>(a);
        else
            k_dense_row<3, true><<<rgrid, DR_THREADS, row_smem, s>>>(a);
        SVB_LAUNCH_CHECK();
        return SVB_OK;
    }
    dim3 grid((d.Dw + 127) / 128, rows, nf * 2);
    if (d.evals) {
        if (d.plane_radius == 2)
            k_dense<2, true><<<grid, 128, 0, s>>>(a);
        else if (d.plane_radius == 3)
            k_dense<3, true><<<grid, 128, 0, s>>>(a);
        else
            k_dense<0, true><<<grid, 128, 0, s>>>(a);
    } else if (d.plane_radius == 2)
        k_dense<2, false><<<grid, 128, 0, s>>>(a);
    else if (d.plane_radius == 3)
        k_dense<3, false><<<grid, 128, 0, s>>>(a);
    else
        k_dense<0, false><<<grid, 128, 0, s>>>(a);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
