// Vertex order for the host Delaunay stage, computed on the device.
//
// The reference's triangulator (Triangle 1.6, src/common_includes/elas/triangle.cpp) starts with two data-parallel
// steps before its inherently sequential divide-and-conquer: the lexicographic (x, y) sort of the vertices
// (triangle.cpp:5882-5903) and the alternating-axis median partition (:5243-5325, :5904-5913).  For duplicate-free
// input their result -- the order in which the recursion meets the vertices -- is unique (host_delaunay.cpp explains
// why only the SETS on either side of each median matter), so it is produced here, one CTA per (frame, image side),
// right behind the support-list compaction, and the host stage starts directly with the recursion (about 15 % less
// host time per frame).  DUPLICATE coordinates (common in the right image of real frames, where x = u - d: two thirds of the
// kitti_mini right lists have some): the reference drops all but the FIRST vertex of every run of equal coordinates in the order its
// unstable randomised quicksort leaves them (triangle.cpp:5183-5229 with the LCG of :3833-3836, 5889-5903), so for such a list one
// thread replays exactly that quicksort on the packed keys -- sequential, a few hundred microseconds, hidden like the rest of this
// latency-bound stage -- and the kernel goes on with the survivors.  Anything else the kernel does not handle -- fewer than 3 or more
// than 4096 points, coordinates outside the packed key range -- is flagged, and the host runs its complete path for that list.
// The flag written for a usable list is the number of vertices that take part (= n without duplicates).
//
//   1. bitonic sort of key = (x + 8192) << 13 | y  -> x rank;   2. bitonic sort of (y << 14 | x + 8192) -> y rank
//   3. k-d split over the two rank-ordered lists: at every level each node of >= 4 vertices is cut at its median along
//      the level's axis; the list ordered along that axis is cut in place, the other one is split stably (one block-wide
//      prefix sum per level).  Nodes of <= 3 stay x-sorted.
#include "svb_internal.h"

namespace svb {

namespace {

constexpr int OR_THREADS = 1024;
constexpr int OR_MAXN = 4096;

struct OrderSmem {
    unsigned long long K[OR_MAXN];  // sort keys: key << 32 | payload
    uint32_t L[3][OR_MAXN];         // rank pairs (xrank << 16 | yrank): x-ordered list, y-ordered list, spare
    uint16_t node_start[OR_MAXN];   // per POSITION: the node (contiguous range) it belongs to at the current level
    uint16_t node_size[OR_MAXN];
    uint16_t orig[OR_MAXN];         // support index of the vertex with x rank i
    uint16_t prefix[OR_MAXN + 2];   // exclusive prefix sums of the "goes to the low side" flags
    int warp_total[OR_THREADS / 32];
    int bad;
};

// Thread t handles the indices i = t (mod 1024); for j < 32 both partners i and i ^ j belong to lanes of one warp, so
// such a step only needs a warp barrier -- unless the step before it exchanged data between warps.
__device__ __forceinline__ void bitonic_sort(unsigned long long *K, int np2) {
    bool prev_wide = true;  // the keys were written block-wide
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool wide = j >= 32;
            if (wide || prev_wide)
                __syncthreads();
            else
                __syncwarp();
            prev_wide = wide;
            for (int i = threadIdx.x; i < np2; i += OR_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = K[i], b = K[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) {
                        K[i] = b;
                        K[ixj] = a;
                    }
                }
            }
        }
    }
    __syncthreads();
}

// The reference's vertexsort (triangle.cpp:5183-5229) on keys a[i] >> 32, replayed exactly: random pivot from the LCG seeded with 1
// (randomnation, :3833-3836; triangleinit resets the seed for every triangulation), Hoare partition with strict comparisons on both
// sides, left subset first.  The recursion and every single partition are sequential by nature -- which element a scan stops at
// depends on all swaps before it -- but the SCANS need not visit elements one by one: a warp looks at 32 elements at a time, turns
// "would the left (right) scan stop here" into a ballot mask, and jumps from stopper to stopper.  The k-th stop of the left scan is
// always an element the partition has not touched yet, so masks taken before a swap stay valid for everything still ahead of the
// two cursors:
//   left scan : next index l in (l_prev, r_prev) with key >= pivot, else l = r_prev (that slot now holds a swapped-in key >= pivot)
//   right scan: next index r in [l, r_prev) from above with key <= pivot, else r = l - 1
//   l < r: swap, go on;  otherwise the partition ends with (left, right) = (l, r)                        (triangle.cpp:5207-5221)
// All lanes run the same control flow; lane 0 swaps.  `stack` holds (start, length) pairs; false if it overflows (host's turn then).
__device__ bool replay_vertexsort(unsigned long long *a, int n, uint32_t *stack, int stack_cap, int lane) {
    unsigned seed = 1u;
    int sp = 2;
    if (lane == 0) {
        stack[0] = 0u;
        stack[1] = (uint32_t)n;
    }
    __syncwarp();
    while (sp > 0) {
        const int len = (int)stack[sp - 1];
        const int lo = (int)stack[sp - 2];
        sp -= 2;
        unsigned long long *s = a + lo;
        if (len == 2) {
            if (lane == 0 && (unsigned)(s[0] >> 32) > (unsigned)(s[1] >> 32)) {
                const unsigned long long t = s[0];
                s[0] = s[1];
                s[1] = t;
            }
            __syncwarp();
            continue;
        }
        seed = (seed * 1366u + 150889u) % 714025u;
        const int pivot = (int)(seed / (714025u / (unsigned)len + 1u));
        const unsigned pv = (unsigned)(s[pivot] >> 32);
        int lc = 0, rc = (len - 1) & ~31;  // bases of the 32-element windows the two scans are in
        unsigned lmask = __ballot_sync(0xFFFFFFFFu, lc + lane < len && (unsigned)(s[lc + lane] >> 32) >= pv);
        unsigned rmask = __ballot_sync(0xFFFFFFFFu, rc + lane < len && (unsigned)(s[rc + lane] >> 32) <= pv);
        int r_prev = len, left, right;
        while (true) {
            while (lmask == 0u && lc + 32 < r_prev) {  // r_prev <= len
                lc += 32;
                lmask = __ballot_sync(0xFFFFFFFFu, lc + lane < len && (unsigned)(s[lc + lane] >> 32) >= pv);
            }
            int l = lmask ? lc + __ffs(lmask) - 1 : r_prev;
            if (l > r_prev) l = r_prev;
            while (rmask == 0u && rc > l) {
                rc -= 32;
                rmask = __ballot_sync(0xFFFFFFFFu, (unsigned)(s[rc + lane] >> 32) <= pv);
            }
            int r = rmask ? rc + 31 - __clz(rmask) : -1;
            if (r < l) r = l - 1;
            left = l;
            right = r;
            if (l >= r) break;
            if (lane == 0) {
                const unsigned long long t = s[l];
                s[l] = s[r];
                s[r] = t;
            }
            lmask &= lmask - 1u;         // l consumed
            rmask &= ~(1u << (r - rc));  // r consumed
            r_prev = r;
            __syncwarp();
        }
        __syncwarp();
        // the reference recurses into the left subset first: push the right one below it
        if (sp + 4 > stack_cap) return false;
        if (right < len - 2) {
            if (lane == 0) {
                stack[sp] = (uint32_t)(lo + right + 1);
                stack[sp + 1] = (uint32_t)(len - right - 1);
            }
            sp += 2;
        }
        if (left > 1) {
            if (lane == 0) {
                stack[sp] = (uint32_t)lo;
                stack[sp + 1] = (uint32_t)left;
            }
            sp += 2;
        }
        __syncwarp();
    }
    return true;
}

// Lists flagged by pass 0 of k_delaunay_order: the reference's vertexsort + duplicate removal (triangle.cpp:5889-5903: the first of every
// run of equal coordinates stays), replayed by a one-warp CTA.  grid: (nf, 2), 32 threads;
// dynamic smem: OR_MAXN keys (8 B) + RS_STACK stack words
constexpr int RS_STACK = 2048;
__global__ void __launch_bounds__(32) k_replay_vertexsort(const int32_t *__restrict__ support_all, const int32_t *__restrict__ nsupport_all,
                                                         int32_t *__restrict__ dup_count_all, unsigned long long *__restrict__ dup_keys_all, int maxS) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    unsigned long long *K = reinterpret_cast<unsigned long long *>(rs_smem);
    uint32_t *stack = reinterpret_cast<uint32_t *>(rs_smem + sizeof(unsigned long long) * OR_MAXN);
    __shared__ int s_m;
    const int f = blockIdx.x, side = blockIdx.y, lane = threadIdx.x;
    int32_t *dup_count = dup_count_all + 2 * f + side;
    if (*dup_count != -1) return;
    const int n = nsupport_all[f];
    const int32_t *support = support_all + (size_t)f * maxS * 3;
    // the input order: vertex i = support point i (elas.cpp:451-461)
    for (int i = lane; i < n; i += 32) {
        const int u = support[3 * i], v = support[3 * i + 1], d = support[3 * i + 2];
        const unsigned xb = (unsigned)((side ? u - d : u) + 8192), yb = (unsigned)v;
        K[i] = (unsigned long long)(xb << 13 | (yb & 8191u)) << 32 | (unsigned)i;
    }
    __syncwarp();
    const bool sorted = replay_vertexsort(K, n, stack, RS_STACK, lane);
    if (lane == 0) {
        int m = 0;
        if (sorted) {
            for (int j = 1; j < n; j++)
                if ((K[m] >> 32) != (K[j] >> 32)) K[++m] = K[j];
            m++;
        }
        s_m = m;
    }
    __syncwarp();
    const int m = s_m;
    unsigned long long *out = dup_keys_all + ((size_t)f * 2 + side) * OR_MAXN;
    for (int i = lane; i < m; i += 32) out[i] = K[i];
    if (lane == 0) *dup_count = m;  // 0: the replay gave up (stack), 1 or 2: fewer than 3 distinct vertices
}

// grid: (nf, 2); blockIdx.y = image side (0: (u, v), 1: (u - d, v), elas.cpp:451-461)
__global__ void __launch_bounds__(OR_THREADS) k_delaunay_order(const int32_t *__restrict__ support_all, const int32_t *__restrict__ nsupport_all,
                                                              int32_t *__restrict__ h_order_all, int32_t *__restrict__ h_ok_all,
                                                              int32_t *__restrict__ d_order_all, int32_t *__restrict__ d_ok_all, int maxS,
                                                              int pass, int32_t *__restrict__ dup_count_all,
                                                              unsigned long long *__restrict__ dup_keys_all) {
    // pass 0: every list; one with duplicate coordinates is only FLAGGED (dup_count = -1) and left.  k_replay_vertexsort then sorts and
    // de-duplicates the flagged lists the way the reference does (dup_count = number of survivors, their keys in dup_keys), and
    // pass 1 -- flagged lists only -- picks up from there.  The replay is one thread's work for up to a millisecond: it runs in a
    // one-warp CTA of its own instead of keeping this kernel's 1024 threads and 100 KB of shared memory resident.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OrderSmem &S = *reinterpret_cast<OrderSmem *>(smem_raw);
    const int f = blockIdx.x, side = blockIdx.y, tid = threadIdx.x;
    int n = nsupport_all[f];  // becomes the number of distinct vertices if the list holds duplicates
    int32_t *dup_count = dup_count_all ? dup_count_all + 2 * f + side : nullptr;
    if (pass == 1 && (!dup_count || *dup_count == 0)) return;  // nothing was flagged for this list
    int32_t *ok_out = h_ok_all + 2 * f + side;
    int32_t *ok_dev = d_ok_all ? d_ok_all + 2 * f + side : nullptr;  // device copies for the divide-and-conquer kernel (k_delaunay.cu)
    if (n < 3 || n > OR_MAXN || n > maxS) {  // uniform
        if (tid == 0) {
            *ok_out = 0;
            if (ok_dev) *ok_dev = 0;
            if (dup_count) *dup_count = 0;
        }
        return;
    }
    const int32_t *support = support_all + (size_t)f * maxS * 3;
    int32_t *order = h_order_all + ((size_t)f * 2 + side) * maxS;
    int np2 = 4;
    while (np2 < n) np2 <<= 1;
    if (tid == 0) S.bad = 0;
    __syncthreads();
    __shared__ int s_dup;
    if (pass == 1) {
        // the reference's own sort and duplicate removal have been replayed: distinct keys, in order, with their support indices
        n = *dup_count;
        if (n < 3) {  // fewer than 3 distinct vertices, or the replay gave up: the host's turn (it finds no triangle, or does it all)
            if (tid == 0) {
                *ok_out = 0;
                if (ok_dev) *ok_dev = 0;
            }
            return;
        }
        np2 = 4;
        while (np2 < n) np2 <<= 1;
        const unsigned long long *keys = dup_keys_all + ((size_t)f * 2 + side) * OR_MAXN;
        for (int i = tid; i < np2; i += OR_THREADS) S.K[i] = i < n ? keys[i] : ~0ull;
        __syncthreads();
    } else {
    // 1. x rank
    for (int i = tid; i < np2; i += OR_THREADS) {
        unsigned long long k = ~0ull;
        if (i < n) {
            const int u = support[3 * i], v = support[3 * i + 1], d = support[3 * i + 2];
            const unsigned xb = (unsigned)((side ? u - d : u) + 8192), yb = (unsigned)v;
            if ((xb >> 14) | (yb >> 13)) S.bad = 1;
            k = (unsigned long long)(xb << 13 | (yb & 8191u)) << 32 | (unsigned)i;
        }
        S.K[i] = k;
    }
    bitonic_sort(S.K, np2);
    if (tid == 0) s_dup = 0;
    __syncthreads();
    for (int i = tid; i < n; i += OR_THREADS)
        if (i > 0 && (S.K[i] >> 32) == (S.K[i - 1] >> 32)) s_dup = 1;  // duplicate coordinates: which one survives is the reference's sort's call
    __syncthreads();
    if (S.bad || (s_dup && !dup_count)) {
        if (tid == 0) {
            *ok_out = 0;
            if (ok_dev) *ok_dev = 0;
            if (dup_count) *dup_count = 0;
        }
        return;
    }
    if (tid == 0 && dup_count) *dup_count = s_dup ? -1 : 0;
    if (s_dup) {
        if (tid == 0) {  // not usable yet; pass 1 overwrites this
            *ok_out = 0;
            if (ok_dev) *ok_dev = 0;
        }
        return;
    }
    }  // pass 0
    // 2. y rank: sort (y, x) keys that carry the x rank
    for (int i = tid; i < np2; i += OR_THREADS) {
        unsigned long long k2 = ~0ull;
        if (i < n) {
            const unsigned long long k = S.K[i];
            const unsigned key = (unsigned)(k >> 32);
            S.orig[i] = (uint16_t)(unsigned)k;
            k2 = (unsigned long long)((key & 8191u) << 14 | (key >> 13)) << 32 | (unsigned)i;
        }
        S.K[i] = k2;  // every thread rewrites only the slot it has just read
    }
    bitonic_sort(S.K, np2);
    int ix = 0, iy = 1, isp = 2;  // which of the three buffers holds the x-ordered list, the y-ordered list, the spare
    for (int r = tid; r < n; r += OR_THREADS) {
        const unsigned xr = (unsigned)S.K[r];
        const uint32_t e = xr << 16 | (unsigned)r;
        S.L[iy][r] = e;
        S.L[ix][xr] = e;
        S.node_start[r] = 0;
        S.node_size[r] = (uint16_t)n;
    }
    __syncthreads();

    // 3. alternating cuts.  Thread t owns the positions [t * per, (t + 1) * per).
    const int per = (n + OR_THREADS - 1) / OR_THREADS;  // 1 .. 4
    const int p0 = tid * per;
    const int lane = tid & 31, wid = tid >> 5;
    for (int axis = 0;; axis ^= 1) {
        const uint32_t *cut = S.L[axis == 0 ? ix : iy];
        const uint32_t *mov = S.L[axis == 0 ? iy : ix];
        uint32_t *dst = S.L[isp];
        uint32_t e[4];
        int s[4], sz[4];
        bool low[4];
        int mine = 0;
        bool any_active = false;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int p = p0 + q;
            low[q] = false;
            s[q] = sz[q] = 0;
            e[q] = 0;
            if (q < per && p < n) {
                s[q] = S.node_start[p];
                sz[q] = S.node_size[p];
                e[q] = mov[p];
                if (sz[q] >= 4) {
                    any_active = true;
                    const uint32_t pv = cut[s[q] + (sz[q] >> 1)];
                    low[q] = axis == 0 ? (e[q] >> 16) < (pv >> 16) : (e[q] & 0xFFFFu) < (pv & 0xFFFFu);
                }
                mine += low[q] ? 1 : 0;
            }
        }
        if (!__syncthreads_or(any_active)) break;
        // block-wide exclusive prefix sum of the flags
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) S.warp_total[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = S.warp_total[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (lane >= o) w += t;
            }
            S.warp_total[lane] = w;  // inclusive
        }
        __syncthreads();
        int run = (wid > 0 ? S.warp_total[wid - 1] : 0) + incl - mine;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int p = p0 + q;
            if (q < per && p < n) {
                S.prefix[p] = (uint16_t)run;
                run += low[q] ? 1 : 0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int p = p0 + q;
            if (q < per && p < n) {
                if (sz[q] >= 4) {
                    const int div = sz[q] >> 1;
                    const int lowrank = (int)S.prefix[p] - (int)S.prefix[s[q]];  // low elements of this node in front of p
                    const int np = low[q] ? s[q] + lowrank : s[q] + div + (p - s[q] - lowrank);
                    dst[np] = e[q];
                    if (p < s[q] + div) {
                        S.node_size[p] = (uint16_t)div;
                    } else {
                        S.node_start[p] = (uint16_t)(s[q] + div);
                        S.node_size[p] = (uint16_t)(sz[q] - div);
                    }
                } else {
                    dst[p] = e[q];
                }
            }
        }
        __syncthreads();
        // the freshly written buffer becomes the list that was moved; its old buffer is the new spare
        if (axis == 0) {
            const int t = iy;
            iy = isp;
            isp = t;
        } else {
            const int t = ix;
            ix = isp;
            isp = t;
        }
    }
    int32_t *order_dev = d_order_all ? d_order_all + ((size_t)f * 2 + side) * maxS : nullptr;
    for (int i = tid; i < n; i += OR_THREADS) {
        const int32_t id = (int32_t)S.orig[S.L[ix][i] >> 16];
        order[i] = id;
        if (order_dev) order_dev[i] = id;
    }
    if (tid == 0) {
        *ok_out = n;  // usable: the number of vertices in `order` (all of them, or the survivors of the duplicate removal)
        if (ok_dev) *ok_dev = n;
    }
}

}  // namespace

size_t delaunay_dup_keys_per_list() { return OR_MAXN; }

// dup_count [nf][2] int32 and dup_keys [nf][2][delaunay_dup_keys_per_list()] uint64: device scratch for lists with duplicate coordinates;
// both null: such lists are flagged for the host instead (h_ok = 0)
int launch_delaunay_order(const Dims &d, const int32_t *support, const int32_t *nsupport, int32_t *h_order, int32_t *h_ok, int32_t *d_order,
                          int32_t *d_ok, int32_t *dup_count, unsigned long long *dup_keys, int nf, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        SVB_CUDA(cudaFuncSetAttribute(k_delaunay_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OrderSmem)));
        configured[dev] = true;
    }
    k_delaunay_order<<<dim3(nf, 2), OR_THREADS, sizeof(OrderSmem), s>>>(support, nsupport, h_order, h_ok, d_order, d_ok, d.maxS, 0, dup_count, dup_keys);
    SVB_LAUNCH_CHECK();
    if (dup_count && dup_keys) {
        const size_t smem = sizeof(unsigned long long) * OR_MAXN + sizeof(uint32_t) * RS_STACK;  // 40 KB
        k_replay_vertexsort<<<dim3(nf, 2), 32, smem, s>>>(support, nsupport, dup_count, dup_keys, d.maxS);
        SVB_LAUNCH_CHECK();
        k_delaunay_order<<<dim3(nf, 2), OR_THREADS, sizeof(OrderSmem), s>>>(support, nsupport, h_order, h_ok, d_order, d_ok, d.maxS, 1, dup_count, dup_keys);
        SVB_LAUNCH_CHECK();
    }
    return SVB_OK;
}

}  // namespace svb
