// Vertex order for the host Delaunay stage, computed on the device.
//
// The reference's triangulator (Triangle 1.6, src/common_includes/elas/triangle.cpp) starts with two data-parallel
// steps before its inherently sequential divide-and-conquer: the lexicographic (x, y) sort of the vertices
// (triangle.cpp:5882-5903) and the alternating-axis median partition (:5243-5325, :5904-5913).  For duplicate-free
// input their result -- the order in which the recursion meets the vertices -- is unique (host_delaunay.cpp explains
// why only the SETS on either side of each median matter), so it is produced here, one CTA per (frame, image side),
// right behind the support-list compaction, and the host stage starts directly with the recursion (about 15 % less
// host time per frame).  Anything the kernel does not handle -- fewer than 3 or more than 4096 points, coordinates
// outside the packed key range, or DUPLICATE coordinates, where the survivor depends on the reference's randomised
// sort -- is flagged, and the host runs its complete path for that list.
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

// grid: (nf, 2); blockIdx.y = image side (0: (u, v), 1: (u - d, v), elas.cpp:451-461)
__global__ void __launch_bounds__(OR_THREADS) k_delaunay_order(const int32_t *__restrict__ support_all, const int32_t *__restrict__ nsupport_all,
                                                              int32_t *__restrict__ h_order_all, int32_t *__restrict__ h_ok_all,
                                                              int32_t *__restrict__ d_order_all, int32_t *__restrict__ d_ok_all, int maxS) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    OrderSmem &S = *reinterpret_cast<OrderSmem *>(smem_raw);
    const int f = blockIdx.x, side = blockIdx.y, tid = threadIdx.x;
    const int n = nsupport_all[f];
    int32_t *ok_out = h_ok_all + 2 * f + side;
    int32_t *ok_dev = d_ok_all ? d_ok_all + 2 * f + side : nullptr;  // device copies for the divide-and-conquer kernel (k_delaunay.cu)
    if (n < 3 || n > OR_MAXN || n > maxS) {  // uniform
        if (tid == 0) {
            *ok_out = 0;
            if (ok_dev) *ok_dev = 0;
        }
        return;
    }
    const int32_t *support = support_all + (size_t)f * maxS * 3;
    int32_t *order = h_order_all + ((size_t)f * 2 + side) * maxS;
    int np2 = 4;
    while (np2 < n) np2 <<= 1;
    if (tid == 0) S.bad = 0;
    __syncthreads();

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
    for (int i = tid; i < n; i += OR_THREADS)
        if (i > 0 && (S.K[i] >> 32) == (S.K[i - 1] >> 32)) S.bad = 1;  // duplicate coordinates: the host decides which survives
    __syncthreads();
    if (S.bad) {
        if (tid == 0) {
            *ok_out = 0;
            if (ok_dev) *ok_dev = 0;
        }
        return;
    }
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
        *ok_out = 1;
        if (ok_dev) *ok_dev = 1;
    }
}

}  // namespace

int launch_delaunay_order(const Dims &d, const int32_t *support, const int32_t *nsupport, int32_t *h_order, int32_t *h_ok, int32_t *d_order,
                          int32_t *d_ok, int nf, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    SVB_CUDA(cudaFuncSetAttribute(k_delaunay_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(OrderSmem)));  // per device
    k_delaunay_order<<<dim3(nf, 2), OR_THREADS, sizeof(OrderSmem), s>>>(support, nsupport, h_order, h_ok, d_order, d_ok, d.maxS);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
