// removeSmallSegments as GPU connected components.
//
// Replaces Elas::removeSmallSegments (src/serial_includes/elas/elas.cpp:1013-1124): 4-connected components over
// the relation "both pixels valid (>= 0) and |D(p) - D(q)| <= speckle_sim_threshold"; every component with fewer
// than speckle_size pixels is set to -10.  The reference grows segments breadth-first from scan-order seeds, but
// the relation is symmetric, so the partition (and therefore the result) does not depend on the order.
// (Invalid pixels are -10 after the L/R check, |(-10) - d| >= 10 > threshold, so they never join a segment.)
//
// Algorithm: two-level lock-free union-find (roots are the minimum index of a component).
//   1. tile    : a CTA labels a 128 x 16 pixel tile entirely in shared memory (run starts by warp ballot, vertical
//                edges with atomicMin hooks), counts the sizes of the tile-local components and writes, per pixel, the
//                global index of its LOCAL root; local roots start as their own parents in the global forest;
//   2. borders : only the edges that cross tile borders (1/16 of the vertical, 1/128 of the horizontal ones) are united
//                in global memory, and only between local roots -- a few thousand nodes instead of W*H;
//   3. totals  : every local root adds its size to its global root.  The tile kernel leaves a compact list of its local roots (a few
//                per tile) and writes sizes only AT roots, so this step touches a few thousand entries instead of scanning two W*H maps;
//   4. prune   : pixels whose global root counts < speckle_size become -10 (step 3 leaves every tile-local root pointing straight at
//                its root, so a pixel gets there in one hop; pixels of a component that is large enough inside its tile carry CCL_KEPT
//                and need no look-up at all).  In the batch pipeline this step rides on the row pass of the gap interpolation
//                (k_gap_rows, k_post.cu); k_ccl_prune serves the staged calls, tap mode and two-map post-processing.
#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ bool similar(float a, float b, float thr) { return a >= 0.f && b >= 0.f && fabsf(__fsub_rn(a, b)) <= thr; }

__device__ __forceinline__ int find_root(const int32_t *labels, int i) {
    SVB_GUARD_ASSERT(i >= 0);
    int p = labels[i];
    while (p != i) {
        SVB_GUARD_ASSERT(p >= 0 && p < i);  // labels only ever point towards smaller indices
        i = p;
        p = labels[i];
    }
    return i;
}

__device__ void unite(int32_t *labels, int a, int b) {
    while (true) {
        a = find_root(labels, a);
        b = find_root(labels, b);
        if (a == b) return;
        if (a > b) {
            int t = a;
            a = b;
            b = t;
        }
        // hook the larger root under the smaller one
        const int old = atomicMin(labels + b, a);
        if (old == b) return;
        b = old;
    }
}

// Every per-pixel kernel below lets a thread walk RPB consecutive rows of its column: 8x fewer, 8x longer CTAs than
// one pixel per thread (the one-pixel form was bound by CTA launch rate and exposed load latency).
constexpr int RPB = 8;
constexpr int TW = 128, TH = 2 * RPB;  // tile: 128 columns x 16 rows, 256 threads (thread = one column, 8 rows)

// grid: (ceil(W/TW), ceil(H/TH), nimg).  labels: -1 for invalid pixels, else the global index (v*W + u) of the pixel's
// tile-local root; sizes = size of the tile-local component at its local root, 0 everywhere else.
__global__ void __launch_bounds__(2 * TW) k_ccl_tile(const float *__restrict__ D_all, int32_t *__restrict__ labels_all,
                                                    int32_t *__restrict__ sizes_all, int32_t *__restrict__ roots_all, int32_t *__restrict__ counts_all,
                                                    int W, int H, float thr, int min_size) {
    __shared__ float sD[TH * TW];
    __shared__ int sL[TH * TW];
    __shared__ int sS[TH * TW];
    __shared__ int s_nroots;
    if (threadIdx.x == 0) s_nroots = 0;
    const size_t img = (size_t)blockIdx.z * (unsigned)(W * H);
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int c = threadIdx.x & (TW - 1), r0 = (threadIdx.x >> 7) * RPB, lane = threadIdx.x & 31;
    const int u = x0 + c;
#pragma unroll
    for (int k = 0; k < RPB; k++) {
        const int r = r0 + k, v = y0 + r;
        sD[r * TW + c] = (u < W && v < H) ? D_all[img + (unsigned)(v * W + u)] : -10.f;
        sS[r * TW + c] = 0;
    }
    __syncthreads();
    // every pixel links to the leftmost pixel of its horizontal run inside the warp's 32 columns (ballot, no atomics);
    // a run that continues into the previous warp links to the pixel just left of this warp
#pragma unroll
    for (int k = 0; k < RPB; k++) {
        const int idx = (r0 + k) * TW + c;
        const float d = sD[idx];
        const bool link = c > 0 && similar(d, sD[idx - 1], thr);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, link);
        const unsigned clear_below = ~bal & ((2u << lane) - 1u);  // lanes <= lane with no left link (lane 31: all bits)
        const unsigned mask = (lane == 31) ? ~bal : clear_below;
        const int start_lane = mask ? (31 - __clz(mask)) : -1;
        int label = -1;  // invalid pixels never join anything (|-10 - x| > thr) and are pruned unconditionally
        if (d >= 0.f) label = start_lane >= 0 ? idx - (lane - start_lane) : idx - lane - 1;
        sL[idx] = label;
    }
    __syncthreads();
    // vertical edges inside the tile; an edge is skipped when the edge one column to the left unites the same two runs
#pragma unroll
    for (int k = 0; k < RPB; k++) {
        const int r = r0 + k, idx = r * TW + c;
        if (r == 0) continue;
        const float d = sD[idx], du = sD[idx - TW];
        if (!similar(d, du, thr)) continue;
        if (c > 0) {
            const float dl = sD[idx - 1], dul = sD[idx - TW - 1];
            if (similar(d, dl, thr) && similar(du, dul, thr) && similar(dl, dul, thr)) continue;
        }
        SVB_GUARD_ASSERT(idx - TW >= 0 && idx < TH * TW && sL[idx] >= 0 && sL[idx - TW] >= 0);
        unite(sL, idx, idx - TW);
    }
    __syncthreads();
    // flatten; the tile-local sizes are counted per run segment (the part of a horizontal run inside one warp's 32 columns): its first
    // lane adds the segment's length to the root
    int root[RPB];
#pragma unroll
    for (int k = 0; k < RPB; k++) {
        const int idx = (r0 + k) * TW + c;
        const int l = sL[idx];
        root[k] = l >= 0 ? find_root(sL, l) : -1;
        // segment starts: valid lanes whose label is not shared with the lane to the left
        const int left = __shfl_up_sync(0xFFFFFFFFu, l, 1);
        const bool start = l >= 0 && (lane == 0 || left != l);
        const unsigned valid = __ballot_sync(0xFFFFFFFFu, l >= 0), starts = __ballot_sync(0xFFFFFFFFu, start);
        if (start) {
            // the segment ends before the next start or the next invalid lane
            const unsigned stop = (starts | ~valid) & (lane == 31 ? 0u : (0xFFFFFFFFu << (lane + 1)));
            const int end = stop ? __ffs(stop) - 1 : 32;
            atomicAdd(&sS[root[k]], end - lane);
        }
    }
    __syncthreads();
    int32_t *labels = labels_all + img;
    int32_t *sizes = sizes_all + img;
    // this tile's list of local roots: room for every pixel of the tile (worst case: no two neighbours similar)
    const int tile_id = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    int32_t *roots = roots_all + (size_t)tile_id * (TW * TH);
#pragma unroll
    for (int k = 0; k < RPB; k++) {
        const int r = r0 + k, v = y0 + r, idx = r * TW + c;
        if (u >= W || v >= H) continue;
        const unsigned g = (unsigned)(v * W + u);
        int label = -1;
        if (root[k] >= 0) {
            const int rr = root[k] >> 7, rc = root[k] & (TW - 1);
            label = (y0 + rr) * W + x0 + rc;
            // a component that is large enough inside this tile alone is kept whatever it joins across the borders: its pixels say so
            // themselves, and the pruning pass skips them without a look-up (root entries are parent pointers and stay plain)
            if (root[k] != idx && sS[root[k]] >= min_size) label |= CCL_KEPT;
        }
        labels[g] = label;
        if (root[k] == idx) {  // a tile-local root: its size is only ever read at this index
            sizes[g] = sS[idx];
            roots[atomicAdd(&s_nroots, 1)] = (int32_t)g;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) counts_all[tile_id] = s_nroots;
}

// Edges across tile borders: rows v = TH, 2 TH, ... (against v - 1) and columns u = TW, 2 TW, ... (against u - 1).
// grid: (ceil((nrow*W + ncol*H) / 256), nimg).  A horizontal-border edge is skipped when the edge one column to the left
// (same border, same tile column) unites the same two components.
__global__ void __launch_bounds__(256) k_ccl_borders(const float *__restrict__ D_all, int32_t *__restrict__ labels_all, int W, int H, float thr) {
    const size_t img = (size_t)blockIdx.y * (unsigned)(W * H);
    const float *D = D_all + img;
    int32_t *labels = labels_all + img;
    const int nrow = (H - 1) / TH, ncol = (W - 1) / TW;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int p, q;
    if (i < nrow * W) {
        const int k = i / W, u = i - k * W;
        p = (k + 1) * TH * W + u;
        q = p - W;
        const float d = D[p], du = D[q];
        if (!similar(d, du, thr)) return;
        if ((u & (TW - 1)) != 0) {
            const float dl = D[p - 1], dul = D[q - 1];
            if (similar(d, dl, thr) && similar(du, dul, thr) && similar(dl, dul, thr)) return;
        }
    } else if (i < nrow * W + ncol * H) {
        const int j = i - nrow * W;
        const int k = j / H, v = j - k * H;
        p = v * W + (k + 1) * TW;
        q = p - 1;
        if (!similar(D[p], D[q], thr)) return;
    } else {
        return;
    }
    SVB_GUARD_ASSERT(p >= 0 && p < W * H && q >= 0 && q < W * H && labels[p] >= 0 && labels[q] >= 0 && (labels[p] & ~CCL_KEPT) < W * H);
    unite(labels, labels[p] & ~CCL_KEPT, labels[q] & ~CCL_KEPT);  // both pixels are valid: their labels are local roots
}

// Every tile-local root adds its size to the root of its component: one warp per tile walks the tile's root list.
// grid: (ceil(tiles / 8), nimg), 256 threads; tiles = tiles of ONE image
__global__ void __launch_bounds__(256) k_ccl_totals(int32_t *__restrict__ labels_all, int32_t *__restrict__ sizes_all,
                                                   const int32_t *__restrict__ roots_all, const int32_t *__restrict__ counts_all, int tiles, int W, int H) {
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= tiles) return;
    const size_t img = (size_t)blockIdx.y * (unsigned)(W * H);
    int32_t *labels = labels_all + img;
    int32_t *sizes = sizes_all + img;
    const int tile_id = blockIdx.y * tiles + t;
    const int32_t *roots = roots_all + (size_t)tile_id * (TW * TH);
    const int n = counts_all[tile_id];
    for (int i = lane; i < n; i += 32) {
        const int g = roots[i];
        if (labels[g] == g) continue;  // also the root of its whole component: its size entry is the accumulator
        const int root = find_root(labels, g);
        atomicAdd(sizes + root, sizes[g]);  // nobody else touches sizes[g]: g is not a global root
        // path compression: from here on every tile-local root points straight at the root of its component, so a pixel finds it in
        // ONE hop (pixel -> local root -> global root).  Racing walkers read either the old parent or the root: both are ancestors.
        labels[g] = root;
    }
}

__global__ void __launch_bounds__(128) k_ccl_prune(float *__restrict__ D_all, const int32_t *__restrict__ labels_all,
                                                  const int32_t *__restrict__ sizes_all, int W, int H, int min_size) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= W) return;
    const size_t img = (size_t)blockIdx.z * (unsigned)(W * H);
    const int32_t *labels = labels_all + img;
    const int v_end = min((int)(blockIdx.y + 1) * RPB, H);
#pragma unroll 4
    for (int v = blockIdx.y * RPB; v < v_end; v++) {
        const int idx = v * W + u;
        const int r = labels[idx];
        if (r < 0) {
            if (1 < min_size) D_all[img + (unsigned)idx] = -10.f;
        } else if (!(r & CCL_KEPT) && sizes_all[img + (unsigned)labels[r]] < min_size) {  // r is a tile-local root; k_ccl_totals compressed its path
            D_all[img + (unsigned)idx] = -10.f;
        }
    }
}

}  // namespace

int ccl_tiles_per_image(const Dims &d) { return ((d.Dw + TW - 1) / TW) * ((d.Dh + TH - 1) / TH); }

int ccl_min_size(const Dims &d, const svb_params &p) {
    // elas.cpp:1017-1022: at half resolution a speckle is sqrt(speckle_size) * 2 pixels
    return d.sub ? (int)(sqrtf((float)p.speckle_size) * 2) : p.speckle_size;
}

// labels / sizes: one int per pixel and image; roots: one int per pixel and image (per-tile root lists); counts: one int per tile
int launch_ccl_label(const Dims &d, const svb_params &p, const float *D, int32_t *labels, int32_t *sizes, int32_t *roots, int32_t *counts, int nimg,
                     cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    const int W = d.Dw, H = d.Dh;
    dim3 tiles((W + TW - 1) / TW, (H + TH - 1) / TH, nimg);
    k_ccl_tile<<<tiles, 2 * TW, 0, s>>>(D, labels, sizes, roots, counts, W, H, p.speckle_sim_threshold, ccl_min_size(d, p));
    SVB_LAUNCH_CHECK();
    const int border_edges = ((H - 1) / TH) * W + ((W - 1) / TW) * H;
    if (border_edges > 0) {
        k_ccl_borders<<<dim3((border_edges + 255) / 256, nimg), 256, 0, s>>>(D, labels, W, H, p.speckle_sim_threshold);
        SVB_LAUNCH_CHECK();
    }
    const int per_image = (int)(tiles.x * tiles.y);
    k_ccl_totals<<<dim3((per_image + 7) / 8, nimg), 256, 0, s>>>(labels, sizes, roots, counts, per_image, W, H);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_ccl_prune(const Dims &d, const svb_params &p, float *D, const int32_t *labels, const int32_t *sizes, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    const int W = d.Dw, H = d.Dh;
    dim3 grid((W + 127) / 128, (H + RPB - 1) / RPB, nimg);
    k_ccl_prune<<<grid, 128, 0, s>>>(D, labels, sizes, W, H, ccl_min_size(d, p));
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_remove_small_segments(const Dims &d, const svb_params &p, float *D, int32_t *labels, int32_t *sizes, int32_t *roots, int32_t *counts,
                                 int nimg, cudaStream_t s) {
    SVB_TRY(launch_ccl_label(d, p, D, labels, sizes, roots, counts, nimg, s));
    return launch_ccl_prune(d, p, D, labels, sizes, nimg, s);
}

}  // namespace svb
