// removeSmallSegments as GPU connected components.
//
// Replaces Elas::removeSmallSegments (src/serial_includes/elas/elas.cpp:1013-1124): 4-connected components over
// the relation "both pixels valid (>= 0) and |D(p) - D(q)| <= speckle_sim_threshold"; every component with fewer
// than speckle_size pixels is set to -10.  The reference grows segments breadth-first from scan-order seeds, but
// the relation is symmetric, so the partition (and therefore the result) does not depend on the order.
// (Invalid pixels are -10 after the L/R check, |(-10) - d| >= 10 > threshold, so they never join a segment.)
//
// Algorithm: lock-free union-find over pixel indices (roots are the minimum index of a component).
//   1. init   : every pixel links to the leftmost pixel of its horizontal run (warp ballot, no atomics);
//   2. merge  : vertical edges and run boundaries are united with atomicMin hooks;
//   3. count  : each pixel finds its root and adds 1 to the root's counter;
//   4. prune  : pixels whose root counts < speckle_size become -10.
#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ bool similar(float a, float b, float thr) { return a >= 0.f && b >= 0.f && fabsf(__fsub_rn(a, b)) <= thr; }

__device__ __forceinline__ int find_root(const int32_t *labels, int i) {
    int p = labels[i];
    while (p != i) {
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
// one pixel per thread (the one-pixel form was bound by CTA launch rate and exposed load latency: ~1.2 TB/s effective).
constexpr int RPB = 8;

// grid: (ceil(W/128), ceil(H/RPB), nimg)   labels are indices local to the image; invalid pixels get label -1
__global__ void __launch_bounds__(128) k_ccl_init(const float *__restrict__ D_all, int32_t *__restrict__ labels_all, int32_t *__restrict__ sizes_all,
                                                 int W, int H, float thr) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t img = (size_t)blockIdx.z * W * H;
    const int lane = threadIdx.x & 31;
    const bool in = u < W;
    const int v_end = min((int)(blockIdx.y + 1) * RPB, H);
#pragma unroll 4
  for (int v = blockIdx.y * RPB; v < v_end; v++) {
    const int idx = v * W + u;
    const float d = in ? D_all[img + idx] : -10.f;
    const float dl = __shfl_up_sync(0xFFFFFFFFu, d, 1);
    const float dleft = (lane == 0) ? ((in && u > 0) ? D_all[img + idx - 1] : -10.f) : dl;
    // linked to the left neighbour?
    const bool link = in && u > 0 && similar(d, dleft, thr);
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, link);
    if (!in) continue;
    // run start inside this warp: the nearest lane at or below `lane` whose link bit is clear
    const unsigned clear_below = ~bal & ((2u << lane) - 1u);  // lanes <= lane with no left link (lane 31: all bits)
    const unsigned mask = (lane == 31) ? ~bal : clear_below;
    const int start_lane = mask ? (31 - __clz(mask)) : -1;
    int label;
    if (d < 0.f) {
        label = -1;  // never joins anything (|-10 - x| > thr); pruned unconditionally (a segment of one pixel)
    } else if (start_lane >= 0) {
        label = idx - (lane - start_lane);
    } else {
        // the run continues into the previous warp: link to the pixel just left of this warp's first lane
        label = idx - lane - 1;
    }
    labels_all[img + idx] = label;
    if (label == idx) sizes_all[img + idx] = 0;  // only run starts can end up as roots (a root is its segment's minimum index)
  }
}

// Vertical edges.  An edge (u,v)-(u,v-1) is skipped when the edge one column to the left already unites the same two
// runs: both pixels are linked to their left neighbours and those neighbours are vertically similar.
__global__ void __launch_bounds__(128) k_ccl_merge(const float *__restrict__ D_all, int32_t *__restrict__ labels_all, int W, int H, float thr) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= W) return;
    const size_t img = (size_t)blockIdx.z * W * H;
    const int v0 = max((int)blockIdx.y * RPB, 1), v_end = min((int)(blockIdx.y + 1) * RPB, H);
    if (v0 >= v_end) return;
    // the row above is carried in registers from one step to the next
    float du = D_all[img + (size_t)(v0 - 1) * W + u];
    float dul = u > 0 ? D_all[img + (size_t)(v0 - 1) * W + u - 1] : -10.f;
    for (int v = v0; v < v_end; v++) {
        const int idx = v * W + u;
        const float d = D_all[img + idx];
        const float dl = u > 0 ? D_all[img + idx - 1] : -10.f;
        if (similar(d, du, thr) && !(u > 0 && similar(d, dl, thr) && similar(du, dul, thr) && similar(dl, dul, thr)))
            unite(labels_all + img, idx, idx - W);
        du = d;
        dul = dl;
    }
}

// Flatten labels to roots and count segment sizes.  A CTA covers a 128 x 8 pixel tile; equal roots are first combined
// inside each warp (match.any), then across the CTA in a small shared-memory hash table, so a large segment costs one
// global atomic per tile instead of one per pixel.
constexpr int CNT_ROWS = 8;
constexpr int CNT_SLOTS = 64;

__global__ void __launch_bounds__(32 * CNT_ROWS) k_ccl_count(int32_t *__restrict__ labels_all, int32_t *__restrict__ sizes_all, int W, int H) {
    __shared__ int s_key[CNT_SLOTS];
    __shared__ int s_val[CNT_SLOTS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < CNT_SLOTS) {
        s_key[threadIdx.x] = -1;
        s_val[threadIdx.x] = 0;
    }
    __syncthreads();
    const size_t img = (size_t)blockIdx.z * W * H;
    int32_t *labels = labels_all + img;
    int32_t *sizes = sizes_all + img;
    const int v = blockIdx.y * CNT_ROWS + wid;
    if (v < H) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int u = blockIdx.x * 128 + k * 32 + lane;
            int r = -1;
            if (u < W) {
                const int idx = v * W + u;
                const int l = labels[idx];
                if (l >= 0) {
                    r = find_root(labels, l);
                    if (r != l) labels[idx] = r;  // roots are fixed once merging has finished
                }
            }
            const unsigned act = __ballot_sync(0xFFFFFFFFu, r >= 0);
            if (r >= 0) {
                const unsigned same = __match_any_sync(act, r);
                if (lane == __ffs(same) - 1) {
                    const int cnt = __popc(same);
                    unsigned h = ((unsigned)r * 2654435761u) >> 26;
                    bool done = false;
#pragma unroll 1
                    for (int t = 0; t < 4 && !done; t++) {
                        const int old = atomicCAS(&s_key[h], -1, r);
                        if (old == -1 || old == r) {
                            atomicAdd(&s_val[h], cnt);
                            done = true;
                        }
                        h = (h + 1) & (CNT_SLOTS - 1);
                    }
                    if (!done) atomicAdd(sizes + r, cnt);
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < CNT_SLOTS && s_key[threadIdx.x] >= 0) atomicAdd(sizes + s_key[threadIdx.x], s_val[threadIdx.x]);
}

__global__ void __launch_bounds__(128) k_ccl_prune(float *__restrict__ D_all, const int32_t *__restrict__ labels_all,
                                                  const int32_t *__restrict__ sizes_all, int W, int H, int min_size) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= W) return;
    const size_t img = (size_t)blockIdx.z * W * H;
    const int v_end = min((int)(blockIdx.y + 1) * RPB, H);
#pragma unroll 4
    for (int v = blockIdx.y * RPB; v < v_end; v++) {
        const int idx = v * W + u;
        const int r = labels_all[img + idx];
        if (r < 0) {
            if (1 < min_size) D_all[img + idx] = -10.f;
        } else if (sizes_all[img + r] < min_size) {
            D_all[img + idx] = -10.f;
        }
    }
}

}  // namespace

int launch_remove_small_segments(const Dims &d, const svb_params &p, float *D, int32_t *labels, int32_t *sizes, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    const int W = d.Dw, H = d.Dh;
    // elas.cpp:1017-1022: at half resolution a speckle is sqrt(speckle_size) * 2 pixels
    const int min_size = d.sub ? (int)(sqrtf((float)p.speckle_size) * 2) : p.speckle_size;
    dim3 grid((W + 127) / 128, (H + RPB - 1) / RPB, nimg);
    k_ccl_init<<<grid, 128, 0, s>>>(D, labels, sizes, W, H, p.speckle_sim_threshold);
    SVB_LAUNCH_CHECK();
    if (H > 1) {
        k_ccl_merge<<<grid, 128, 0, s>>>(D, labels, W, H, p.speckle_sim_threshold);
        SVB_LAUNCH_CHECK();
    }
    dim3 gc((W + 127) / 128, (H + CNT_ROWS - 1) / CNT_ROWS, nimg);
    k_ccl_count<<<gc, 32 * CNT_ROWS, 0, s>>>(labels, sizes, W, H);
    SVB_LAUNCH_CHECK();
    k_ccl_prune<<<grid, 128, 0, s>>>(D, labels, sizes, W, H, min_size);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
