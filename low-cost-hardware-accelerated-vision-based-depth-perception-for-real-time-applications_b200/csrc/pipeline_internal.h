// Internal structures of the pipeline, shared by pipeline.cu and band_split.cu.  Nothing here crosses the C-ABI.
#pragma once

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "host_delaunay.h"
#include "svb_internal.h"
#include "thread_pool.h"

namespace svb {

constexpr int MAX_LANES = 8;  // arenas in flight; the context uses n_lanes of them (default 8, SVB_LANES overrides)

struct Lane {
    cudaStream_t own_stream = nullptr;  // created with the lane
    cudaStream_t stream = nullptr;      // the stream the lane's work is issued on (own_stream, or lane 0's in single-stream mode)
    cudaEvent_t ev_a = nullptr, ev_done = nullptr;
    // device
    uint8_t *img[2] = {nullptr, nullptr};        // [chunk][H][bpl]
    uint8_t *img_tight[2] = {nullptr, nullptr};  // [chunk][H][W] landing buffer of the host-buffer batch path (allocated on first use)
    uint8_t *desc[2] = {nullptr, nullptr};       // = desc_base + padding: the dense matcher's unguarded loads may run a few
    uint8_t *desc_base[2] = {nullptr, nullptr};  //   hundred descriptors past either end of the arena (k_dense.cu)
    int16_t *dcan_raw = nullptr, *dcan = nullptr;
    int32_t *support = nullptr, *nsupport = nullptr;
    int32_t *sf_changed = nullptr;  // per-sweep flags of the multi-CTA lattice filter (k_support.cu)
    int32_t *tri[2] = {nullptr, nullptr};
    int32_t *ntri = nullptr;    // [2*chunk] triangle counts (2f + side), then [chunk] first triangle of frame f in tri[]
    int32_t *trioff = nullptr;  // = ntri + 2*chunk
    PlaneRec *rec[2] = {nullptr, nullptr};
    uint32_t *grid_tmp = nullptr, *grid[2] = {nullptr, nullptr};
    int32_t *owner[2] = {nullptr, nullptr};
    float *Draw = nullptr;  // [2][chunk][N]
    float *Dlr = nullptr;   // [2][chunk][N]
    float *Dtmp = nullptr;  // [2][chunk][N]
    int32_t *labels = nullptr, *sizes = nullptr;  // [2][chunk][N]
    int owner_gen = 0;  // generation of the owner maps' entries (stage_b)
    int32_t *ccl_roots = nullptr, *ccl_counts = nullptr;  // per-tile lists of tile-local roots and their lengths (k_ccl.cu)
    uint8_t *dmap = nullptr;
    // pinned host
    int32_t *h_support = nullptr, *h_nsupport = nullptr, *h_tri[2] = {nullptr, nullptr}, *h_ntri = nullptr;
    int32_t *h_order = nullptr;     // [chunk][2][maxS] recursion order of the Delaunay stage (k_order.cu), written by the device
    int32_t *h_order_ok = nullptr;  // [chunk][2] 1 = h_order is valid for that list
    int32_t *d_order = nullptr;     // device copy of h_order for the device divide-and-conquer (k_delaunay.cu)
    int32_t *d_order_ok = nullptr;  // device copy of h_order_ok
    int32_t *dup_count = nullptr;   // [chunk][2] k_order.cu: -1 = list flagged for the vertex-sort replay, then the number of distinct vertices
    unsigned long long *dup_keys = nullptr;  // [chunk][2][4096] sorted, de-duplicated keys of the flagged lists
    int32_t *h_dd_done = nullptr;   // [chunk][2] mapped host: 1 = the device triangulated that list (its count is in h_ntri), 0 = host's turn
    std::vector<uint8_t> host_made; // [chunk][2] this chunk's lists that the host stage produced (they need an H2D copy)
    bool unpacked = false;          // layout of this chunk's triangle lists: frame f at f * (maxT + 8) (device stage) or packed back to back
    int first_frame = -1;           // batch paths: index of the chunk's first frame in the batch (per-frame status), else -1
};

// CUDA events bracketing every stage of one chunk (stage timing): [i] is recorded in front of stage i, [ST_COUNT] after
// the last stage, a_end behind the D2H that ends stage A.
struct StageEvents {
    cudaEvent_t ev[ST_COUNT + 1] = {};
    cudaEvent_t a_end = nullptr;
    bool a_done = false, b_done = false;
};

struct Tap {
    std::string name;
    void *dev = nullptr;
    size_t bytes = 0;
};

}  // namespace svb

struct svb_context {
    svb_params p;
    svb::Dims d;
    int chunk = 1;
    int device = 0;
    int mean_mode = SVB_MEAN_SERIAL_QUANTISED;
    svb::Lane lanes[svb::MAX_LANES];
    int n_lanes = 3;
    std::unique_ptr<svb::ThreadPool> pool;
    std::vector<svb::DelaunayScratch> scratch;
    svb::Calib calib;
    bool have_calib = false;
    // tap mode (single frame)
    bool tap_mode = false;
    std::vector<svb::Tap> taps;
    float *planes_ref[2] = {nullptr, nullptr};  // [maxT][6], tap mode only
    std::vector<int32_t> inject_tri[2];
    bool inject[2] = {false, false};
    // generatePointCloud path: BGRA staging on the device
    uint8_t *bgra[2] = {nullptr, nullptr};
    cudaEvent_t ev_pc[4] = {nullptr, nullptr, nullptr, nullptr};
    uint8_t *bgra_batch[2] = {nullptr, nullptr};  // svb_batch_upload_bgra: staging for `chunk` BGRA frames per side
    size_t bgra_batch_frames = 0;
    bool last_want_D = false, last_want_P = false;  // outputs of the last batch call (svb_batch_device_ptrs)
    // resident batch stores
    uint8_t *in_img[2] = {nullptr, nullptr};
    size_t in_frames = 0;
    float *out_D1 = nullptr;
    size_t out_D1_frames = 0;
    double *out_points = nullptr;
    size_t out_points_frames = 0;
    std::vector<int32_t> frame_nsupport;  // support points of every frame of the last batch call (< 3: the frame failed)
    // stats
    svb_stats stats;
    bool stage_timing = false;
    bool single_stream = false;
    bool fused_post = true;         // SVB_FUSED_POST=0: stage-by-stage mean / median / reproject kernels everywhere (k_post_fused.cu otherwise)
    bool points_float_disp = false; // SVB_OUT_POINTS_FLOATDISP of the call in flight
    bool delaunay_device = true;    // SVB_DELAUNAY_DEVICE=0: the divide-and-conquer runs on the host for every list (k_delaunay.cu otherwise)
    int dups_policy = 0;            // lists with duplicate coordinates: 0 = by host thread count, 1 = device, 2 = host (SVB_DELAUNAY_DUPS)
    int dd_cap = 2048;              // vertex capacity the device divide-and-conquer is launched with (shared memory); follows the lists seen
    int host_threads = 1;           // hardware threads of the host (a large list spreads its subtrees over those the list-level pool leaves idle)
    bool gpu_order = true;  // SVB_GPU_ORDER=0: the host stage sorts and partitions the vertices itself
    std::vector<svb::StageEvents> stage_ev;  // one set per chunk of the call in flight
    std::mutex mu;
};


namespace svb {
// host Delaunay stage of one chunk whose support lists have arrived in lane L's pinned buffers (pipeline.cu)
// unpacked: the layout the device divide-and-conquer writes (frame f's lists at triangle f * (maxT + 8)); only lists the device left
// (h_dd_done = 0) or that are injected are made here.  Packed: every list is made here, back to back (one H2D copy per side).
int stage_host(svb_context *c, Lane &L, int nf, bool unpacked = false);
}  // namespace svb
