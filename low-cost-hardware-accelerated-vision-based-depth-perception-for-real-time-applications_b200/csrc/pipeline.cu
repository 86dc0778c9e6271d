// Context, device arenas, the frame-batch pipeline and every svb_* entry point of include/elas_b200.h.
//
// One svb_context owns n_lanes (default 3) independent lanes; a lane holds the device arenas for `chunk` frames, one CUDA
// stream and pinned staging buffers.  A batch is cut into chunks that go round-robin over the lanes:
//     stage A (GPU)  descriptor -> support matching -> lattice filters/compaction -> D2H support list
//     host stage     Delaunay triangulation (worker pool, one job per frame and side)
//     stage B (GPU)  H2D triangles -> planes -> grid -> raster -> dense matching -> L/R check -> speckle removal
//                    -> gap interpolation -> adaptive mean -> median -> u8 conversion + reprojection
// While the host triangulates chunk k, the GPU runs stage B of chunk k-1 and stage A of chunk k+1.
// Frames never interact, so there is no collective anywhere.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "host_delaunay.h"
#include "svb_internal.h"
#include "thread_pool.h"

namespace svb {

static thread_local char g_err[1024] = "";
thread_local long long g_launch_counter = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int make_dims(const svb_params &p, int W, int H, Dims *out) {
    Dims d;
    memset(&d, 0, sizeof(d));
    // 8192: coordinate range of the device vertex order (k_order.cu) and of the host stage's radix sort (host_delaunay.cpp), and
    // every shared-memory line buffer of the post filters fits (4 rows x 8192 px x 4 B)
    if (W < 16 || H < 16 || W > 8192 || H > 8192) {
        set_error("unsupported image size %dx%d (16 .. 8192 per side)", W, H);
        return SVB_ERR_ARG;
    }
    if (p.disp_max < 0 || p.disp_max > 4095 || p.grid_size < 1 || p.candidate_stepsize < 1) {
        set_error("invalid parameters (disp_max=%d grid_size=%d candidate_stepsize=%d)", p.disp_max, p.grid_size, p.candidate_stepsize);
        return SVB_ERR_ARG;
    }
    if (!(p.speckle_sim_threshold < 10.0f)) {
        set_error("speckle_sim_threshold >= 10 would connect invalid (-10) pixels in the reference; unsupported");
        return SVB_ERR_UNSUPPORTED;
    }
    d.W = W;
    d.H = H;
    d.N = W * H;
    d.bpl = W + 15 - (W - 1) % 16;  // elas.cpp:37
    d.IN = (size_t)d.bpl * H;
    d.sub = p.subsampling ? 1 : 0;
    d.Dw = d.sub ? W / 2 : W;
    d.Dh = d.sub ? H / 2 : H;
    d.DN = d.Dw * d.Dh;
    d.step = p.candidate_stepsize;
    if (d.sub) d.step += d.step % 2;  // elas.cpp:376-378: at half resolution only every second descriptor row exists
    d.cw = (W + d.step - 1) / d.step;  // elas.cpp:383-386
    d.ch = (H + d.step - 1) / d.step;
    d.gw = (int)ceil((float)W / (float)p.grid_size);  // elas.cpp:88-89
    d.gh = (int)ceil((float)H / (float)p.grid_size);
    d.gwords = (((p.disp_max + 1 + 31) / 32) + 3) & ~3;  // multiple of 4: cells are read as 16-byte vectors
    d.maxS = (d.cw - 1) * (d.ch - 1) + 6;
    d.maxT = 2 * d.maxS;
    // elas.cpp:828-832
    const float two_sigma_squared = 2 * p.sigma * p.sigma;
    d.plane_radius = (int32_t)fmaxf((float)ceil(p.sigma * p.sradius), (float)2.0);
    if (d.plane_radius > 7) {
        set_error("plane radius %d > 7 unsupported", d.plane_radius);
        return SVB_ERR_UNSUPPORTED;
    }
    for (int delta = 0; delta < 8; delta++)
        d.P[delta] = (int32_t)((-log(p.gamma + exp(-delta * delta / two_sigma_squared)) + log(p.gamma)) / p.beta);
    // The dense matcher packs (cost + bias) << 13 | phase | d into one unsigned key (k_dense.cu); cost = SAD + P[.] is signed in the
    // reference (elas.cpp:672-684) and P <= 0 grows with 1 / beta, so the bias follows the table instead of assuming the presets.
    int minP = 0, maxP = 0;
    for (int delta = 0; delta <= d.plane_radius; delta++) {
        minP = d.P[delta] < minP ? d.P[delta] : minP;
        maxP = d.P[delta] > maxP ? d.P[delta] : maxP;
    }
    d.desc_pad = ((size_t)p.disp_max + W + 1024) * 16;
    d.cost_bias = minP < -16 ? -minP : 16;
    if ((long long)4080 + maxP + d.cost_bias >= (1 << 19) || minP < -(1 << 18)) {
        set_error("prior table out of range for the packed matching key (P in [%d, %d]; gamma=%g beta=%g sigma=%g)", minP, maxP, p.gamma, p.beta,
                  p.sigma);
        return SVB_ERR_UNSUPPORTED;
    }
    *out = d;
    return SVB_OK;
}

static const char *kStageNames[ST_COUNT] = {"descriptor", "support_match", "support_filter", "delaunay_device", "d2h_support", "h2d_triangles",
                                            "planes",     "grid",          "raster",         "dense_match", "lr_check",
                                            "remove_small_segments", "gap_interpolation", "adaptive_mean", "median", "reproject", "post_fused"};

}  // namespace svb

#include "pipeline_internal.h"

using namespace svb;

namespace {

template <typename T>
int dev_alloc(T **p, size_t count) {
    *p = nullptr;
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T) + 256);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    return SVB_OK;
}
template <typename T>
int host_alloc(T **p, size_t count) {
    *p = nullptr;
    cudaError_t e = cudaHostAlloc((void **)p, count * sizeof(T) + 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) {
        set_error("cudaHostAlloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    return SVB_OK;
}

int lane_create(svb_context *c, Lane &L) {
    const Dims &d = c->d;
    const size_t C = (size_t)c->chunk, N = (size_t)d.N;
    SVB_CUDA(cudaStreamCreateWithFlags(&L.own_stream, cudaStreamNonBlocking));
    L.stream = L.own_stream;
    SVB_CUDA(cudaEventCreateWithFlags(&L.ev_a, cudaEventDisableTiming));
    SVB_CUDA(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
    for (int s = 0; s < 2; s++) {
        SVB_TRY(dev_alloc(&L.img[s], C * d.IN));
        SVB_CUDA(cudaMemset(L.img[s], 0, C * d.IN));  // the padding columns are never written again
        {
            // zero-filled guard bands of (disp_max + W + 1024) descriptors in front of and behind the descriptor arena
            const size_t pad = d.desc_pad;
            SVB_TRY(dev_alloc(&L.desc_base[s], C * N * 16 + 2 * pad));
            SVB_CUDA(cudaMemset(L.desc_base[s], 0, C * N * 16 + 2 * pad));
            L.desc[s] = L.desc_base[s] + pad;
        }
        SVB_TRY(dev_alloc(&L.tri[s], C * (d.maxT + 8) * 3));  // + 8: stage_host() reserves 2 n + 8 triangles per frame
        SVB_TRY(dev_alloc(&L.rec[s], C * d.maxT));
        SVB_TRY(dev_alloc(&L.grid[s], C * d.gw * d.gh * d.gwords));
        SVB_TRY(dev_alloc(&L.owner[s], C * N));
        SVB_CUDA(cudaMemset(L.owner[s], 0xFF, C * N * sizeof(int32_t)));  // generation-tagged entries (stage_b): cleared here and when the generation wraps
        SVB_TRY(host_alloc(&L.h_tri[s], C * (d.maxT + 8) * 3));
    }
    SVB_TRY(dev_alloc(&L.dcan_raw, C * d.cw * d.ch));
    SVB_TRY(dev_alloc(&L.dcan, C * d.cw * d.ch + 2));  // + 2: the filters set flag bits with 32-bit atomics on the word holding a cell
    SVB_TRY(dev_alloc(&L.support, C * d.maxS * 3));
    SVB_TRY(dev_alloc(&L.nsupport, C));
    SVB_TRY(dev_alloc(&L.sf_changed, C * SUPPORT_FILTER_SCRATCH_INTS));
    SVB_TRY(dev_alloc(&L.ntri, C * 3));
    L.trioff = L.ntri + 2 * C;
    SVB_TRY(dev_alloc(&L.grid_tmp, C * 2 * d.gw * d.gh * d.gwords));
    SVB_TRY(dev_alloc(&L.Draw, 2 * C * N));
    SVB_TRY(dev_alloc(&L.Dlr, 2 * C * N));
    SVB_CUDA(cudaMemset(L.Dlr, 0, 2 * C * N * sizeof(float)));  // with subsampling only the first DN floats of a map are ever written
    SVB_TRY(dev_alloc(&L.Dtmp, 2 * C * N));
    SVB_TRY(dev_alloc(&L.labels, 2 * C * N));
    SVB_TRY(dev_alloc(&L.sizes, 2 * C * N));
    SVB_TRY(dev_alloc(&L.ccl_roots, 2 * C * (size_t)ccl_tiles_per_image(d) * 2048));  // per-tile root lists: room for every pixel of a 128 x 16 tile
    SVB_TRY(dev_alloc(&L.ccl_counts, 2 * C * (size_t)ccl_tiles_per_image(d)));
    SVB_TRY(dev_alloc(&L.dmap, C * N));
    SVB_TRY(host_alloc(&L.h_support, C * d.maxS * 3));
    SVB_TRY(host_alloc(&L.h_nsupport, C));
    SVB_TRY(host_alloc(&L.h_ntri, C * 3));
    SVB_TRY(host_alloc(&L.h_order, C * 2 * d.maxS));
    SVB_TRY(host_alloc(&L.h_order_ok, C * 2));
    memset(L.h_order_ok, 0, sizeof(int32_t) * C * 2);
    SVB_TRY(dev_alloc(&L.d_order, C * 2 * d.maxS));
    SVB_TRY(dev_alloc(&L.d_order_ok, C * 2));
    SVB_CUDA(cudaMemset(L.d_order_ok, 0, sizeof(int32_t) * C * 2));
    SVB_TRY(dev_alloc(&L.dup_count, C * 2));
    SVB_CUDA(cudaMemset(L.dup_count, 0, sizeof(int32_t) * C * 2));
    SVB_TRY(dev_alloc(&L.dup_keys, C * 2 * delaunay_dup_keys_per_list()));
    SVB_TRY(host_alloc(&L.h_dd_done, C * 2));
    memset(L.h_dd_done, 0, sizeof(int32_t) * C * 2);
    L.host_made.assign(C * 2, 0);
    return SVB_OK;
}

void lane_destroy(Lane &L) {
    for (int s = 0; s < 2; s++) {
        cudaFree(L.img[s]);
        cudaFree(L.img_tight[s]);
        cudaFree(L.desc_base[s]);
        cudaFree(L.tri[s]);
        cudaFree(L.rec[s]);
        cudaFree(L.grid[s]);
        cudaFree(L.owner[s]);
        cudaFreeHost(L.h_tri[s]);
    }
    cudaFree(L.dcan_raw);
    cudaFree(L.dcan);
    cudaFree(L.support);
    cudaFree(L.nsupport);
    cudaFree(L.sf_changed);
    cudaFree(L.ntri);
    cudaFree(L.grid_tmp);
    cudaFree(L.Draw);
    cudaFree(L.Dlr);
    cudaFree(L.Dtmp);
    cudaFree(L.labels);
    cudaFree(L.sizes);
    cudaFree(L.ccl_roots);
    cudaFree(L.ccl_counts);
    cudaFree(L.dmap);
    cudaFreeHost(L.h_support);
    cudaFreeHost(L.h_nsupport);
    cudaFreeHost(L.h_ntri);
    cudaFreeHost(L.h_order);
    cudaFreeHost(L.h_order_ok);
    cudaFree(L.d_order);
    cudaFree(L.d_order_ok);
    cudaFreeHost(L.h_dd_done);
    cudaFree(L.dup_count);
    cudaFree(L.dup_keys);
    if (L.ev_a) cudaEventDestroy(L.ev_a);
    if (L.ev_done) cudaEventDestroy(L.ev_done);
    if (L.own_stream) cudaStreamDestroy(L.own_stream);
    L = Lane();
}

struct StageTimer {
    Lane &L;
    StageEvents *se;
    StageTimer(Lane &L_, StageEvents *se_) : L(L_), se(se_) {}
    int mark(int idx) {
        if (!se) return SVB_OK;
        SVB_CUDA(cudaEventRecord(se->ev[idx], L.stream));
        return SVB_OK;
    }
};

// make sure `n` per-chunk event sets exist and mark them unused
int stage_events_prepare(svb_context *c, int n) {
    if (!c->stage_timing) return SVB_OK;
    while ((int)c->stage_ev.size() < n) {
        StageEvents se;
        for (int i = 0; i <= ST_COUNT; i++) SVB_CUDA(cudaEventCreate(&se.ev[i]));
        SVB_CUDA(cudaEventCreate(&se.a_end));
        c->stage_ev.push_back(se);
    }
    for (auto &se : c->stage_ev) se.a_done = se.b_done = false;
    return SVB_OK;
}

StageEvents *stage_events_of(svb_context *c, int chunk_index) {
    return (c->stage_timing && chunk_index < (int)c->stage_ev.size()) ? &c->stage_ev[chunk_index] : nullptr;
}

// accumulate per-stage times of every chunk of a call whose device work has completed
void stage_events_collect(svb_context *c) {
    if (!c->stage_timing) return;
    for (auto &se : c->stage_ev) {
        for (int i = 0; i < ST_COUNT; i++) {
            const bool in_a = i < ST_H2D_TRIANGLES;
            if (in_a ? !se.a_done : !se.b_done) continue;
            float ms = 0.f;
            // the D2H of the support lists ends stage A; the host stage separates it from the next device stage
            if (cudaEventElapsedTime(&ms, se.ev[i], i == ST_D2H_SUPPORT ? se.a_end : se.ev[i + 1]) == cudaSuccess)
                c->stats.stage_ms[i] += ms;
        }
    }
    cudaGetLastError();
}

int tap_store(svb_context *c, const char *name, const void *dev_src, size_t bytes, cudaStream_t s) {
    if (!c->tap_mode) return SVB_OK;
    Tap *t = nullptr;
    for (auto &x : c->taps)
        if (x.name == name) t = &x;
    if (!t) {
        c->taps.push_back(Tap());
        t = &c->taps.back();
        t->name = name;
    }
    if (t->bytes < bytes || !t->dev) {
        if (t->dev) cudaFree(t->dev);
        SVB_CUDA(cudaMalloc(&t->dev, bytes + 256));
    }
    t->bytes = bytes;
    SVB_CUDA(cudaMemcpyAsync(t->dev, dev_src, bytes, cudaMemcpyDeviceToDevice, s));
    return SVB_OK;
}

// Lists with duplicate coordinates (two thirds of the right-image lists of real frames): the reference keeps whichever duplicate its
// randomised quicksort leaves first, so someone has to replay that sort -- a sequential job.  On the device it costs about 1.4 ms of
// latency per chunk (k_replay_vertexsort, hidden only in part), on the host about 0.3 ms of CPU per frame for the whole list.  Which is
// better depends on how many host threads this context has: SVB_DELAUNAY_DUPS=device|host decides, the default is the host stage when
// it has at least 8 worker threads (one GPU on a 16-core box), the device otherwise (8 GPUs sharing 32 cores).
bool dups_on_device(const svb_context *c) {
    if (c->dups_policy == 1) return true;
    if (c->dups_policy == 2) return false;
    return c->pool && c->pool->size() < 8;
}

// ---- stage A: images (device) -> support lists (device + pinned host) ------------------------------
int stage_a(svb_context *c, Lane &L, const uint8_t *img1, const uint8_t *img2, int nf, StageEvents *se) {
    const Dims &d = c->d;
    StageTimer T(L, se);
    SVB_TRY(T.mark(ST_DESCRIPTOR));
    SVB_TRY(launch_descriptor(d, img1, L.desc[0], nf, L.stream));
    SVB_TRY(launch_descriptor(d, img2, L.desc[1], nf, L.stream));
    SVB_TRY(T.mark(ST_SUPPORT_MATCH));
    SVB_TRY(launch_support_match(d, c->p, L.desc[0], L.desc[1], L.dcan_raw, nf, L.stream));
    SVB_TRY(T.mark(ST_SUPPORT_FILTER));
    // the kernel writes the lists into the mapped pinned buffers itself: no device-to-host copy is queued
    SVB_TRY(launch_support_filter(d, c->p, L.dcan_raw, L.dcan, L.support, L.nsupport, L.h_support, L.h_nsupport, L.sf_changed, nf, L.stream));
    SVB_TRY(T.mark(ST_DELAUNAY_DEVICE));
    // ... and the order in which the host's divide-and-conquer will meet the vertices (sort + alternating cuts)
    L.unpacked = false;
    if (c->gpu_order) {
        SVB_TRY(launch_delaunay_order(d, L.support, L.nsupport, L.h_order, L.h_order_ok, L.d_order, L.d_order_ok, dups_on_device(c) ? L.dup_count : nullptr,
                                      dups_on_device(c) ? L.dup_keys : nullptr, nf, L.stream));
        // ... and then the divide-and-conquer itself, straight into the lane's triangle arena (k_delaunay.cu); the host stage only
        // takes the lists the device leaves to it (duplicate coordinates, more points than the launch has shared memory for)
        if (c->delaunay_device && !c->inject[0] && !c->inject[1]) {
            SVB_TRY(launch_delaunay_levels(d, L.support, L.nsupport, L.d_order, L.d_order_ok, L.tri[0], L.tri[1], L.h_ntri, L.h_dd_done, nf, c->dd_cap,
                                           L.stream));
            L.unpacked = true;
        }
    }
    SVB_TRY(T.mark(ST_D2H_SUPPORT));
    if (se) {
        SVB_CUDA(cudaEventRecord(se->a_end, L.stream));
        se->a_done = true;
    }
    SVB_CUDA(cudaEventRecord(L.ev_a, L.stream));
    return SVB_OK;
}

// ---- host stage ----------------------------------------------------------------------------------------
}  // namespace
int svb::stage_host(svb_context *c, Lane &L, int nf, bool unpacked) {
    const Dims &d = c->d;
    SVB_CUDA(cudaEventSynchronize(L.ev_a));
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<double> per_worker(c->pool->size(), 0.0);
    int32_t *h_trioff = L.h_ntri + 2 * c->chunk;
    std::fill(L.host_made.begin(), L.host_made.end(), (uint8_t)0);
    std::vector<int> jobs;  // 2 f + side of the lists made here
    jobs.reserve(2 * (size_t)nf);
    int max_n = 0;
    if (unpacked) {
        // The device wrote frame f's lists at triangle f * (maxT + 8) of the lane's arena and their sizes into h_ntri; what it left
        // (h_dd_done = 0) is triangulated here into the same slot of the pinned mirror and copied up list by list.
        for (int f = 0; f < nf; f++) {
            int n = L.h_nsupport[f];
            if (n < 0 || n > d.maxS) n = L.h_nsupport[f] = 0;
            max_n = std::max(max_n, n);
            h_trioff[f] = f * (d.maxT + 8);
            for (int side = 0; side < 2; side++) {
                if (c->inject[side] || L.h_dd_done[2 * f + side] != 1) {
                    jobs.push_back(2 * f + side);
                } else {
                    c->stats.delaunay_lists_device++;
                }
                L.h_dd_done[2 * f + side] = 0;
            }
        }
        h_trioff[nf] = nf * (d.maxT + 8);
        // shared-memory capacity of the next launches follows the lists seen (a list above it is simply the host's)
        if (max_n > c->dd_cap * 4 / 5 && c->dd_cap < 4096) c->dd_cap = std::min(4096, ((max_n * 5 / 4 + 511) / 512) * 512);
    } else {
        // The triangle lists of a chunk are packed back to back: a triangulation of n points has at most 2n - 5 triangles,
        // so frame f gets room for 2 n_f (or the injected list) starting at h_trioff[f]; only that much crosses PCIe.
        int off = 0;
        for (int f = 0; f < nf; f++) {
            int n = L.h_nsupport[f];
            if (n < 0 || n > d.maxS) n = L.h_nsupport[f] = 0;
            int cap = 2 * n + 8;
            for (int side = 0; side < 2; side++)
                if (c->inject[side]) cap = std::max(cap, (int)(c->inject_tri[side].size() / 3));
            h_trioff[f] = off;
            off += cap;
            jobs.push_back(2 * f);
            jobs.push_back(2 * f + 1);
        }
        h_trioff[nf] = off;  // total (h_ntri holds 3*chunk ints + slack, h_trioff[chunk] is the slack slot)
        if ((size_t)off > (size_t)c->chunk * (d.maxT + 8)) {
            set_error("triangle lists do not fit the arena (%d > %zu)", off, (size_t)c->chunk * (d.maxT + 8));
            return SVB_ERR_ARG;
        }
    }
    // fewer lists than threads (one 4K frame): a large list may spread its subtrees over the threads that would idle otherwise
    // (these are threads of the triangulation's own, not the pool's: a one-frame context has a pool of two)
    const int par_threads = jobs.empty() ? 1 : std::max(1, c->host_threads / (int)jobs.size());
    if (!jobs.empty())
        c->pool->parallel_for((int)jobs.size(), [&](int j, int worker) {
            const auto w0 = std::chrono::steady_clock::now();
            c->scratch[worker].par_threads = par_threads;
            const int f = jobs[j] >> 1, side = jobs[j] & 1;
            const int n = L.h_nsupport[f];
            const int cap = unpacked ? d.maxT + 8 : h_trioff[f + 1] - h_trioff[f];
            int32_t *out = L.h_tri[side] + (size_t)h_trioff[f] * 3;
            int m = 0;
            if (c->inject[side]) {
                m = (int)(c->inject_tri[side].size() / 3);
                if (m > cap) m = cap;
                memcpy(out, c->inject_tri[side].data(), sizeof(int32_t) * 3 * m);
            } else if (n >= 3) {
                const int32_t *sup = L.h_support + (size_t)f * d.maxS * 3;
                m = -1;
                const int nv = L.h_order_ok[2 * f + side];  // vertices in the device's order (duplicates already removed), 0 = not usable
                if (c->gpu_order && nv >= 3)
                    m = delaunay_support_ordered(sup, n, side, L.h_order + ((size_t)f * 2 + side) * d.maxS, nv, out, cap, c->scratch[worker]);
                if (m < 0) m = delaunay_support(sup, n, side, out, cap, c->scratch[worker]);
                if (m > cap) m = cap;
            }
            L.h_ntri[2 * f + side] = m;
            L.host_made[2 * f + side] = 1;
            per_worker[worker] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        });
    const auto t1 = std::chrono::steady_clock::now();
    c->stats.delaunay_ms_wall += std::chrono::duration<double, std::milli>(t1 - t0).count();
    for (double v : per_worker) c->stats.delaunay_ms_total += v;
    c->stats.delaunay_lists_host += (int64_t)jobs.size();
    for (int f = 0; f < nf; f++) {
        c->stats.support_points += L.h_nsupport[f];
        c->stats.triangles += L.h_ntri[2 * f] + L.h_ntri[2 * f + 1];
        if (L.h_nsupport[f] < 3) c->stats.frames_failed++;
        if (L.first_frame >= 0 && (size_t)(L.first_frame + f) < c->frame_nsupport.size()) c->frame_nsupport[L.first_frame + f] = L.h_nsupport[f];
    }
    return SVB_OK;
}

namespace {
// ---- stage B: triangles (pinned host) -> disparity / points -------------------------------------------
// out_D1 / out_points may be null.  The final maps stay in L.Dlr ([0] = left, [1] = right).
int stage_b(svb_context *c, Lane &L, int nf, float *out_D1, double *out_points, StageEvents *se) {
    const Dims &d = c->d;
    const svb_params &p = c->p;
    const size_t N = (size_t)d.N, C = (size_t)c->chunk, DN = (size_t)d.DN;
    StageTimer T(L, se);
    int max_tri = 0, max_support = 0;
    for (int i = 0; i < 2 * nf; i++) max_tri = L.h_ntri[i] > max_tri ? L.h_ntri[i] : max_tri;
    for (int i = 0; i < nf; i++) max_support = L.h_nsupport[i] > max_support ? L.h_nsupport[i] : max_support;
    SVB_TRY(T.mark(ST_H2D_TRIANGLES));
    SVB_CUDA(cudaMemcpyAsync(L.ntri, L.h_ntri, sizeof(int32_t) * 3 * C, cudaMemcpyHostToDevice, L.stream));
    if (L.unpacked) {
        // the device made the lists in place; only what the host stage produced goes up, list by list (normally nothing)
        for (int i = 0; i < 2 * nf; i++)
            if (L.host_made[i] && L.h_ntri[i] > 0) {
                const size_t at = (size_t)L.h_ntri[2 * C + (i >> 1)] * 3;
                SVB_CUDA(cudaMemcpyAsync(L.tri[i & 1] + at, L.h_tri[i & 1] + at, sizeof(int32_t) * 3 * (size_t)L.h_ntri[i], cudaMemcpyHostToDevice, L.stream));
            }
    } else {
        const size_t tri_total = (size_t)L.h_ntri[2 * C + nf];  // packed size of this chunk's lists, in triangles
        for (int s = 0; s < 2; s++)
            if (tri_total) SVB_CUDA(cudaMemcpyAsync(L.tri[s], L.h_tri[s], sizeof(int32_t) * 3 * tri_total, cudaMemcpyHostToDevice, L.stream));
    }
    SVB_TRY(T.mark(ST_PLANES));
    float *pr1 = (c->tap_mode && nf == 1) ? c->planes_ref[0] : nullptr;
    float *pr2 = (c->tap_mode && nf == 1) ? c->planes_ref[1] : nullptr;
    SVB_TRY(launch_planes(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, pr1, pr2, L.rec[0], L.rec[1], nf, max_tri, L.stream));
    SVB_TRY(T.mark(ST_GRID));
    SVB_TRY(launch_grid(d, p, L.support, L.nsupport, L.grid_tmp, L.grid[0], L.grid[1], nf, max_support, L.stream));
    SVB_TRY(T.mark(ST_RASTER));
    // Outside tap mode the owner maps are not cleared per chunk: their entries carry a generation number (svb_internal.h) that moves
    // on with every chunk of the lane; the maps are cleared when the 7-bit number wraps (and once at creation).
    int owner_gen = 0;
    if (!c->tap_mode) {
        if (++L.owner_gen > OWNER_GEN_MAX) {
            for (int sd = 0; sd < 2; sd++) SVB_CUDA(cudaMemsetAsync(L.owner[sd], 0xFF, (size_t)C * N * sizeof(int32_t), L.stream));
            L.owner_gen = 1;
        }
        owner_gen = L.owner_gen;
    }
    SVB_TRY(launch_raster(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, L.owner[0], L.owner[1], nf, max_tri, L.stream, owner_gen));
    SVB_TRY(T.mark(ST_DENSE));
    float *D1raw = L.Draw, *D2raw = L.Draw + C * N;
    float *D1 = L.Dlr, *D2 = L.Dlr + C * N;
    SVB_TRY(launch_dense(d, p, L.desc[0], L.desc[1], L.owner[0], L.owner[1], L.rec[0], L.rec[1], L.grid[0], L.grid[1], D1raw, D2raw, nf,
                         L.stream, owner_gen));
    SVB_TRY(T.mark(ST_LR));
    const bool both = !p.postprocess_only_left;
    const bool need_d2 = both || c->tap_mode || (out_D1 == nullptr && out_points == nullptr);
    SVB_TRY(launch_lr_check(d, p, D1raw, D2raw, D1, need_d2 ? D2 : nullptr, nf, L.stream));
    if (c->tap_mode && nf == 1) {
        SVB_TRY(tap_store(c, "D1raw", D1raw, DN * 4, L.stream));
        SVB_TRY(tap_store(c, "D2raw", D2raw, DN * 4, L.stream));
        SVB_TRY(tap_store(c, "D1lr", D1, DN * 4, L.stream));
        SVB_TRY(tap_store(c, "D2lr", D2, DN * 4, L.stream));
    }
    // post-processing chain; when both maps are processed they are handled as 2*nf independent images, which
    // needs the left and right blocks to be adjacent: true when nf == chunk, otherwise run the sides separately
    const int passes = both ? 2 : 1;
    SVB_TRY(T.mark(ST_SEGMENTS));
    // one map to post-process and no taps: the pruning step of the speckle removal rides on the row pass of the gap interpolation
    const bool prune_in_gap = passes == 1 && !c->tap_mode;
    for (int s = 0; s < passes; s++) {
        if (prune_in_gap)
            SVB_TRY(launch_ccl_label(d, p, D1, L.labels, L.sizes, L.ccl_roots, L.ccl_counts, nf, L.stream));
        else
            SVB_TRY(launch_remove_small_segments(d, p, s ? D2 : D1, L.labels, L.sizes, L.ccl_roots, L.ccl_counts, nf, L.stream));
    }
    if (c->tap_mode && nf == 1) {
        SVB_TRY(tap_store(c, "D1seg", D1, DN * 4, L.stream));
        if (both) SVB_TRY(tap_store(c, "D2seg", D2, DN * 4, L.stream));
    }
    SVB_TRY(T.mark(ST_GAP));
    // (the validity bits go through Dtmp, which the tail kernels only use later)
    for (int s = 0; s < passes; s++)
        SVB_TRY(launch_gap(d, p, s ? D2 : D1, reinterpret_cast<uint32_t *>(L.Dtmp), prune_in_gap ? L.labels : nullptr, prune_in_gap ? L.sizes : nullptr, nf,
                           L.stream));
    if (c->tap_mode && nf == 1) {
        SVB_TRY(tap_store(c, "D1gap", D1, DN * 4, L.stream));
        if (both) SVB_TRY(tap_store(c, "D2gap", D2, DN * 4, L.stream));
    }
    // Tail of the chain.  Full-resolution maps outside tap mode: ONE kernel does adaptive mean, median, the final map, the u8 map
    // and the point cloud (k_post_fused.cu); tap mode and the half-resolution chain keep the stage-by-stage kernels.
    const bool fused = c->fused_post && !c->tap_mode && !d.sub;
    const bool float_disp = c->points_float_disp;
    SVB_TRY(T.mark(ST_MEAN));
    if (!fused && p.filter_adaptive_mean) {
        for (int s = 0; s < passes; s++) SVB_TRY(launch_adaptive_mean(d, c->mean_mode, s ? D2 : D1, L.Dtmp, nf, L.stream));
        if (c->tap_mode && nf == 1) {
            SVB_TRY(tap_store(c, "D1mean", D1, DN * 4, L.stream));
            if (both) SVB_TRY(tap_store(c, "D2mean", D2, DN * 4, L.stream));
        }
    }
    SVB_TRY(T.mark(ST_MEDIAN));
    if (!fused && p.filter_median) {
        for (int s = 0; s < passes; s++) SVB_TRY(launch_median(d, s ? D2 : D1, L.Dtmp, nf, L.stream));
        if (c->tap_mode && nf == 1) {
            SVB_TRY(tap_store(c, "D1med", D1, DN * 4, L.stream));
            if (both) SVB_TRY(tap_store(c, "D2med", D2, DN * 4, L.stream));
        }
    }
    SVB_TRY(T.mark(ST_REPROJECT));
    if (!fused) {
        if (out_D1) SVB_CUDA(cudaMemcpyAsync(out_D1, D1, DN * 4 * nf, cudaMemcpyDeviceToDevice, L.stream));
        if (out_points) {
            if (float_disp)
                SVB_TRY(launch_reproject_float(d, c->calib, D1, out_points, nf, L.stream));
            else
                SVB_TRY(launch_reproject(d, c->calib, D1, L.dmap, out_points, nf, L.stream));
        }
    }
    SVB_TRY(T.mark(ST_POST_FUSED));
    if (fused) {
        for (int s = 0; s < passes; s++) {
            float *src = s ? D2 : D1;
            // the final left map goes straight into the batch store when there is one; otherwise (Elas::process) it comes back into the
            // lane's map through the scratch arena, because a tile reads its neighbours' halos and cannot work in place
            float *dst = (s == 0 && out_D1) ? out_D1 : L.Dtmp;
            SVB_TRY(launch_post_fused(d, p, c->mean_mode, c->calib, src, dst, (s == 0 && out_points) ? L.dmap : nullptr,
                                      s == 0 ? out_points : nullptr, float_disp ? 1 : 0, nf, L.stream));
            if (dst == L.Dtmp) SVB_CUDA(cudaMemcpyAsync(src, L.Dtmp, DN * 4 * nf, cudaMemcpyDeviceToDevice, L.stream));
        }
    }
    // Batch outputs of a frame with fewer than 3 support points: Elas::process returns without touching D (elas.cpp:64-69) and the
    // driver's maps start as zeros (stereo_vision.cu:311-312), so the frame's disparity is 0 everywhere, like svb_point_cloud_bgra
    if (out_D1 || out_points)
        for (int f = 0; f < nf; f++)
            if (L.h_nsupport[f] < 3) {
                SVB_CUDA(cudaMemsetAsync(D1 + (size_t)f * DN, 0, DN * 4, L.stream));
                if (out_D1) SVB_CUDA(cudaMemsetAsync(out_D1 + (size_t)f * DN, 0, DN * 4, L.stream));
                if (out_points) SVB_TRY(launch_reproject(d, c->calib, D1 + (size_t)f * DN, L.dmap + (size_t)f * N, out_points + (size_t)f * N * 3, 1, L.stream));
            }
    SVB_TRY(T.mark(ST_COUNT));
    if (se) se->b_done = true;
    SVB_CUDA(cudaEventRecord(L.ev_done, L.stream));
    return SVB_OK;
}

int tap_after_a(svb_context *c, Lane &L) {
    const Dims &d = c->d;
    const size_t N = (size_t)d.N;
    SVB_TRY(tap_store(c, "desc1", L.desc[0], N * 16, L.stream));
    SVB_TRY(tap_store(c, "desc2", L.desc[1], N * 16, L.stream));
    SVB_TRY(tap_store(c, "dcan_raw", L.dcan_raw, (size_t)d.cw * d.ch * 2, L.stream));
    SVB_TRY(tap_store(c, "dcan", L.dcan, (size_t)d.cw * d.ch * 2, L.stream));
    return SVB_OK;
}

int ensure_calib(svb_context *c) {
    if (c->have_calib) return SVB_OK;
    // identity-like default: Q = I, XR = I, XT = 0 (callers set the real one with svb_set_calibration)
    memset(&c->calib, 0, sizeof(c->calib));
    for (int i = 0; i < 4; i++) c->calib.Q[5 * i] = 1.0;
    for (int i = 0; i < 3; i++) c->calib.XR[4 * i] = 1.0;
    return SVB_OK;
}

void stats_reset(svb_context *c) {
    memset(&c->stats, 0, sizeof(c->stats));
    g_launch_counter = 0;
}

}  // namespace

// =====================================================================================================
//                                          C-ABI
// =====================================================================================================
extern "C" {

const char *svb_last_error(void) { return g_err; }
const char *svb_version(void) { return "elas_b200 0.1 (sm_100a)"; }

int svb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int svb_default_params(int setting, svb_params *o) {
    if (!o) return SVB_ERR_ARG;
    // src/parallel_includes/elas/elas.h:90-113 (ROBOTICS) and :117-140 (MIDDLEBURY)
    const bool rob = setting == SVB_ROBOTICS;
    o->disp_min = 0;
    o->disp_max = 255;
    o->support_threshold = rob ? 0.85f : 0.95f;
    o->support_texture = 10;
    o->candidate_stepsize = 5;
    o->incon_window_size = 5;
    o->incon_threshold = 5;
    o->incon_min_support = 5;
    o->add_corners = rob ? 0 : 1;
    o->grid_size = 20;
    o->beta = 0.02f;
    o->gamma = rob ? 3.f : 5.f;
    o->sigma = 1.f;
    o->sradius = rob ? 2.f : 3.f;
    o->match_texture = rob ? 1 : 0;
    o->lr_threshold = 2;
    o->speckle_sim_threshold = 1.f;
    o->speckle_size = 200;
    o->ipol_gap_width = rob ? 3 : 5000;
    o->filter_median = rob ? 0 : 1;
    o->filter_adaptive_mean = rob ? 1 : 0;
    o->postprocess_only_left = rob ? 1 : 0;
    o->subsampling = 0;
    if (setting == SVB_PIPELINE) {  // stereo_vision.cu:315-319
        o->postprocess_only_left = 1;
        o->filter_adaptive_mean = 1;
    }
    return SVB_OK;
}

svb_context *svb_create(const svb_params *params, int width, int height, int chunk, int device) {
    if (!params) {
        set_error("params is NULL");
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
        return nullptr;
    }
    int caller_device = -1;
    if (cudaGetDevice(&caller_device) != cudaSuccess) caller_device = -1;
    if (device < 0) device = caller_device >= 0 ? caller_device : 0;
    // the caller's current device is restored on every way out (RAII)
    struct RestoreDevice {
        int dev;
        ~RestoreDevice() {
            if (dev >= 0) cudaSetDevice(dev);
        }
    } restore{caller_device};
    if (device >= ndev) {
        set_error("device %d out of range (%d devices)", device, ndev);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    svb_context *c = new svb_context();
    c->p = *params;
    c->device = device;
    c->chunk = chunk < 1 ? 1 : chunk;
    if (make_dims(c->p, width, height, &c->d) != SVB_OK) {
        delete c;
        return nullptr;
    }
    {
        const char *g = getenv("SVB_GPU_ORDER");
        if (g && atoi(g) == 0) c->gpu_order = false;
        const char *dd = getenv("SVB_DELAUNAY_DEVICE");
        if (dd && atoi(dd) == 0) c->delaunay_device = false;
        c->dd_cap = std::min(c->chunk == 1 ? 4096 : 2048, std::max(c->d.maxS, 3));  // single-frame contexts: one CTA per side, take all it can
        const char *dp = getenv("SVB_DELAUNAY_DUPS");
        if (dp && !strcmp(dp, "device")) c->dups_policy = 1;
        if (dp && !strcmp(dp, "host")) c->dups_policy = 2;
        const char *fp = getenv("SVB_FUSED_POST");
        if (fp && atoi(fp) == 0) c->fused_post = false;
        const char *e = getenv("SVB_LANES");
        const int n = e ? atoi(e) : 8;  // measured: 8 lanes hide the latency-bound stages (lattice filters, Delaunay, vertex-sort replay) better than 4: + 2 % synthetic, + 7 % kitti_mini
        c->n_lanes = n < 1 ? 1 : (n > MAX_LANES ? MAX_LANES : n);
        if (c->chunk == 1) c->n_lanes = 1;  // single-frame contexts never pipeline
    }
    for (int i = 0; i < c->n_lanes; i++)
        if (lane_create(c, c->lanes[i]) != SVB_OK) {
            svb_destroy(c);
            return nullptr;
        }
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = hw ? (int)hw : 4;
    if (nthreads > 2 * c->chunk) nthreads = 2 * c->chunk;
    if (nthreads > 64) nthreads = 64;
    c->host_threads = hw ? (int)hw : 4;
    c->pool.reset(new ThreadPool(nthreads));
    c->scratch.resize(c->pool->size());
    memset(&c->stats, 0, sizeof(c->stats));
    ensure_calib(c);
    return c;
}

void svb_destroy(svb_context *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < MAX_LANES; i++)
        if (c->lanes[i].own_stream) lane_destroy(c->lanes[i]);
    for (auto &t : c->taps) cudaFree(t.dev);
    for (auto &se : c->stage_ev) {
        for (int i = 0; i <= ST_COUNT; i++)
            if (se.ev[i]) cudaEventDestroy(se.ev[i]);
        if (se.a_end) cudaEventDestroy(se.a_end);
    }
    for (int s = 0; s < 2; s++) {
        cudaFree(c->planes_ref[s]);
        cudaFree(c->in_img[s]);
        cudaFree(c->bgra[s]);
        cudaFree(c->bgra_batch[s]);
    }
    for (int i = 0; i < 4; i++)
        if (c->ev_pc[i]) cudaEventDestroy(c->ev_pc[i]);
    cudaFree(c->out_D1);
    cudaFree(c->out_points);
    cudaFree(c->d.evals);
    delete c;
}

int svb_set_mean_mode(svb_context *c, int mode) {
    if (!c || (mode != 0 && mode != 1)) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    c->mean_mode = mode;
    return SVB_OK;
}

int svb_set_delaunay_threads(svb_context *c, int n) {
    if (!c || n < 1) return SVB_ERR_ARG;
    c->pool->resize(n);
    c->scratch.resize(c->pool->size());
    return SVB_OK;
}

int svb_set_stage_timing(svb_context *c, int on) {
    if (!c) return SVB_ERR_ARG;
    c->stage_timing = on != 0;
    return SVB_OK;
}

int svb_set_single_stream(svb_context *c, int on) {
    if (!c) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    SVB_CUDA(cudaDeviceSynchronize());
    c->single_stream = on != 0;
    for (int i = 0; i < c->n_lanes; i++) c->lanes[i].stream = c->single_stream ? c->lanes[0].own_stream : c->lanes[i].own_stream;
    return SVB_OK;
}

int svb_set_eval_counting(svb_context *c, int on) {
    if (!c) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    SVB_CUDA(cudaDeviceSynchronize());
    if (on && !c->d.evals) {
        SVB_CUDA(cudaMalloc((void **)&c->d.evals, 2 * sizeof(unsigned long long)));
        SVB_CUDA(cudaMemset(c->d.evals, 0, 2 * sizeof(unsigned long long)));
    } else if (!on && c->d.evals) {
        cudaFree(c->d.evals);
        c->d.evals = nullptr;
    }
    return SVB_OK;
}

int svb_get_eval_counts(svb_context *c, uint64_t *support_hypotheses, uint64_t *dense_hypotheses) {
    if (!c) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->d.evals) {
        set_error("svb_get_eval_counts: counting is off (svb_set_eval_counting)");
        return SVB_ERR_ARG;
    }
    SVB_CUDA(cudaSetDevice(c->device));
    SVB_CUDA(cudaDeviceSynchronize());
    unsigned long long h[2] = {0, 0};
    SVB_CUDA(cudaMemcpy(h, c->d.evals, sizeof(h), cudaMemcpyDeviceToHost));
    SVB_CUDA(cudaMemset(c->d.evals, 0, sizeof(h)));
    if (support_hypotheses) *support_hypotheses = h[0];
    if (dense_hypotheses) *dense_hypotheses = h[1];
    return SVB_OK;
}

int svb_set_tap_mode(svb_context *c, int on) {
    if (!c) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    c->tap_mode = on != 0;
    if (c->tap_mode && !c->planes_ref[0])
        for (int s = 0; s < 2; s++) SVB_TRY(dev_alloc(&c->planes_ref[s], (size_t)c->d.maxT * 6));
    return SVB_OK;
}

int svb_inject_triangles(svb_context *c, int side, const int32_t *tri, int n) {
    if (!c || side < 0 || side > 1) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (n < 0 || !tri) {
        c->inject[side] = false;
        c->inject_tri[side].clear();
        return SVB_OK;
    }
    c->inject_tri[side].assign(tri, tri + (size_t)3 * n);
    c->inject[side] = true;
    return SVB_OK;
}

int svb_set_calibration(svb_context *c, const double *Q16, const double *XR9, const double *XT3) {
    if (!c || !Q16) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    memcpy(c->calib.Q, Q16, sizeof(double) * 16);
    if (XR9)
        memcpy(c->calib.XR, XR9, sizeof(double) * 9);
    else {
        memset(c->calib.XR, 0, sizeof(c->calib.XR));
        for (int i = 0; i < 3; i++) c->calib.XR[4 * i] = 1.0;
    }
    if (XT3)
        memcpy(c->calib.XT, XT3, sizeof(double) * 3);
    else
        memset(c->calib.XT, 0, sizeof(c->calib.XT));
    c->have_calib = true;
    return SVB_OK;
}

// ---- Elas::process ----------------------------------------------------------------------------------------
int svb_process(svb_context *c, const uint8_t *I1, const uint8_t *I2, int stride, float *D1, float *D2) {
    if (!c || !I1 || !I2 || !D1 || !D2 || stride < c->d.W) {
        set_error("svb_process: bad argument");
        return SVB_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    stats_reset(c);
    const Dims &d = c->d;
    Lane &L = c->lanes[0];
    const size_t N = (size_t)d.N, C = (size_t)c->chunk;
    // elas.cpp:33-50: rows are copied out of the caller's stride
    SVB_CUDA(cudaMemcpy2DAsync(L.img[0], d.bpl, I1, stride, d.W, d.H, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpy2DAsync(L.img[1], d.bpl, I2, stride, d.W, d.H, cudaMemcpyHostToDevice, L.stream));
    SVB_TRY(stage_events_prepare(c, 1));
    StageEvents *se = stage_events_of(c, 0);
    SVB_TRY(stage_a(c, L, L.img[0], L.img[1], 1, se));
    if (c->tap_mode) SVB_TRY(tap_after_a(c, L));
    SVB_TRY(stage_host(c, L, 1, L.unpacked));
    c->stats.frames = 1;
    const int n = L.h_nsupport[0];
    if (c->tap_mode) {
        SVB_TRY(tap_store(c, "support", L.support, (size_t)n * 12, L.stream));
    }
    if (n < 3) {
        // elas.cpp:64-69: "ERROR: Need at least 3 support points!", D1/D2 untouched
        SVB_CUDA(cudaStreamSynchronize(L.stream));
        set_error("need at least 3 support points (got %d)", n);
        c->stats.kernel_launches = g_launch_counter;
        return SVB_ERR_FEW_SUPPORT;
    }
    SVB_TRY(stage_b(c, L, 1, nullptr, nullptr, se));
    if (c->tap_mode) {
        SVB_TRY(tap_store(c, "tri1", L.tri[0], (size_t)L.h_ntri[0] * 12, L.stream));
        SVB_TRY(tap_store(c, "tri2", L.tri[1], (size_t)L.h_ntri[1] * 12, L.stream));
        SVB_TRY(tap_store(c, "planes1", c->planes_ref[0], (size_t)L.h_ntri[0] * 24, L.stream));
        SVB_TRY(tap_store(c, "planes2", c->planes_ref[1], (size_t)L.h_ntri[1] * 24, L.stream));
        SVB_TRY(tap_store(c, "owner1", L.owner[0], (size_t)d.DN * 4, L.stream));
        SVB_TRY(tap_store(c, "owner2", L.owner[1], (size_t)d.DN * 4, L.stream));
        // grids in the reference's list layout
        const size_t gbytes = (size_t)d.gw * d.gh * (c->p.disp_max + 2) * 4;
        for (int s = 0; s < 2; s++) {
            int32_t *tmp = nullptr;
            SVB_TRY(dev_alloc(&tmp, gbytes / 4));
            int r = launch_grid_expand(d, c->p, L.grid[s], tmp, L.stream);
            if (r == SVB_OK) r = tap_store(c, s ? "grid2" : "grid1", tmp, gbytes, L.stream);
            cudaStreamSynchronize(L.stream);
            cudaFree(tmp);
            SVB_TRY(r);
        }
    }
    // D1 / D2 are (W/2) x (H/2) with subsampling (elas.h:157-160)
    SVB_CUDA(cudaMemcpyAsync(D1, L.Dlr, (size_t)d.DN * 4, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaMemcpyAsync(D2, L.Dlr + C * N, (size_t)d.DN * 4, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    stage_events_collect(c);
    c->stats.kernel_launches = g_launch_counter;
    return SVB_OK;
}

int64_t svb_tap(svb_context *c, const char *name, void *dst, int64_t cap) {
    if (!c || !name || !dst) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    for (auto &t : c->taps)
        if (t.name == name) {
            if ((int64_t)t.bytes > cap) {
                set_error("tap %s needs %zu bytes, capacity %lld", name, t.bytes, (long long)cap);
                return SVB_ERR_ARG;
            }
            SVB_CUDA(cudaSetDevice(c->device));
            SVB_CUDA(cudaMemcpy(dst, t.dev, t.bytes, cudaMemcpyDeviceToHost));
            return (int64_t)t.bytes;
        }
    set_error("no such tap: %s", name);
    return SVB_ERR_ARG;
}

// ---- stage-isolated entry points -------------------------------------------------------------------------
#define STAGE_PROLOG()                              \
    if (!c) return SVB_ERR_ARG;                     \
    std::lock_guard<std::mutex> lk(c->mu);          \
    SVB_CUDA(cudaSetDevice(c->device));             \
    const Dims &d = c->d;                           \
    Lane &L = c->lanes[0];                          \
    const size_t N = (size_t)d.N;                   \
    (void)N;                                        \
    (void)L;

int svb_stage_descriptor(svb_context *c, const uint8_t *I, int stride, uint8_t *desc_out) {
    STAGE_PROLOG();
    if (!I || !desc_out || stride < d.W) return SVB_ERR_ARG;
    SVB_CUDA(cudaMemcpy2DAsync(L.img[0], d.bpl, I, stride, d.W, d.H, cudaMemcpyHostToDevice, L.stream));
    SVB_TRY(launch_descriptor(d, L.img[0], L.desc[0], 1, L.stream));
    SVB_CUDA(cudaMemcpyAsync(desc_out, L.desc[0], N * 16, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    return SVB_OK;
}

int svb_stage_support(svb_context *c, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int16_t *dcan, int32_t *support, int cap,
                      int *n_out) {
    STAGE_PROLOG();
    if (!desc1 || !desc2) return SVB_ERR_ARG;
    SVB_CUDA(cudaMemcpyAsync(L.desc[0], desc1, N * 16, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpyAsync(L.desc[1], desc2, N * 16, cudaMemcpyHostToDevice, L.stream));
    SVB_TRY(launch_support_match(d, c->p, L.desc[0], L.desc[1], L.dcan_raw, 1, L.stream));
    SVB_TRY(launch_support_filter(d, c->p, L.dcan_raw, L.dcan, L.support, L.nsupport, L.h_support, L.h_nsupport, L.sf_changed, 1, L.stream));
    const size_t cb = (size_t)d.cw * d.ch * 2;
    if (dcan_raw) SVB_CUDA(cudaMemcpyAsync(dcan_raw, L.dcan_raw, cb, cudaMemcpyDeviceToHost, L.stream));
    if (dcan) SVB_CUDA(cudaMemcpyAsync(dcan, L.dcan, cb, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    const int n = L.h_nsupport[0];
    if (n_out) *n_out = n;
    if (support) memcpy(support, L.h_support, (size_t)(n < cap ? n : cap) * 12);
    return SVB_OK;
}

// The integer predicates of the Delaunay stage (delaunay_mesh.h) are exact for x in [-8192, 16383] and y in [0, 8191] -- every list a
// frame of up to 8192 x 8192 pixels with disparities up to 4095 can produce, corner points (u + d) included.  Lists handed to the stage
// entry points directly are held to the same range instead of overflowing silently.
static bool support_coordinates_in_range(const int32_t *support, int n, int right_image) {
    for (int i = 0; i < n; i++) {
        const long long x = right_image ? (long long)support[3 * i] - support[3 * i + 2] : support[3 * i], y = support[3 * i + 1];
        if (x < -8192 || x > 16383 || y < 0 || y > 8191) {
            set_error("support point %d: (%lld, %lld) outside the Delaunay stage's coordinate range (x -8192 .. 16383, y 0 .. 8191)", i, x, y);
            return false;
        }
    }
    return true;
}

int svb_stage_delaunay(const int32_t *support, int n, int right_image, int32_t *tri, int cap, int *n_tri_out) {
    if (!support || !tri || n < 0 || cap < 0) return SVB_ERR_ARG;
    if (!support_coordinates_in_range(support, n, right_image)) return SVB_ERR_ARG;
    DelaunayScratch scratch;
    if (const char *pt = getenv("SVB_DELAUNAY_PAR")) scratch.par_threads = atoi(pt);  // tests: subtrees of a large list on threads of their own
    const int m = delaunay_support(support, n, right_image, tri, cap, scratch);
    if (n_tri_out) *n_tri_out = m;
    return SVB_OK;
}

// The host half of the pipeline's Delaunay stage on its own (no GPU needed): `order` is what k_order.cu would deliver.
int svb_stage_delaunay_ordered(const int32_t *support, int n, int right_image, const int32_t *order, int32_t *tri, int cap, int *n_tri_out) {
    if (!support || !order || !tri || n < 0 || cap < 0) return SVB_ERR_ARG;
    if (!support_coordinates_in_range(support, n, right_image)) return SVB_ERR_ARG;
    DelaunayScratch scratch;
    const int m = delaunay_support_ordered(support, n, right_image ? 1 : 0, order, n, tri, cap, scratch);
    if (m < 0) {
        set_error("svb_stage_delaunay_ordered: `order` is not a permutation of 0..n-1");
        return SVB_ERR_ARG;
    }
    if (n_tri_out) *n_tri_out = m;
    return SVB_OK;
}

// The device's share of the Delaunay stage restated on the host (no GPU needed): levels below `host_levels` built level by level in
// 16-bit records like k_delaunay.cu, the rest by the host recursion.  `order` as for svb_stage_delaunay_ordered.
int svb_stage_delaunay_levels(const int32_t *support, int n, int right_image, const int32_t *order, int host_levels, int32_t *tri, int cap,
                              int *n_tri_out) {
    if (!support || !order || !tri || n < 0 || cap < 0) return SVB_ERR_ARG;
    if (!support_coordinates_in_range(support, n, right_image)) return SVB_ERR_ARG;
    DelaunayScratch scratch;
    const int m = delaunay_support_levels(support, n, right_image ? 1 : 0, order, host_levels, tri, cap, scratch);
    if (m < 0) {
        set_error("svb_stage_delaunay_levels: n > 4096 or `order` is not a permutation of 0..n-1");
        return SVB_ERR_ARG;
    }
    if (n_tri_out) *n_tri_out = m;
    return SVB_OK;
}

// Host Delaunay stage exactly as the pipeline runs it: the device orders the vertices (k_order.cu), the host recurses;
// *used_device_order: 2 = the device made the whole list (k_delaunay.cu), 1 = device vertex order + host recursion, 0 = complete host
// path (duplicates, > 4096 points, ...).
int svb_stage_delaunay_pipeline(svb_context *c, const int32_t *support, int n, int right_image, int32_t *tri, int cap, int *n_tri_out,
                                int *used_device_order) {
    STAGE_PROLOG();
    if (!support || !tri || n < 0 || cap < 0 || n > d.maxS) return SVB_ERR_ARG;
    if (!support_coordinates_in_range(support, n, right_image)) return SVB_ERR_ARG;
    L.h_nsupport[0] = n;
    memcpy(L.h_support, support, (size_t)n * 12);
    SVB_CUDA(cudaMemcpyAsync(L.nsupport, L.h_nsupport, 4, cudaMemcpyHostToDevice, L.stream));
    if (n) SVB_CUDA(cudaMemcpyAsync(L.support, L.h_support, (size_t)n * 12, cudaMemcpyHostToDevice, L.stream));
    L.h_order_ok[0] = L.h_order_ok[1] = 0;
    L.h_dd_done[0] = L.h_dd_done[1] = 0;
    SVB_TRY(launch_delaunay_order(d, L.support, L.nsupport, L.h_order, L.h_order_ok, L.d_order, L.d_order_ok, dups_on_device(c) ? L.dup_count : nullptr,
                                  dups_on_device(c) ? L.dup_keys : nullptr, 1, L.stream));
    if (c->delaunay_device)
        SVB_TRY(launch_delaunay_levels(d, L.support, L.nsupport, L.d_order, L.d_order_ok, L.tri[0], L.tri[1], L.h_ntri, L.h_dd_done, 1, c->dd_cap, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    const int side = right_image ? 1 : 0;
    int m = -1, used = 0;
    if (L.h_dd_done[side] == 1) {
        // the device made the whole list (k_delaunay.cu)
        m = L.h_ntri[side];
        const int take = m < cap ? m : cap;
        if (take > 0) SVB_CUDA(cudaMemcpy(tri, L.tri[side], sizeof(int32_t) * 3 * (size_t)take, cudaMemcpyDeviceToHost));
        used = 2;
    } else if (L.h_order_ok[side] >= 3) {
        m = delaunay_support_ordered(L.h_support, n, side, L.h_order + (size_t)side * d.maxS, L.h_order_ok[side], tri, cap, c->scratch[0]);
        used = m >= 0;
    }
    if (m < 0) m = delaunay_support(L.h_support, n, side, tri, cap, c->scratch[0]);
    L.h_order_ok[0] = L.h_order_ok[1] = 0;
    L.h_dd_done[0] = L.h_dd_done[1] = 0;
    if (n_tri_out) *n_tri_out = m;
    if (used_device_order) *used_device_order = used;
    return SVB_OK;
}

static int upload_support_and_tris(svb_context *c, Lane &L, const int32_t *support, int n, const int32_t *tri1, int m1, const int32_t *tri2,
                                   int m2) {
    const Dims &d = c->d;
    if (n < 0 || n > d.maxS || m1 > d.maxT || m2 > d.maxT) {
        set_error("support/triangle list too large (n=%d maxS=%d, m=%d/%d maxT=%d)", n, d.maxS, m1, m2, d.maxT);
        return SVB_ERR_ARG;
    }
    L.h_nsupport[0] = n;
    L.h_ntri[0] = m1 > 0 ? m1 : 0;
    L.h_ntri[1] = m2 > 0 ? m2 : 0;
    L.h_ntri[2 * c->chunk] = 0;  // one frame: its lists start at triangle 0
    memcpy(L.h_support, support, (size_t)n * 12);
    if (tri1 && m1 > 0) memcpy(L.h_tri[0], tri1, (size_t)m1 * 12);
    if (tri2 && m2 > 0) memcpy(L.h_tri[1], tri2, (size_t)m2 * 12);
    SVB_CUDA(cudaMemcpyAsync(L.nsupport, L.h_nsupport, 4, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpyAsync(L.ntri, L.h_ntri, sizeof(int32_t) * 3 * c->chunk, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpyAsync(L.support, L.h_support, (size_t)n * 12, cudaMemcpyHostToDevice, L.stream));
    if (tri1 && m1 > 0) SVB_CUDA(cudaMemcpyAsync(L.tri[0], L.h_tri[0], (size_t)m1 * 12, cudaMemcpyHostToDevice, L.stream));
    if (tri2 && m2 > 0) SVB_CUDA(cudaMemcpyAsync(L.tri[1], L.h_tri[1], (size_t)m2 * 12, cudaMemcpyHostToDevice, L.stream));
    return SVB_OK;
}

int svb_stage_planes(svb_context *c, const int32_t *support, int n, const int32_t *tri, int m, float *planes) {
    STAGE_PROLOG();
    if (!support || !tri || !planes) return SVB_ERR_ARG;
    float *tmp = nullptr;
    SVB_TRY(dev_alloc(&tmp, (size_t)d.maxT * 6));
    int r = upload_support_and_tris(c, L, support, n, tri, m, nullptr, 0);
    if (r == SVB_OK) r = launch_planes(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, tmp, nullptr, L.rec[0], L.rec[1], 1, m, L.stream);
    if (r == SVB_OK && cudaMemcpyAsync(planes, tmp, (size_t)m * 24, cudaMemcpyDeviceToHost, L.stream) != cudaSuccess) r = SVB_ERR_CUDA;
    cudaStreamSynchronize(L.stream);
    cudaFree(tmp);
    return r;
}

int svb_stage_grid(svb_context *c, const int32_t *support, int n, int right_image, int32_t *grid) {
    STAGE_PROLOG();
    if (!support || !grid) return SVB_ERR_ARG;
    const size_t gcount = (size_t)d.gw * d.gh * (c->p.disp_max + 2);
    int32_t *tmp = nullptr;
    SVB_TRY(dev_alloc(&tmp, gcount));
    int r = upload_support_and_tris(c, L, support, n, nullptr, 0, nullptr, 0);
    if (r == SVB_OK) r = launch_grid(d, c->p, L.support, L.nsupport, L.grid_tmp, L.grid[0], L.grid[1], 1, n, L.stream);
    if (r == SVB_OK) r = launch_grid_expand(d, c->p, L.grid[right_image ? 1 : 0], tmp, L.stream);
    if (r == SVB_OK && cudaMemcpyAsync(grid, tmp, gcount * 4, cudaMemcpyDeviceToHost, L.stream) != cudaSuccess) r = SVB_ERR_CUDA;
    cudaStreamSynchronize(L.stream);
    cudaFree(tmp);
    return r;
}

int svb_stage_disparity(svb_context *c, const int32_t *support, int n, const int32_t *tri, int m, const uint8_t *desc1, const uint8_t *desc2,
                        int right_image, float *D) {
    STAGE_PROLOG();
    if (!support || !tri || !desc1 || !desc2 || !D) return SVB_ERR_ARG;
    const size_t C = (size_t)c->chunk;
    SVB_CUDA(cudaMemcpyAsync(L.desc[0], desc1, N * 16, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpyAsync(L.desc[1], desc2, N * 16, cudaMemcpyHostToDevice, L.stream));
    // the same list is used for both sides; only the requested side is read back
    SVB_TRY(upload_support_and_tris(c, L, support, n, tri, m, tri, m));
    SVB_TRY(launch_planes(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, nullptr, nullptr, L.rec[0], L.rec[1], 1, m, L.stream));
    SVB_TRY(launch_grid(d, c->p, L.support, L.nsupport, L.grid_tmp, L.grid[0], L.grid[1], 1, n, L.stream));
    SVB_TRY(launch_raster(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, L.owner[0], L.owner[1], 1, m, L.stream));
    SVB_TRY(launch_dense(d, c->p, L.desc[0], L.desc[1], L.owner[0], L.owner[1], L.rec[0], L.rec[1], L.grid[0], L.grid[1], L.Draw, L.Draw + C * N,
                         1, L.stream));
    SVB_CUDA(cudaMemcpyAsync(D, right_image ? L.Draw + C * N : L.Draw, (size_t)d.DN * 4, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    return SVB_OK;
}

int svb_stage_lr_check(svb_context *c, float *D1, float *D2) {
    STAGE_PROLOG();
    if (!D1 || !D2) return SVB_ERR_ARG;
    const size_t C = (size_t)c->chunk;
    const size_t DB = (size_t)d.DN * 4;
    SVB_CUDA(cudaMemcpyAsync(L.Draw, D1, DB, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpyAsync(L.Draw + C * N, D2, DB, cudaMemcpyHostToDevice, L.stream));
    SVB_TRY(launch_lr_check(d, c->p, L.Draw, L.Draw + C * N, L.Dlr, L.Dlr + C * N, 1, L.stream));
    SVB_CUDA(cudaMemcpyAsync(D1, L.Dlr, DB, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaMemcpyAsync(D2, L.Dlr + C * N, DB, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    return SVB_OK;
}

static int stage_inplace(svb_context *c, float *D, int which) {
    STAGE_PROLOG();
    if (!D) return SVB_ERR_ARG;
    SVB_CUDA(cudaMemcpyAsync(L.Dlr, D, (size_t)d.DN * 4, cudaMemcpyHostToDevice, L.stream));
    if (which == 0) SVB_TRY(launch_remove_small_segments(d, c->p, L.Dlr, L.labels, L.sizes, L.ccl_roots, L.ccl_counts, 1, L.stream));
    if (which == 1) SVB_TRY(launch_gap(d, c->p, L.Dlr, reinterpret_cast<uint32_t *>(L.Dtmp), nullptr, nullptr, 1, L.stream));
    if (which == 2) SVB_TRY(launch_adaptive_mean(d, c->mean_mode, L.Dlr, L.Dtmp, 1, L.stream));
    if (which == 3) SVB_TRY(launch_median(d, L.Dlr, L.Dtmp, 1, L.stream));
    SVB_CUDA(cudaMemcpyAsync(D, L.Dlr, (size_t)d.DN * 4, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    return SVB_OK;
}
int svb_stage_remove_small_segments(svb_context *c, float *D) { return stage_inplace(c, D, 0); }
int svb_stage_gap_interpolation(svb_context *c, float *D) { return stage_inplace(c, D, 1); }
int svb_stage_adaptive_mean(svb_context *c, float *D) { return stage_inplace(c, D, 2); }
int svb_stage_median(svb_context *c, float *D) { return stage_inplace(c, D, 3); }

int svb_stage_reproject(svb_context *c, const float *D, const double *Q16, const double *XR9, const double *XT3, uint8_t *dmap_out,
                        double *points_out) {
    STAGE_PROLOG();
    if (!D || !Q16 || !points_out) return SVB_ERR_ARG;
    Calib cal;
    memcpy(cal.Q, Q16, sizeof(cal.Q));
    memset(cal.XR, 0, sizeof(cal.XR));
    memset(cal.XT, 0, sizeof(cal.XT));
    for (int i = 0; i < 3; i++) cal.XR[4 * i] = 1.0;
    if (XR9) memcpy(cal.XR, XR9, sizeof(cal.XR));
    if (XT3) memcpy(cal.XT, XT3, sizeof(cal.XT));
    double *pts = nullptr;
    SVB_TRY(dev_alloc(&pts, N * 3));
    int r = SVB_OK;
    if (cudaMemcpyAsync(L.Dlr, D, N * 4, cudaMemcpyHostToDevice, L.stream) != cudaSuccess) r = SVB_ERR_CUDA;
    if (r == SVB_OK) r = launch_reproject(d, cal, L.Dlr, L.dmap, pts, 1, L.stream);
    if (r == SVB_OK && cudaMemcpyAsync(points_out, pts, N * 24, cudaMemcpyDeviceToHost, L.stream) != cudaSuccess) r = SVB_ERR_CUDA;
    if (r == SVB_OK && dmap_out && cudaMemcpyAsync(dmap_out, L.dmap, N, cudaMemcpyDeviceToHost, L.stream) != cudaSuccess) r = SVB_ERR_CUDA;
    cudaStreamSynchronize(L.stream);
    cudaFree(pts);
    if (r == SVB_ERR_CUDA) set_error("svb_stage_reproject: CUDA failure: %s", cudaGetErrorString(cudaGetLastError()));
    return r;
}

// publishPointCloud on a u8 map of any size (no context: the map need not have the size a context was created for)
int svb_reproject_u8(const uint8_t *dmap, int width, int height, const double *Q16, const double *XR9, const double *XT3, double *points_out) {
    if (!dmap || !Q16 || !points_out || width < 1 || height < 1) return SVB_ERR_ARG;
    Calib cal;
    memcpy(cal.Q, Q16, sizeof(cal.Q));
    memset(cal.XR, 0, sizeof(cal.XR));
    memset(cal.XT, 0, sizeof(cal.XT));
    for (int i = 0; i < 3; i++) cal.XR[4 * i] = 1.0;
    if (XR9) memcpy(cal.XR, XR9, sizeof(cal.XR));
    if (XT3) memcpy(cal.XT, XT3, sizeof(cal.XT));
    const size_t N = (size_t)width * height;
    uint8_t *d_map = nullptr;
    double *d_pts = nullptr;
    SVB_CUDA(cudaMalloc((void **)&d_map, N));
    cudaError_t e = cudaMalloc((void **)&d_pts, N * 24);
    if (e == cudaSuccess) e = cudaMemcpy(d_map, dmap, N, cudaMemcpyHostToDevice);
    int r = SVB_OK;
    if (e == cudaSuccess) r = launch_reproject_u8(cal, d_map, d_pts, width, height, nullptr);
    if (e == cudaSuccess && r == SVB_OK) e = cudaMemcpy(points_out, d_pts, N * 24, cudaMemcpyDeviceToHost);
    cudaFree(d_map);
    cudaFree(d_pts);
    if (e != cudaSuccess) {
        set_error("svb_reproject_u8: %s", cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    return r;
}

// ---- batch pipeline -----------------------------------------------------------------------------------------
static int ensure_store(void **p, size_t *have_frames, size_t want_frames, size_t bytes_per_frame) {
    if (*p && *have_frames >= want_frames) return SVB_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have_frames = 0;
    cudaError_t e = cudaMalloc(p, want_frames * bytes_per_frame + 256);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu bytes) for a batch store: %s", want_frames * bytes_per_frame, cudaGetErrorString(e));
        return SVB_ERR_CUDA;
    }
    *have_frames = want_frames;
    return SVB_OK;
}

int svb_batch_upload(svb_context *c, const uint8_t *left, const uint8_t *right, int n_frames) {
    if (!c || !left || !right || n_frames < 1) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    const size_t N = (size_t)c->d.N;
    if (!(c->in_img[0] && c->in_frames >= (size_t)n_frames)) {
        for (int s = 0; s < 2; s++) {
            if (c->in_img[s]) cudaFree(c->in_img[s]);
            c->in_img[s] = nullptr;
        }
        c->in_frames = 0;
        for (int s = 0; s < 2; s++) {
            SVB_TRY(dev_alloc(&c->in_img[s], (size_t)n_frames * c->d.IN));
            SVB_CUDA(cudaMemset(c->in_img[s], 0, (size_t)n_frames * c->d.IN));
        }
        c->in_frames = n_frames;
    }
    // tight rows on the host, Dims::bpl bytes per line on the device; the rows of consecutive frames follow one another in both
    SVB_CUDA(cudaMemcpy2D(c->in_img[0], c->d.bpl, left, c->d.W, c->d.W, (size_t)n_frames * c->d.H, cudaMemcpyHostToDevice));
    SVB_CUDA(cudaMemcpy2D(c->in_img[1], c->d.bpl, right, c->d.W, c->d.W, (size_t)n_frames * c->d.H, cudaMemcpyHostToDevice));
    return SVB_OK;
}

// The input side of generatePointCloud for a whole batch: BGRA frames (height x width x 4 bytes, what sv.py hands over, sv.py:185-188)
// go up in groups of `chunk` frames and are converted to gray on the device (cv::cvtColor(BGRA2GRAY), stereo_vision.cu:346-347) straight
// into the resident input store.
int svb_batch_upload_bgra(svb_context *c, const uint8_t *left_bgra, const uint8_t *right_bgra, int n_frames) {
    if (!c || !left_bgra || !right_bgra || n_frames < 1) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    const size_t N = (size_t)c->d.N;
    if (!(c->in_img[0] && c->in_frames >= (size_t)n_frames)) {
        for (int s = 0; s < 2; s++) {
            if (c->in_img[s]) cudaFree(c->in_img[s]);
            c->in_img[s] = nullptr;
        }
        c->in_frames = 0;
        for (int s = 0; s < 2; s++) {
            SVB_TRY(dev_alloc(&c->in_img[s], (size_t)n_frames * c->d.IN));
            SVB_CUDA(cudaMemset(c->in_img[s], 0, (size_t)n_frames * c->d.IN));
        }
        c->in_frames = n_frames;
    }
    const int G = c->chunk;  // frames per staging group
    if (c->bgra_batch_frames < (size_t)G) {
        for (int s = 0; s < 2; s++) {
            if (c->bgra_batch[s]) cudaFree(c->bgra_batch[s]);
            c->bgra_batch[s] = nullptr;
        }
        c->bgra_batch_frames = 0;
        for (int s = 0; s < 2; s++) SVB_TRY(dev_alloc(&c->bgra_batch[s], (size_t)G * N * 4));
        c->bgra_batch_frames = G;
    }
    cudaStream_t st = c->lanes[0].stream;
    for (int f0 = 0; f0 < n_frames; f0 += G) {
        const int nf = std::min(G, n_frames - f0);
        const uint8_t *src[2] = {left_bgra, right_bgra};
        for (int s = 0; s < 2; s++) {
            SVB_CUDA(cudaMemcpyAsync(c->bgra_batch[s], src[s] + (size_t)f0 * N * 4, (size_t)nf * N * 4, cudaMemcpyHostToDevice, st));
            // one launch converts the whole group: the rows of its frames follow one another in the staging buffer and in the store
            SVB_TRY(launch_bgra_to_gray(c->bgra_batch[s], c->in_img[s] + (size_t)f0 * c->d.IN, c->d.W, nf * c->d.H, c->d.bpl, st));
        }
    }
    SVB_CUDA(cudaStreamSynchronize(st));
    return SVB_OK;
}

// Device pointers of the last batch call's resident results, so that a consumer on the same GPU (another CUDA library, a renderer, the
// next stage of a perception stack) reads them in place instead of paying 13 MB of PCIe per frame: D1 = n_frames x (H x W) float,
// points = n_frames x (H x W) x {x, y, z} double (either may come back NULL when the call did not produce it).  Valid until the next
// batch call on this context or svb_destroy; work of the producing call is complete when svb_batch_run returns.
int svb_batch_device_ptrs(svb_context *c, float **D1_dev, double **points_dev, int *n_frames, int *device) {
    if (!c) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (D1_dev) *D1_dev = c->last_want_D ? c->out_D1 : nullptr;
    if (points_dev) *points_dev = c->last_want_P ? c->out_points : nullptr;
    if (n_frames) *n_frames = (int)c->stats.frames;
    if (device) *device = c->device;
    return SVB_OK;
}

// Shared driver of the resident and the host-buffer (e2e) batch paths.
static int batch_drive(svb_context *c, int n_frames, int flags, const uint8_t *h_left, const uint8_t *h_right, float *h_D1, double *h_points) {
    const Dims &d = c->d;
    const size_t N = (size_t)d.N;
    const int C = c->chunk;
    const bool from_host = h_left != nullptr;
    const bool want_D = (flags & SVB_OUT_DISPARITY) != 0, want_P = (flags & (SVB_OUT_POINTS | SVB_OUT_POINTS_FLOATDISP)) != 0;
    if ((flags & SVB_OUT_POINTS) && (flags & SVB_OUT_POINTS_FLOATDISP)) {
        set_error("batch: SVB_OUT_POINTS and SVB_OUT_POINTS_FLOATDISP share the point-cloud store; pick one");
        return SVB_ERR_ARG;
    }
    c->points_float_disp = (flags & SVB_OUT_POINTS_FLOATDISP) != 0;
    c->last_want_D = want_D;
    c->last_want_P = want_P;
    const size_t DN = (size_t)d.DN;
    if (want_P && d.sub) {
        set_error("batch: point clouds with subsampling are only defined for the single-frame generatePointCloud path");
        return SVB_ERR_UNSUPPORTED;
    }
    if (!want_D && !want_P) {
        set_error("batch: flags select no output");
        return SVB_ERR_ARG;
    }
    stats_reset(c);
    c->frame_nsupport.assign((size_t)n_frames, 0);
    if (want_D) SVB_TRY(ensure_store((void **)&c->out_D1, &c->out_D1_frames, n_frames, DN * 4));
    if (want_P) SVB_TRY(ensure_store((void **)&c->out_points, &c->out_points_frames, n_frames, N * 24));
    const int nchunks = (n_frames + C - 1) / C;
    SVB_TRY(stage_events_prepare(c, nchunks));
    cudaEvent_t ev0, ev1;
    SVB_CUDA(cudaEventCreate(&ev0));
    SVB_CUDA(cudaEventCreate(&ev1));
    SVB_CUDA(cudaDeviceSynchronize());
    SVB_CUDA(cudaEventRecord(ev0, c->lanes[0].stream));
    const int LANES = c->n_lanes;
    for (int l = 1; l < LANES; l++) SVB_CUDA(cudaStreamWaitEvent(c->lanes[l].stream, ev0, 0));

    auto frames_of = [&](int k) { return (k + 1) * C <= n_frames ? C : n_frames - k * C; };
    auto issue_a = [&](int k) -> int {
        Lane &L = c->lanes[k % LANES];
        const int nf = frames_of(k);
        const size_t off = (size_t)k * C * N;
        L.first_frame = k * C;
        if (from_host) {
            // over PCIe as ONE contiguous transfer per side (a pitched host-to-device copy is issued row by row and measured 15 % slower
            // end to end), then re-pitched to Dims::bpl bytes per line on the device
            const uint8_t *h_src[2] = {h_left + off, h_right + off};
            for (int s = 0; s < 2; s++) {
                if (!L.img_tight[s]) SVB_TRY(dev_alloc(&L.img_tight[s], (size_t)C * N));
                SVB_CUDA(cudaMemcpyAsync(L.img_tight[s], h_src[s], (size_t)nf * N, cudaMemcpyHostToDevice, L.stream));
                SVB_CUDA(cudaMemcpy2DAsync(L.img[s], d.bpl, L.img_tight[s], d.W, d.W, (size_t)nf * d.H, cudaMemcpyDeviceToDevice, L.stream));
            }
            return stage_a(c, L, L.img[0], L.img[1], nf, stage_events_of(c, k));
        }
        const size_t off_dev = (size_t)k * C * d.IN;
        return stage_a(c, L, c->in_img[0] + off_dev, c->in_img[1] + off_dev, nf, stage_events_of(c, k));
    };
    // The host stage runs on its own thread, one chunk ahead of the launches: while this thread queues stage B of
    // chunk k and stage A of chunk k + LANES, the Delaunay workers already triangulate chunk k + 1.  `issued` counts
    // the chunks whose stage A (and with it the event the host stage waits for) has been queued.
    struct HostStage {
        std::mutex mu;
        std::condition_variable cv;
        int issued = 0, done = 0, err = SVB_OK;
        bool stop = false;
        char msg[1024] = "";  // the error text is thread-local: carried over to the calling thread
    } hs;
    std::thread host_thread([&] {
        cudaSetDevice(c->device);
        for (int k = 0; k < nchunks; k++) {
            {
                std::unique_lock<std::mutex> lk(hs.mu);
                hs.cv.wait(lk, [&] { return hs.issued > k || hs.stop; });
                if (hs.stop) return;
            }
            const int rc = stage_host(c, c->lanes[k % LANES], frames_of(k), c->lanes[k % LANES].unpacked);
            std::lock_guard<std::mutex> lk(hs.mu);
            if (rc != SVB_OK) {
                hs.err = rc;
                snprintf(hs.msg, sizeof(hs.msg), "%s", g_err);
                hs.done = nchunks;
                hs.cv.notify_all();
                return;
            }
            hs.done = k + 1;
            hs.cv.notify_all();
        }
    });
    auto drive = [&]() -> int {
        for (int k = 0; k < nchunks && k < LANES; k++) {
            SVB_TRY(issue_a(k));
            std::lock_guard<std::mutex> lk(hs.mu);
            hs.issued = k + 1;
            hs.cv.notify_all();
        }
        for (int k = 0; k < nchunks; k++) {
            Lane &L = c->lanes[k % LANES];
            const int nf = frames_of(k);
            {
                std::unique_lock<std::mutex> lk(hs.mu);
                hs.cv.wait(lk, [&] { return hs.done > k; });
                if (hs.err != SVB_OK) {
                    set_error("%s", hs.msg);
                    return hs.err;
                }
            }
            const size_t off = (size_t)k * C * N, offD = (size_t)k * C * DN;
            SVB_TRY(stage_b(c, L, nf, want_D ? c->out_D1 + offD : nullptr, want_P ? c->out_points + off * 3 : nullptr, stage_events_of(c, k)));
            if (from_host) {
                if (want_D && h_D1) SVB_CUDA(cudaMemcpyAsync(h_D1 + offD, c->out_D1 + offD, nf * DN * 4, cudaMemcpyDeviceToHost, L.stream));
                if (want_P && h_points)
                    SVB_CUDA(cudaMemcpyAsync(h_points + off * 3, c->out_points + off * 3, nf * N * 24, cudaMemcpyDeviceToHost, L.stream));
            }
            if (k + LANES < nchunks) {
                SVB_TRY(issue_a(k + LANES));
                std::lock_guard<std::mutex> lk(hs.mu);
                hs.issued = k + LANES + 1;
                hs.cv.notify_all();
            }
        }
        return SVB_OK;
    };
    const int drive_rc = drive();
    {
        std::lock_guard<std::mutex> lk(hs.mu);
        hs.stop = true;
        hs.cv.notify_all();
    }
    host_thread.join();
    for (int l = 0; l < LANES; l++) c->lanes[l].first_frame = -1;
    if (drive_rc != SVB_OK) {
        cudaDeviceSynchronize();
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        return drive_rc;
    }
    for (int l = 1; l < LANES; l++) {
        SVB_CUDA(cudaEventRecord(c->lanes[l].ev_done, c->lanes[l].stream));
        SVB_CUDA(cudaStreamWaitEvent(c->lanes[0].stream, c->lanes[l].ev_done, 0));
    }
    SVB_CUDA(cudaEventRecord(ev1, c->lanes[0].stream));
    SVB_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    SVB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    stage_events_collect(c);
    c->stats.gpu_ms_total = ms;
    c->stats.frames = n_frames;
    c->stats.kernel_launches = g_launch_counter;
    SVB_CUDA(cudaGetLastError());
    return SVB_OK;
}

int svb_batch_run(svb_context *c, int n_frames, int flags) {
    if (!c || n_frames < 1) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    if (!c->in_img[0] || c->in_frames < (size_t)n_frames) {
        set_error("svb_batch_run: only %zu frames are resident (upload first)", c->in_frames);
        return SVB_ERR_ARG;
    }
    return batch_drive(c, n_frames, flags, nullptr, nullptr, nullptr, nullptr);
}

int svb_batch_run_host(svb_context *c, const uint8_t *left, const uint8_t *right, int n_frames, int flags, float *D1_out, double *points_out) {
    if (!c || !left || !right || n_frames < 1) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    return batch_drive(c, n_frames, flags, left, right, D1_out, points_out);
}

int svb_batch_frame_support(svb_context *c, int32_t *nsupport_out, int n_frames) {
    if (!c || !nsupport_out || n_frames < 0) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if ((size_t)n_frames > c->frame_nsupport.size()) {
        set_error("svb_batch_frame_support: the last batch call had %zu frames", c->frame_nsupport.size());
        return SVB_ERR_ARG;
    }
    memcpy(nsupport_out, c->frame_nsupport.data(), sizeof(int32_t) * (size_t)n_frames);
    return SVB_OK;
}

int svb_batch_download_disparity(svb_context *c, int frame, float *out) {
    if (!c || !out || frame < 0) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if ((size_t)frame >= c->out_D1_frames || !c->out_D1) return SVB_ERR_ARG;
    SVB_CUDA(cudaSetDevice(c->device));
    SVB_CUDA(cudaMemcpy(out, c->out_D1 + (size_t)frame * c->d.DN, (size_t)c->d.DN * 4, cudaMemcpyDeviceToHost));
    return SVB_OK;
}

int svb_batch_download_points(svb_context *c, int frame, double *out) {
    if (!c || !out || frame < 0) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if ((size_t)frame >= c->out_points_frames || !c->out_points) return SVB_ERR_ARG;
    SVB_CUDA(cudaSetDevice(c->device));
    SVB_CUDA(cudaMemcpy(out, c->out_points + (size_t)frame * c->d.N * 3, (size_t)c->d.N * 24, cudaMemcpyDeviceToHost));
    return SVB_OK;
}

// ---- generatePointCloud body: BGRA in, double3 point cloud out (stereo_vision.cu:596-618) ------------------------
int svb_point_cloud_bgra(svb_context *c, const uint8_t *left_bgra, const uint8_t *right_bgra, double *points_out, uint8_t *dmap_out, float *D1_out,
                         double *times_ms) {
    if (!c || !left_bgra || !right_bgra || !points_out) {
        set_error("svb_point_cloud_bgra: bad argument");
        return SVB_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    SVB_CUDA(cudaSetDevice(c->device));
    stats_reset(c);
    const Dims &d = c->d;
    Lane &L = c->lanes[0];
    const size_t N = (size_t)d.N;
    for (int s = 0; s < 2; s++)
        if (!c->bgra[s]) SVB_TRY(dev_alloc(&c->bgra[s], N * 4));
    for (int i = 0; i < 4; i++)
        if (!c->ev_pc[i]) SVB_CUDA(cudaEventCreate(&c->ev_pc[i]));
    SVB_TRY(ensure_store((void **)&c->out_points, &c->out_points_frames, 1, N * 24));
    SVB_CUDA(cudaMemcpyAsync(c->bgra[0], left_bgra, N * 4, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaMemcpyAsync(c->bgra[1], right_bgra, N * 4, cudaMemcpyHostToDevice, L.stream));
    SVB_CUDA(cudaEventRecord(c->ev_pc[0], L.stream));
    // imgCallback_video(): cvtColor(BGRA2GRAY) of both images (stereo_vision.cu:346-347)
    SVB_TRY(launch_bgra_to_gray(c->bgra[0], L.img[0], d.W, d.H, d.bpl, L.stream));
    SVB_TRY(launch_bgra_to_gray(c->bgra[1], L.img[1], d.W, d.H, d.bpl, L.stream));
    SVB_TRY(stage_events_prepare(c, 1));
    StageEvents *se = stage_events_of(c, 0);
    SVB_TRY(stage_a(c, L, L.img[0], L.img[1], 1, se));
    SVB_TRY(stage_host(c, L, 1, L.unpacked));
    c->stats.frames = 1;
    int rc = SVB_OK;
    if (L.h_nsupport[0] < 3) {
        // elas.cpp:64-69 leaves the zero-initialised leftdpf alone (stereo_vision.cu:312): disparity 0 everywhere
        set_error("need at least 3 support points (got %d)", L.h_nsupport[0]);
        rc = SVB_ERR_FEW_SUPPORT;
        SVB_CUDA(cudaMemsetAsync(L.Dlr, 0, N * 4, L.stream));
        SVB_CUDA(cudaEventRecord(c->ev_pc[1], L.stream));
        SVB_TRY(launch_reproject(d, c->calib, L.Dlr, L.dmap, c->out_points, 1, L.stream));
    } else {
        // generateDisparityMap(): Elas::process, then convertTo(CV_8UC1, 4.0) and publishPointCloud()'s kernel.  At full resolution
        // both are fused with the last filters (k_post_fused.cu), so the "point-cloud part" of times_ms is the copy of the cloud only;
        // with subsampling the reference projects the full-size buffer whose first (W/2)*(H/2) floats hold the map: stage by stage.
        c->points_float_disp = false;
        if (d.sub || !c->fused_post) {
            SVB_TRY(stage_b(c, L, 1, nullptr, nullptr, se));
            SVB_CUDA(cudaEventRecord(c->ev_pc[1], L.stream));
            SVB_TRY(launch_reproject(d, c->calib, L.Dlr, L.dmap, c->out_points, 1, L.stream));
        } else {
            SVB_TRY(stage_b(c, L, 1, nullptr, c->out_points, se));
            SVB_CUDA(cudaEventRecord(c->ev_pc[1], L.stream));
        }
    }
    SVB_CUDA(cudaMemcpyAsync(points_out, c->out_points, N * 24, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaEventRecord(c->ev_pc[2], L.stream));
    if (dmap_out) SVB_CUDA(cudaMemcpyAsync(dmap_out, L.dmap, N, cudaMemcpyDeviceToHost, L.stream));
    if (D1_out) SVB_CUDA(cudaMemcpyAsync(D1_out, L.Dlr, N * 4, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    if (times_ms) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, c->ev_pc[0], c->ev_pc[1]);
        cudaEventElapsedTime(&b, c->ev_pc[1], c->ev_pc[2]);
        times_ms[0] = a;
        times_ms[1] = b;
    }
    stage_events_collect(c);
    c->stats.kernel_launches = g_launch_counter;
    return rc;
}

// BGRA -> gray on its own (parity harness for the colour conversion)
int svb_stage_bgra_to_gray(svb_context *c, const uint8_t *bgra, uint8_t *gray_out) {
    STAGE_PROLOG();
    if (!bgra || !gray_out) return SVB_ERR_ARG;
    if (!c->bgra[0]) SVB_TRY(dev_alloc(&c->bgra[0], N * 4));
    SVB_CUDA(cudaMemcpyAsync(c->bgra[0], bgra, N * 4, cudaMemcpyHostToDevice, L.stream));
    SVB_TRY(launch_bgra_to_gray(c->bgra[0], L.img[0], d.W, d.H, d.bpl, L.stream));
    SVB_CUDA(cudaMemcpy2DAsync(gray_out, d.W, L.img[0], d.bpl, d.W, d.H, cudaMemcpyDeviceToHost, L.stream));
    SVB_CUDA(cudaStreamSynchronize(L.stream));
    return SVB_OK;
}

int svb_get_stats(svb_context *c, svb_stats *out) {
    if (!c || !out) return SVB_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    *out = c->stats;
    return SVB_OK;
}

const char *svb_stage_name(int id) { return (id >= 0 && id < ST_COUNT) ? kStageNames[id] : ""; }

void *svb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void svb_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
