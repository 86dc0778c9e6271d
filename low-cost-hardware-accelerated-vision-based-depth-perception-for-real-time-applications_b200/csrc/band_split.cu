// Row-band split of ONE frame over several GPUs with NVLink peer-to-peer exchange (BASELINE.json configs[3]:
// 3840x2160, disparity range 512).  New relative to the reference, which is single-GPU; the result is bit-identical
// to Elas::process on one device (SURVEY.md 8e describes what is row-local and what is global).
//
// Device b owns image rows [r0_b, r1_b).  One process, one stream per device, cross-device ordering by CUDA events:
//   1. H2D   image rows [r0-3, r1+3) of both images (the 3-row Sobel/descriptor input halo is re-read from the host)
//   2. GPU b descriptors of rows [r0, r1)
//   3. P2P   the two descriptor halo rows above and below the band are fetched from the neighbours' HBM
//            (support matching reads anchors at v -+ 2): 2 images x 2 sides x 2 rows x 16 W bytes per band edge
//   4. GPU b support matching of the lattice rows that fall into the band
//   5. P2P   lattice rows -> device 0 (a few hundred KB); device 0 runs the (global, order-dependent) lattice filters
//            and writes the support list into mapped host memory
//   6. host  Delaunay (both triangulations)
//   7. H2D   support list + triangle lists to every device
//   8. GPU b planes, grid (whole frame: tiny), owner map + dense matching + L/R check of the band's rows
//   9. P2P   band rows of both L/R-checked maps -> device 0
//  10. GPU 0 speckle removal, gap interpolation, adaptive mean, median on the whole frame (CCL and the column pass
//            are global along columns), D2H
// Peer copies are cudaMemcpyPeerAsync: over NVLink / NVSwitch when peer access is enabled, staged by the driver
// otherwise; the same device may appear several times in the list (used to test the logic on one GPU).
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "pipeline_internal.h"

using namespace svb;

struct svb_band_group {
    svb_params p;
    Dims d;
    int n = 0;
    std::vector<svb_context *> ctx;
    std::vector<int> r0, r1;      // image rows of band b
    std::vector<int> vc0, vc1;    // lattice rows of band b
    std::vector<cudaEvent_t> ev_desc, ev_match, ev_lr, ev_up;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    svb_band_stats stats;
};

namespace {

int peer_copy(void *dst, int dst_dev, const void *src, int src_dev, size_t bytes, cudaStream_t s, svb_band_group *g) {
    if (!bytes) return SVB_OK;
    SVB_CUDA(cudaMemcpyPeerAsync(dst, dst_dev, src, src_dev, bytes, s));
    g->stats.p2p_bytes += (int64_t)bytes;
    g->stats.p2p_copies++;
    return SVB_OK;
}

}  // namespace

extern "C" {

svb_band_group *svb_band_create(const svb_params *params, int width, int height, const int *devices, int n_devices) {
    if (!params || !devices || n_devices < 1 || n_devices > 16) {
        set_error("svb_band_create: bad argument");
        return nullptr;
    }
    svb_band_group *g = new svb_band_group();
    g->p = *params;
    g->n = n_devices;
    if (make_dims(g->p, width, height, &g->d) != SVB_OK) {
        delete g;
        return nullptr;
    }
    if (g->d.sub) {
        set_error("svb_band_create: subsampling is not supported by the row-band split");
        delete g;
        return nullptr;
    }
    if (height < 16 * n_devices) {
        set_error("svb_band_create: %d rows are too few for %d bands", height, n_devices);
        delete g;
        return nullptr;
    }
    memset(&g->stats, 0, sizeof(g->stats));
    for (int b = 0; b < n_devices; b++) {
        svb_context *c = svb_create(params, width, height, 1, devices[b]);
        if (!c) {
            for (auto *x : g->ctx) svb_destroy(x);
            delete g;
            return nullptr;
        }
        g->ctx.push_back(c);
    }
    // peer access between every pair of distinct devices (NVLink 5 / NVSwitch: every peer at full bandwidth)
    for (int a = 0; a < n_devices; a++)
        for (int b = 0; b < n_devices; b++) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (can) {
                cudaSetDevice(devices[a]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
                if (e == cudaSuccess) g->stats.peer_links++;
                else if (e == cudaErrorPeerAccessAlreadyEnabled) {
                    g->stats.peer_links++;
                    cudaGetLastError();
                } else
                    cudaGetLastError();
            }
        }
    const Dims &d = g->d;
    for (int b = 0; b < n_devices; b++) {
        const int r0 = (int)((long long)height * b / n_devices), r1 = (int)((long long)height * (b + 1) / n_devices);
        g->r0.push_back(r0);
        g->r1.push_back(r1);
        // lattice rows vc >= 1 whose image row vc * step lies in [r0, r1)
        int v0 = std::max(1, (r0 + d.step - 1) / d.step), v1 = std::min(d.ch, (r1 + d.step - 1) / d.step);
        if (v1 < v0) v1 = v0;
        g->vc0.push_back(v0);
        g->vc1.push_back(v1);
    }
    g->ev_desc.resize(n_devices);
    g->ev_match.resize(n_devices);
    g->ev_lr.resize(n_devices);
    g->ev_up.resize(n_devices);
    for (int b = 0; b < n_devices; b++) {
        cudaSetDevice(devices[b]);
        cudaEventCreateWithFlags(&g->ev_desc[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g->ev_match[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g->ev_lr[b], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&g->ev_up[b], cudaEventDisableTiming);
    }
    cudaSetDevice(devices[0]);
    cudaEventCreate(&g->ev_begin);
    cudaEventCreate(&g->ev_end);
    return g;
}

void svb_band_destroy(svb_band_group *g) {
    if (!g) return;
    for (int b = 0; b < g->n; b++) {
        cudaSetDevice(g->ctx[b]->device);
        cudaDeviceSynchronize();
        cudaEventDestroy(g->ev_desc[b]);
        cudaEventDestroy(g->ev_match[b]);
        cudaEventDestroy(g->ev_lr[b]);
        cudaEventDestroy(g->ev_up[b]);
    }
    cudaEventDestroy(g->ev_begin);
    cudaEventDestroy(g->ev_end);
    for (auto *c : g->ctx) svb_destroy(c);
    delete g;
}

int svb_band_get_stats(svb_band_group *g, svb_band_stats *out) {
    if (!g || !out) return SVB_ERR_ARG;
    *out = g->stats;
    return SVB_OK;
}

int svb_band_process(svb_band_group *g, const uint8_t *I1, const uint8_t *I2, int stride, float *D1, float *D2) {
    if (!g || !I1 || !I2 || !D1 || !D2 || stride < g->d.W) {
        set_error("svb_band_process: bad argument");
        return SVB_ERR_ARG;
    }
    const Dims &d = g->d;
    const svb_params &p = g->p;
    const int n = g->n, W = d.W, H = d.H;
    const size_t N = (size_t)d.N;
    const int peer_links = g->stats.peer_links;
    memset(&g->stats, 0, sizeof(g->stats));
    g->stats.peer_links = peer_links;
    g->stats.bands = n;
    const auto t_begin = std::chrono::steady_clock::now();
    svb_context *c0 = g->ctx[0];
    Lane &L0 = c0->lanes[0];
    SVB_CUDA(cudaSetDevice(c0->device));
    SVB_CUDA(cudaEventRecord(g->ev_begin, L0.stream));

    // ---- 1-2: input rows and descriptors of every band -------------------------------------------------------
    for (int b = 0; b < n; b++) {
        svb_context *c = g->ctx[b];
        Lane &L = c->lanes[0];
        SVB_CUDA(cudaSetDevice(c->device));
        const int i0 = std::max(g->r0[b] - 3, 0), i1 = std::min(g->r1[b] + 3, H);
        const uint8_t *src[2] = {I1, I2};
        for (int s = 0; s < 2; s++)
            SVB_CUDA(cudaMemcpy2DAsync(L.img[s] + (size_t)i0 * d.bpl, d.bpl, src[s] + (size_t)i0 * stride, stride, W, i1 - i0, cudaMemcpyHostToDevice,
                                       L.stream));
        for (int s = 0; s < 2; s++) SVB_TRY(launch_descriptor_rows(d, L.img[s], L.desc[s], 1, g->r0[b], g->r1[b], L.stream));
        SVB_CUDA(cudaEventRecord(g->ev_desc[b], L.stream));
    }
    // ---- 3-4: descriptor halo rows from the neighbours over P2P, then support matching of the band ------------
    for (int b = 0; b < n; b++) {
        svb_context *c = g->ctx[b];
        Lane &L = c->lanes[0];
        SVB_CUDA(cudaSetDevice(c->device));
        for (int nb = b - 1; nb <= b + 1; nb += 2) {
            if (nb < 0 || nb >= n) continue;
            SVB_CUDA(cudaStreamWaitEvent(L.stream, g->ev_desc[nb], 0));
            // rows [r0-2, r0) live on the band above, rows [r1, r1+2) on the band below
            const int h0 = nb < b ? std::max(g->r0[b] - 2, g->r0[nb]) : g->r1[b];
            const int h1 = nb < b ? g->r0[b] : std::min(g->r1[b] + 2, g->r1[nb]);
            Lane &Ln = g->ctx[nb]->lanes[0];
            for (int s = 0; s < 2 && h1 > h0; s++)
                SVB_TRY(peer_copy(L.desc[s] + (size_t)h0 * W * 16, c->device, Ln.desc[s] + (size_t)h0 * W * 16, g->ctx[nb]->device,
                                  (size_t)(h1 - h0) * W * 16, L.stream, g));
        }
        SVB_TRY(launch_support_match_rows(d, p, L.desc[0], L.desc[1], L.dcan_raw, 1, g->vc0[b], g->vc1[b], L.stream));
        SVB_CUDA(cudaEventRecord(g->ev_match[b], L.stream));
    }
    // ---- 5: lattice rows to device 0, global lattice filters there -------------------------------------------
    SVB_CUDA(cudaSetDevice(c0->device));
    for (int b = 1; b < n; b++) {
        SVB_CUDA(cudaStreamWaitEvent(L0.stream, g->ev_match[b], 0));
        const size_t off = (size_t)g->vc0[b] * d.cw, cnt = (size_t)(g->vc1[b] - g->vc0[b]) * d.cw;
        SVB_TRY(peer_copy(L0.dcan_raw + off, c0->device, g->ctx[b]->lanes[0].dcan_raw + off, g->ctx[b]->device, cnt * sizeof(int16_t), L0.stream,
                          g));
    }
    SVB_TRY(launch_dcan_border(d, L0.dcan_raw, 1, L0.stream));
    SVB_TRY(launch_support_filter(d, p, L0.dcan_raw, L0.dcan, L0.support, L0.nsupport, L0.h_support, L0.h_nsupport, L0.sf_changed, 1, L0.stream));
    SVB_CUDA(cudaEventRecord(L0.ev_a, L0.stream));
    // ---- 6: host Delaunay -----------------------------------------------------------------------------------
    c0->stats.delaunay_ms_total = c0->stats.delaunay_ms_wall = 0;
    SVB_TRY(stage_host(c0, L0, 1));
    g->stats.delaunay_ms = c0->stats.delaunay_ms_wall;
    const int ns = L0.h_nsupport[0];
    g->stats.support_points = ns;
    if (ns < 3) {
        for (int b = 0; b < n; b++) {
            SVB_CUDA(cudaSetDevice(g->ctx[b]->device));
            SVB_CUDA(cudaStreamSynchronize(g->ctx[b]->lanes[0].stream));
        }
        set_error("need at least 3 support points (got %d)", ns);
        return SVB_ERR_FEW_SUPPORT;  // elas.cpp:64-69: D1 / D2 untouched
    }
    const size_t C = 1;
    const int32_t *h_trioff = L0.h_ntri + 2 * C;
    const size_t tri_total = (size_t)h_trioff[1];
    const int max_tri = std::max(L0.h_ntri[0], L0.h_ntri[1]);
    // ---- 7-8: lists to every device; planes, grid, owner map, dense matching and L/R check of the band ----------
    for (int b = 0; b < n; b++) {
        svb_context *c = g->ctx[b];
        Lane &L = c->lanes[0];
        SVB_CUDA(cudaSetDevice(c->device));
        SVB_CUDA(cudaMemcpyAsync(L.nsupport, L0.h_nsupport, sizeof(int32_t), cudaMemcpyHostToDevice, L.stream));
        SVB_CUDA(cudaMemcpyAsync(L.support, L0.h_support, (size_t)ns * 12, cudaMemcpyHostToDevice, L.stream));
        SVB_CUDA(cudaMemcpyAsync(L.ntri, L0.h_ntri, sizeof(int32_t) * 3 * C, cudaMemcpyHostToDevice, L.stream));
        for (int s = 0; s < 2; s++)
            if (tri_total) SVB_CUDA(cudaMemcpyAsync(L.tri[s], L0.h_tri[s], sizeof(int32_t) * 3 * tri_total, cudaMemcpyHostToDevice, L.stream));
        SVB_TRY(launch_planes(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, nullptr, nullptr, L.rec[0], L.rec[1], 1, max_tri, L.stream));
        SVB_TRY(launch_grid(d, p, L.support, L.nsupport, L.grid_tmp, L.grid[0], L.grid[1], 1, ns, L.stream));
        SVB_TRY(launch_raster_rows(d, L.support, L.tri[0], L.tri[1], L.ntri, L.trioff, L.owner[0], L.owner[1], 1, max_tri, g->r0[b], g->r1[b],
                                   L.stream));
        float *D1raw = L.Draw, *D2raw = L.Draw + N;
        SVB_TRY(launch_dense_rows(d, p, L.desc[0], L.desc[1], L.owner[0], L.owner[1], L.rec[0], L.rec[1], L.grid[0], L.grid[1], D1raw, D2raw, 1,
                                  g->r0[b], g->r1[b], L.stream));
        SVB_TRY(launch_lr_check_rows(d, p, D1raw, D2raw, L.Dlr, L.Dlr + N, 1, g->r0[b], g->r1[b], L.stream));
        SVB_CUDA(cudaEventRecord(g->ev_lr[b], L.stream));
    }
    // ---- 9: gather the band rows of both maps on device 0 -----------------------------------------------------
    SVB_CUDA(cudaSetDevice(c0->device));
    for (int b = 1; b < n; b++) {
        SVB_CUDA(cudaStreamWaitEvent(L0.stream, g->ev_lr[b], 0));
        const size_t off = (size_t)g->r0[b] * W, cnt = (size_t)(g->r1[b] - g->r0[b]) * W * sizeof(float);
        Lane &Lb = g->ctx[b]->lanes[0];
        SVB_TRY(peer_copy(L0.Dlr + off, c0->device, Lb.Dlr + off, g->ctx[b]->device, cnt, L0.stream, g));
        SVB_TRY(peer_copy(L0.Dlr + N + off, c0->device, Lb.Dlr + N + off, g->ctx[b]->device, cnt, L0.stream, g));
    }
    // ---- 10: post-processing of the whole frame on device 0 (elas.cpp:111-135) ---------------------------------
    const int passes = p.postprocess_only_left ? 1 : 2;
    for (int s = 0; s < passes; s++) {
        float *D = L0.Dlr + (size_t)s * N;
        SVB_TRY(launch_remove_small_segments(d, p, D, L0.labels, L0.sizes, L0.ccl_roots, L0.ccl_counts, 1, L0.stream));
        SVB_TRY(launch_gap(d, p, D, reinterpret_cast<uint32_t *>(L0.Dtmp), nullptr, nullptr, 1, L0.stream));
        if (p.filter_adaptive_mean) SVB_TRY(launch_adaptive_mean(d, c0->mean_mode, D, L0.Dtmp, 1, L0.stream));
        if (p.filter_median) SVB_TRY(launch_median(d, D, L0.Dtmp, 1, L0.stream));
    }
    SVB_CUDA(cudaMemcpyAsync(D1, L0.Dlr, N * 4, cudaMemcpyDeviceToHost, L0.stream));
    SVB_CUDA(cudaMemcpyAsync(D2, L0.Dlr + N, N * 4, cudaMemcpyDeviceToHost, L0.stream));
    SVB_CUDA(cudaEventRecord(g->ev_end, L0.stream));
    for (int b = n - 1; b >= 0; b--) {
        SVB_CUDA(cudaSetDevice(g->ctx[b]->device));
        SVB_CUDA(cudaStreamSynchronize(g->ctx[b]->lanes[0].stream));
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g->ev_begin, g->ev_end);
    g->stats.gpu_ms = ms;
    g->stats.wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    g->stats.triangles = L0.h_ntri[0] + L0.h_ntri[1];
    return SVB_OK;
}

}  // extern "C"
