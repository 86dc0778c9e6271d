// Device share of the Delaunay stage: the divide-and-conquer itself, level by level.
//
// Replaces, for duplicate-free lists of up to 4096 support points, the recursion of the reference's triangulator
// (triangulate("zQB") -> divconqrecurse / mergehulls, src/common_includes/elas/triangle.cpp:5362-5815, called from
// Elas::computeDelaunayTriangulation, src/serial_includes/elas/elas.cpp:442-501).  The decision rules are the ones of
// delaunay_mesh.h -- the SAME source the host stage compiles -- so the triangle list, its order and the corner rotation are the
// reference's (tests: level-synchronous restatement on the CPU against the oracle, and this kernel against the oracle on the GPU).
//
// Why this parallelises although the merge is sequential: records are never freed, a subtree of c vertices creates exactly 2c - 2
// of them in depth-first order, so every node of the recursion tree knows in advance where its records go.  All nodes of one depth
// are independent; the kernel walks the tree bottom-up with one barrier per level, one thread per node (one WARP per node once a
// level has no more nodes than warps, so that the long merges near the root do not diverge against each other).  The mesh lives in
// shared memory: 8 bytes per vertex plus 16-byte records (16-bit handles and vertex ids).  What is left is latency -- the root merge
// is one thread chasing handles -- which the frame-batch pipeline hides behind the other lanes' kernels: one CTA per (frame, side),
// 64 per chunk.  The finished list (live records in creation order, ghosts dropped: triangle.cpp:7449-7500) is written to the
// frame's slot of the device triangle arena; nothing but two counters per list crosses PCIe.
//
// Lists the vertex-order kernel flagged (more than 4096 points, coordinates outside its key range) or that exceed this launch's
// shared-memory capacity are left to the host stage (host_delaunay.cpp).  Duplicate coordinates are resolved by k_order.cu.
#include "delaunay_mesh.h"
#include "svb_internal.h"

namespace svb {

namespace {

constexpr int DD_THREADS = 256;

// grid: (nf, 2); blockIdx.y = image side.  Dynamic shared memory: (cap_n + 1) points, then 2 cap_n records of 8 x uint16.
__global__ void __launch_bounds__(DD_THREADS) k_delaunay_levels(const int32_t *__restrict__ support_all, const int32_t *__restrict__ nsupport_all,
                                                               const int32_t *__restrict__ order_all, const int32_t *__restrict__ order_ok_all,
                                                               int32_t *__restrict__ tri1_all, int32_t *__restrict__ tri2_all,
                                                               int32_t *__restrict__ h_ntri, int32_t *__restrict__ h_done, int maxS, int tri_stride,
                                                               int cap_n) {
    extern __shared__ __align__(16) unsigned char dd_smem[];
    __shared__ int s_warp[DD_THREADS / 32];
    __shared__ int s_base;
    Pt *P = reinterpret_cast<Pt *>(dd_smem) + 1;  // P[-1]: the NULL vertex (its coordinates are read, never used)
    uint16_t *R = reinterpret_cast<uint16_t *>(dd_smem + sizeof(Pt) * (size_t)(cap_n + 1));
    const int f = blockIdx.x, side = blockIdx.y, tid = threadIdx.x;
    // vertices that take part: the whole list, or the survivors of the reference's duplicate removal (k_order.cu); 0 = list not usable
    const int n = order_ok_all[2 * f + side];
    if (n < 3 || n > cap_n || n > maxS || n > nsupport_all[f]) {  // uniform: the host stage takes this list
        if (tid == 0) h_done[2 * f + side] = 0;
        return;
    }
    const int32_t *support = support_all + (size_t)f * maxS * 3;
    const int32_t *order = order_all + ((size_t)f * 2 + side) * maxS;
    for (int i = tid; i < n; i += DD_THREADS) {
        const int32_t *sp = support + 3 * order[i];
        P[i] = Pt{side ? sp[0] - sp[2] : sp[0], sp[1]};  // elas.cpp:451-461
    }
    if (tid < 8) R[tid] = tid < 4 ? (uint16_t)0 : (uint16_t)0xFFFF;  // record 0: "outer space"
    if (tid == 8) P[-1] = Pt{0, 0};
    __syncthreads();

    MeshT<uint16_t> mesh;
    mesh.P = P;
    mesh.R = R;
    mesh.ntri = 0;
    for (int depth = delaunay_max_depth(n); depth >= 0; depth--) {
        const int nodes = 1 << depth;
        if (nodes <= DD_THREADS / 32) {
            if ((tid & 31) == 0 && (tid >> 5) < nodes) {
                const DelaunayNode nd = delaunay_node_at(n, depth, tid >> 5);
                if (nd.exists) mesh.build_node(nd);
            }
        } else {
            for (int k = tid; k < nodes; k += DD_THREADS) {
                const DelaunayNode nd = delaunay_node_at(n, depth, k);
                if (nd.exists) mesh.build_node(nd);
            }
        }
        __syncthreads();
    }

    // the list: live records in creation order, corners (org, dest, apex), as indices into the support list
    int32_t *tri = (side ? tri2_all : tri1_all) + (size_t)f * tri_stride * 3;
    const int last = 2 * n - 2;  // records 1 .. last
    const int lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int start = 1; start <= last; start += DD_THREADS) {
        const int t = start + tid;
        int a = -1, b = -1, c = -1;
        if (t <= last) {
            const uint16_t *r = R + 8 * t + 4;
            a = (int)(int16_t)r[1];
            b = (int)(int16_t)r[2];
            c = (int)(int16_t)r[0];
        }
        const bool keep = (a | b | c) >= 0;  // bounding records carry a NULL vertex (removeghosts, triangle.cpp:5817-5859)
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < wid; w++) off += s_warp[w];
        if (keep) {
            const int pos = off + __popc(bal & ((1u << lane) - 1u));
            tri[3 * pos + 0] = order[a];
            tri[3 * pos + 1] = order[b];
            tri[3 * pos + 2] = order[c];
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < DD_THREADS / 32; w++) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    if (tid == 0) {
        h_ntri[2 * f + side] = s_base;
        h_done[2 * f + side] = 1;
    }
}

}  // namespace

size_t delaunay_levels_smem(int cap_n) { return sizeof(Pt) * (size_t)(cap_n + 1) + (size_t)16 * 2 * cap_n; }

// order / order_ok: DEVICE copies of what k_delaunay_order produced; h_ntri / h_done: mapped host memory, [2 * nf]
int launch_delaunay_levels(const Dims &d, const int32_t *support, const int32_t *nsupport, const int32_t *order, const int32_t *order_ok,
                           int32_t *tri1, int32_t *tri2, int32_t *h_ntri, int32_t *h_done, int nf, int cap_n, cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    if (cap_n > 4096) cap_n = 4096;  // 16-bit handles
    const size_t smem = delaunay_levels_smem(cap_n);
    static int configured[64] = {};  // per device: the largest dynamic shared-memory size opted in to so far
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && configured[dev] < (int)smem) {
        cudaError_t e = cudaFuncSetAttribute(k_delaunay_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(k_delaunay_levels, %zu): %s", smem, cudaGetErrorString(e));
            return SVB_ERR_CUDA;
        }
        configured[dev] = (int)smem;
    }
    k_delaunay_levels<<<dim3(nf, 2), DD_THREADS, smem, s>>>(support, nsupport, order, order_ok, tri1, tri2, h_ntri, h_done, d.maxS, d.maxT + 8, cap_n);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
