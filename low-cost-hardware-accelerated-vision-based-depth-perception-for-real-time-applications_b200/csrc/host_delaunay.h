// Host stage: Delaunay triangulation of the support points (the one inherently sequential stage of the path).
//
// Replaces Elas::computeDelaunayTriangulation -> triangulate("zQB")
// (src/serial_includes/elas/elas.cpp:442-501, src/common_includes/elas/triangle.cpp:8116).
#pragma once
#include <stdint.h>
#include <vector>

namespace svb {

// Reusable scratch memory so that the per-frame stage performs no heap allocation in steady state.
struct DelaunayScratch {
    std::vector<int32_t> storage;  // coordinates (with a sentinel in front), triangle records, sort arrays
    // > 1: a large list (>= 2048 vertices) may use that many threads of its own: the subtrees of one recursion depth write disjoint,
    // pre-determined record ranges (delaunay_mesh.h), so they are built concurrently and the levels above are merged afterwards.
    // What one 4K frame needs (two lists of 14 700 points and nothing else to overlap with); batches parallelise over lists instead.
    int par_threads = 1;
};

// support: n x {u,v,d}.  Left image (right_image = 0) triangulates (u,v), right image (u-d,v).
// tri_out receives up to `cap` triangles as {c1,c2,c3} = indices into `support`, in the exact order and corner
// rotation the reference emits.  Returns the number of triangles (which may exceed cap; only cap are stored).
int delaunay_support(const int32_t *support, int n, int right_image, int32_t *tri_out, int cap, DelaunayScratch &scratch);

// Same, starting from the recursion order computed on the device (k_order.cu): order[i] = index into `support` of the
// i-th vertex after the lexicographic sort and the alternating-axis partition of a DUPLICATE-FREE point set.
// Returns -1 if `order` is not usable (the caller then runs delaunay_support).
// m = number of entries of `order` (n, or fewer when the device has already removed duplicate coordinates the way the reference does).
int delaunay_support_ordered(const int32_t *support, int n, int right_image, const int32_t *order, int m, int32_t *tri_out, int cap,
                             DelaunayScratch &scratch);

// The device's share of the stage, restated on the host (tests, no GPU needed): the levels of the recursion tree from the leaves up
// to depth `host_levels` are built level by level in 16-bit records -- exactly what k_delaunay.cu does with one thread per node --
// and the host recursion finishes the levels above (host_levels = 0: nothing is left for it but the list).  n <= 4096.
int delaunay_support_levels(const int32_t *support, int n, int right_image, const int32_t *order, int host_levels, int32_t *tri_out, int cap,
                            DelaunayScratch &scratch);

// Same on explicit integer coordinates.
int delaunay_xy(const int32_t *x, const int32_t *y, int n, int32_t *tri_out, int cap, DelaunayScratch &scratch);

}  // namespace svb
