// Internal declarations shared by the CUDA translation units of libelas_b200.so.
// Nothing here crosses the C-ABI (include/elas_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/elas_b200.h"

namespace svb {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
extern thread_local long long g_launch_counter;  // kernels launched by this library on this thread

#define SVB_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            svb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
            return SVB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define SVB_TRY(call)              \
    do {                           \
        int _r = (call);           \
        if (_r != SVB_OK) return _r; \
    } while (0)

#define SVB_LAUNCH_CHECK()                                                                          \
    do {                                                                                            \
        svb::g_launch_counter++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess) {                                                                    \
            svb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return SVB_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// ---- guard build (make GUARD=1 -> lib_guard/libelas_b200.so) ----------------------------------------------------------------
// Every load / store whose index the kernels do not test at run time ("loads never need a guard" arguments in k_dense.cu,
// k_support.cu, k_prior.cu, the shared-memory tiles of k_post.cu / k_ccl.cu / k_post_fused.cu) asserts its range in this build and
// traps with file:line; `pytest -m gpu` is run against it once per round (profiles/rNN_guard_pytest.log).  The product build
// compiles the macro away.
#ifdef SVB_GUARD
#define SVB_GUARD_ASSERT(cond)                                                         \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            printf("SVB_GUARD %s:%d: %s violated (block %d,%d,%d thread %d)\n", __FILE__, __LINE__, #cond, (int)blockIdx.x, (int)blockIdx.y, \
                   (int)blockIdx.z, (int)threadIdx.x);                                 \
            __trap();                                                                  \
        }                                                                              \
    } while (0)
#else
#define SVB_GUARD_ASSERT(cond) ((void)0)
#endif

// ---- geometry derived from (params, width, height) ---------------------------------------------
struct Dims {
    int W, H, N;          // image size, N = W*H
    int bpl;              // bytes per line of every device-resident input image: W rounded up to 16 (the reference's own padding,
                          // elas.cpp:37) -- what lets the descriptor kernel fetch its tiles with TMA (row strides must be 16-byte multiples)
    size_t IN;            // bytes per input image on the device: bpl * H
    int sub;              // param.subsampling: disparities only for even (u, v), maps are (W/2) x (H/2) (elas.h:81-83)
    int Dw, Dh, DN;       // disparity map size: W x H, or W/2 x H/2 with subsampling; DN = Dw*Dh
    int step, cw, ch;     // candidate grid (elas.cpp:376-386)
    int gw, gh, gwords;   // disparity grid cells (elas.cpp:88-89) and 32-bit words per cell bitmask
    int maxS;             // capacity of the support list  ((cw-1)*(ch-1) + 6 corner points)
    int maxT;             // capacity of one triangle list (2*maxS is an upper bound for a planar triangulation)
    int plane_radius;     // elas.cpp:832
    int P[8];             // prior table P[|d - d_plane|] for deltas 0..plane_radius (elas.cpp:831)
    size_t desc_pad;      // bytes of zero-filled guard band in front of and behind every descriptor arena (the matchers' unguarded loads)
    int cost_bias;        // added to every matching cost so that SAD + P stays >= 0 in the unsigned packed key: max(16, -min P)
    // optional device counters (svb_set_eval_counting): [0] support-matching hypotheses (64-byte SADs), [1] dense-matching
    // hypotheses (16-byte SADs), counted the way the reference evaluates them; nullptr = the kernels without counting
    unsigned long long *evals;
};

int make_dims(const svb_params &p, int W, int H, Dims *out);

// plane of one triangle for one side, as consumed by the dense matcher
struct __align__(16) PlaneRec {
    float a, b, c;  // plane in the image being matched (t1* for left, t2* for right)
    int valid;      // fabs(a) < 0.7 && fabs(other side's a) < 0.7 (elas.cpp:910)
};

enum StageId {
    ST_DESCRIPTOR = 0,
    ST_SUPPORT_MATCH,
    ST_SUPPORT_FILTER,
    ST_DELAUNAY_DEVICE,  // device half of the Delaunay stage: vertex order (k_order.cu) and the lower merge levels (k_delaunay.cu)
    ST_D2H_SUPPORT,
    ST_H2D_TRIANGLES,
    ST_PLANES,
    ST_GRID,
    ST_RASTER,
    ST_DENSE,
    ST_LR,
    ST_SEGMENTS,
    ST_GAP,
    ST_MEAN,
    ST_MEDIAN,
    ST_REPROJECT,
    ST_POST_FUSED,  // adaptive mean + median + u8 + reprojection in one kernel (k_post_fused.cu); the three stages before it are then empty
    ST_COUNT
};

// ---- kernel launchers (each one is batched: `nimg`/`nf` independent images or frames) ----------
// k_descriptor.cu
// img: nimg images of d.IN bytes each, d.bpl bytes per line
int launch_descriptor(const Dims &d, const uint8_t *img, uint8_t *desc, int nimg, cudaStream_t s);
// *_rows variants restrict a stage to image rows [row0, row1) / lattice rows [vc0, vc1): the row-band split (band_split.cu)
int launch_descriptor_rows(const Dims &d, const uint8_t *img, uint8_t *desc, int nimg, int row0, int row1, cudaStream_t s);  // d.sub: even rows only
// k_order.cu: recursion order of the host Delaunay stage for every (frame, side); h_order [nf][2][maxS], h_ok [nf][2] (mapped host memory)
// d_order / d_ok: optional device copies (same layout) for launch_delaunay_levels
// dup_count [nf][2] / dup_keys [nf][2][delaunay_dup_keys_per_list()]: optional device scratch; with it, lists with duplicate coordinates
// are sorted and de-duplicated on the device exactly like the reference does it, without it they are flagged for the host
size_t delaunay_dup_keys_per_list();
int launch_delaunay_order(const Dims &d, const int32_t *support, const int32_t *nsupport, int32_t *h_order, int32_t *h_ok, int32_t *d_order,
                          int32_t *d_ok, int32_t *dup_count, unsigned long long *dup_keys, int nf, cudaStream_t s);
// k_delaunay.cu: the divide-and-conquer itself on the device, one CTA per (frame, side); lists of up to cap_n (<= 4096) points whose
// order is usable are triangulated into tri1 / tri2 at frame f * (maxT + 8) (h_ntri = count, h_done = 1, mapped host memory),
// the others get h_done = 0 and are left to the host stage
size_t delaunay_levels_smem(int cap_n);
int launch_delaunay_levels(const Dims &d, const int32_t *support, const int32_t *nsupport, const int32_t *order, const int32_t *order_ok,
                           int32_t *tri1, int32_t *tri2, int32_t *h_ntri, int32_t *h_done, int nf, int cap_n, cudaStream_t s);
// k_support.cu
int launch_support_match(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int nf,
                         cudaStream_t s);
int launch_dcan_border(const Dims &d, int16_t *dcan_raw, int nf, cudaStream_t s);
int launch_support_match_rows(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int nf, int vc0,
                              int vc1, cudaStream_t s);
// h_support / h_nsupport: device-accessible (mapped pinned) host copies written by the kernel itself, may be null;
// changed: device scratch of SUPPORT_FILTER_SCRATCH_INTS int32 per frame (per-sweep flags of the multi-CTA path for large lattices)
constexpr int SUPPORT_FILTER_SCRATCH_INTS = 10;
int launch_support_filter(const Dims &d, const svb_params &p, const int16_t *dcan_raw, int16_t *dcan, int32_t *support, int32_t *nsupport,
                          int32_t *h_support, int32_t *h_nsupport, int32_t *changed, int nf, cudaStream_t s);
// k_prior.cu
// tri1 / tri2 hold the frames' triangle lists packed back to back: frame f starts at triangle trioff[f] (both sides)
int launch_planes(const Dims &d, const int32_t *support, const int32_t *tri1, const int32_t *tri2, const int32_t *ntri, const int32_t *trioff,
                  float *planes_ref1, float *planes_ref2, PlaneRec *rec1, PlaneRec *rec2, int nf, int max_tri, cudaStream_t s);
int launch_grid(const Dims &d, const svb_params &p, const int32_t *support, const int32_t *nsupport, uint32_t *tmp, uint32_t *grid1,
                uint32_t *grid2, int nf, int max_support, cudaStream_t s);
int launch_grid_expand(const Dims &d, const svb_params &p, const uint32_t *grid, int32_t *grid_ref, cudaStream_t s);
// The owner map holds, per pixel, the index of the last triangle that covers it (-1: none).  gen > 0: entries are (gen << 24 | index) and
// the map is NOT cleared -- entries of earlier generations are smaller than any of this one (atomicMax keeps the new ones) and read as
// "none" through owner_untag; the caller clears the map before it reuses a generation number.  gen = 0: plain indices, map cleared first.
constexpr int OWNER_GEN_SHIFT = 24, OWNER_GEN_MAX = 127, OWNER_INDEX_MASK = 0xFFFFFF;
#ifdef __CUDACC__
__device__ __forceinline__ int owner_untag(int raw, int gen) { return (raw >> OWNER_GEN_SHIFT) == gen ? (raw & OWNER_INDEX_MASK) : -1; }
#endif
int launch_raster(const Dims &d, const int32_t *support, const int32_t *tri1, const int32_t *tri2, const int32_t *ntri, const int32_t *trioff,
                  int32_t *owner1, int32_t *owner2, int nf, int max_tri, cudaStream_t s, int gen = 0);
int launch_raster_rows(const Dims &d, const int32_t *support, const int32_t *tri1, const int32_t *tri2, const int32_t *ntri, const int32_t *trioff,
                       int32_t *owner1, int32_t *owner2, int nf, int max_tri, int row0, int row1, cudaStream_t s, int gen = 0);
// k_dense.cu
int launch_dense(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, const int32_t *owner1, const int32_t *owner2,
                 const PlaneRec *rec1, const PlaneRec *rec2, const uint32_t *grid1, const uint32_t *grid2, float *D1, float *D2, int nf,
                 cudaStream_t s, int owner_gen = 0);
int launch_dense_rows(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, const int32_t *owner1, const int32_t *owner2,
                      const PlaneRec *rec1, const PlaneRec *rec2, const uint32_t *grid1, const uint32_t *grid2, float *D1, float *D2, int nf,
                      int row0, int row1, cudaStream_t s, int owner_gen = 0);
// k_post.cu
int launch_lr_check(const Dims &d, const svb_params &p, const float *D1in, const float *D2in, float *D1out, float *D2out, int nf,
                    cudaStream_t s);
int launch_lr_check_rows(const Dims &d, const svb_params &p, const float *D1in, const float *D2in, float *D1out, float *D2out, int nf, int row0,
                         int row1, cudaStream_t s);
// scratch: nimg * Dh * ceil(Dw / 32) words (row-pass validity bits for the column pass).  labels / sizes non-null: the row pass also does
// the pruning step of the speckle removal (launch_ccl_label has run, launch_ccl_prune has not) on maps whose invalid pixels are all -10
int launch_gap(const Dims &d, const svb_params &p, float *D, uint32_t *scratch, const int32_t *labels, const int32_t *sizes, int nimg, cudaStream_t s);
constexpr int CCL_KEPT = 0x40000000;  // flag in a non-root pixel's label: its component is large enough inside its tile (pixel indices are < 2^26)
int launch_adaptive_mean(const Dims &d, int mean_mode, float *D, float *tmp, int nimg, cudaStream_t s);
int launch_median(const Dims &d, float *D, float *tmp, int nimg, cudaStream_t s);
// k_ccl.cu
// labels, sizes, roots: one int32 per pixel and image; counts: one per 128 x 16 tile and image (ccl_tiles_per_image)
int ccl_tiles_per_image(const Dims &d);
int ccl_min_size(const Dims &d, const svb_params &p);
int launch_ccl_label(const Dims &d, const svb_params &p, const float *D, int32_t *labels, int32_t *sizes, int32_t *roots, int32_t *counts, int nimg,
                     cudaStream_t s);
int launch_ccl_prune(const Dims &d, const svb_params &p, float *D, const int32_t *labels, const int32_t *sizes, int nimg, cudaStream_t s);
int launch_remove_small_segments(const Dims &d, const svb_params &p, float *D, int32_t *labels, int32_t *sizes, int32_t *roots, int32_t *counts,
                                 int nimg, cudaStream_t s);
// k_reproject.cu
struct Calib;
// k_post_fused.cu: adaptive mean -> median -> final map (+ u8 map + point cloud) in one pass; full-resolution maps only
int launch_post_fused(const Dims &d, const svb_params &p, int mean_mode, const Calib &cal, const float *Din, float *Dout, uint8_t *dmap,
                      double *points, int float_disp, int nimg, cudaStream_t s);
struct Calib {
    double Q[16];  // row-major 4x4 (cv::stereoRectify's Q, stereo_vision.cu:447)
    double XR[9];  // row-major 3x3
    double XT[3];
};
int launch_reproject(const Dims &d, const Calib &c, const float *D, uint8_t *dmap, double *points, int nf, cudaStream_t s);
// the float disparity itself enters Q (invalid pixels as 0): SVB_OUT_POINTS_FLOATDISP outside the fused kernel
int launch_reproject_float(const Dims &d, const Calib &c, const float *D, double *points, int nf, cudaStream_t s);
int launch_reproject_u8(const Calib &c, const uint8_t *dmap, double *points, int W, int H, cudaStream_t s);
// k_convert.cu
// bgra: rows x W pixels, tight; gray: rows of `pitch` bytes (rows may span several frames stored back to back)
int launch_bgra_to_gray(const uint8_t *bgra, uint8_t *gray, int W, int rows, int pitch, cudaStream_t s);

}  // namespace svb
