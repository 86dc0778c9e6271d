/* elas_b200.h -- C-ABI of the B200-native ELAS stereo hot path (libelas_b200.so).
 *
 * Plain pointers and sizes only; no C++ or framework types cross this boundary.
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * repository root).  The reference's own exported C symbols (generatePointCloud / clean /
 * getColor, src/parallel_includes/main/stereo_vision.cu:113-127,574-637) are declared in
 * include/stereo_vision_c.h and live in the same shared library.
 *
 * Error convention: functions returning int return 0 on success and a negative svb_status on
 * failure; svb_last_error() gives the message.  There is NO CPU fallback: without a CUDA device
 * svb_create() fails with SVB_ERR_NO_DEVICE.
 */
#ifndef ELAS_B200_H
#define ELAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* POD mirror of Elas::parameters (src/parallel_includes/elas/elas.h:58-83): same field order,
 * bools widened to int32 so that every field is 4 bytes. */
typedef struct svb_params {
    int32_t disp_min;
    int32_t disp_max;
    float support_threshold;
    int32_t support_texture;
    int32_t candidate_stepsize;
    int32_t incon_window_size;
    int32_t incon_threshold;
    int32_t incon_min_support;
    int32_t add_corners;
    int32_t grid_size;
    float beta;
    float gamma;
    float sigma;
    float sradius;
    int32_t match_texture;
    int32_t lr_threshold;
    float speckle_sim_threshold;
    int32_t speckle_size;
    int32_t ipol_gap_width;
    int32_t filter_median;
    int32_t filter_adaptive_mean;
    int32_t postprocess_only_left;
    int32_t subsampling;
} svb_params;

enum svb_setting { SVB_ROBOTICS = 0, SVB_MIDDLEBURY = 1, SVB_PIPELINE = 2 };

enum svb_status {
    SVB_OK = 0,
    SVB_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the library never computes on the CPU */
    SVB_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    SVB_ERR_ARG = -3,         /* invalid argument */
    SVB_ERR_UNSUPPORTED = -4, /* parameter combination not implemented (e.g. subsampling) */
    SVB_ERR_FEW_SUPPORT = -5  /* < 3 support points: outputs left untouched, like elas.cpp:64-69 */
};

/* adaptive-mean weight form (SURVEY.md finding 3) */
enum svb_mean_mode {
    SVB_MEAN_SERIAL_QUANTISED = 0, /* serial reference: bit-mask "abs" (elas.cpp:1329) -> weights {4,2,0} */
    SVB_MEAN_TRUE_ABS = 1          /* parallel reference: max(-x,x) (src/parallel_includes/elas/elas.cpp:1552) */
};

typedef struct svb_context svb_context;

const char *svb_last_error(void);
const char *svb_version(void);
int svb_device_count(void);

/* Elas::parameters(setting) (src/parallel_includes/elas/elas.h:86-142).  SVB_PIPELINE is the preset
 * generateDisparityMap() builds (src/parallel_includes/main/stereo_vision.cu:315-319):
 * MIDDLEBURY + postprocess_only_left + filter_adaptive_mean. */
int svb_default_params(int setting, svb_params *out);

/* Replaces `ElasGPU elas(param)` (src/parallel_includes/elas/elas_gpu.h:26-33) plus the lazy
 * ElasGPU::memInit (elas_gpu.cu:289-306): all device arenas, pinned staging buffers, streams and
 * the host Delaunay worker pool are created once, sized for `chunk` frames of width x height in
 * flight per lane (chunk = 1: one lane; chunk > 1: 8 lanes, SVB_LANES in the environment overrides; roughly 1.4 GB of device memory per lane for 32 frames of 1242x375).  device < 0 picks the current device. */
svb_context *svb_create(const svb_params *params, int width, int height, int chunk, int device);
void svb_destroy(svb_context *ctx);
int svb_set_mean_mode(svb_context *ctx, int mode);
int svb_set_delaunay_threads(svb_context *ctx, int n_threads);
/* on: every lane issues its GPU work on ONE stream (kernels never overlap one another, so the per-stage CUDA-event
 * times of svb_stats are exact); off (default): one stream per lane. */
int svb_set_single_stream(svb_context *ctx, int on);
/* Measurement aid ("pixel-disparity evaluations"): while on, the two matching kernels also count the
 * hypotheses the reference algorithm evaluates -- support matching: every d of a candidate's range, forward and
 * backward (elas.cpp:330, one hypothesis = four 16-byte SADs); dense matching: grid candidates outside the plane band
 * plus the band, warped column inside the image (elas.cpp:759-793, one 16-byte SAD each).  svb_get_eval_counts returns
 * the totals since the last call and resets them.  Off by default: the timed kernels carry no counting code. */
int svb_set_eval_counting(svb_context *ctx, int on);
int svb_get_eval_counts(svb_context *ctx, uint64_t *support_hypotheses, uint64_t *dense_hypotheses);

/* Elas::process (src/parallel_includes/elas/elas.h:151-160, serial semantics of
 * src/serial_includes/elas/elas.cpp:31-150).  Host buffers; I1/I2 are u8 with `stride` bytes per
 * line; D1/D2 are caller-allocated width*height floats. */
int svb_process(svb_context *ctx, const uint8_t *I1, const uint8_t *I2, int stride, float *D1, float *D2);

/* ---- stage taps of the most recent svb_process() call (parity harness) ---------------------
 * With tap mode on, svb_process() snapshots every intermediate; svb_tap copies one to the host.
 * names: desc1 desc2 dcan_raw dcan support tri1 tri2 planes1 planes2 grid1 grid2 owner1 owner2
 *        D1raw D2raw D1lr D2lr D1seg D1gap D1mean D1med (and the D2* counterparts)
 * Returns the number of bytes written, or a negative svb_status. */
int svb_set_tap_mode(svb_context *ctx, int on);
int64_t svb_tap(svb_context *ctx, const char *name, void *dst, int64_t capacity_bytes);

/* Use a caller-supplied triangle list (n x {c1,c2,c3}) for side 0 (left) / 1 (right) in the next
 * svb_process() calls instead of the host Delaunay stage; n < 0 restores the built-in stage.
 * Replaces the output of Elas::computeDelaunayTriangulation (elas.cpp:442-501) for
 * stage-isolated parity (SURVEY.md finding 7). */
int svb_inject_triangles(svb_context *ctx, int side, const int32_t *tri, int n);

/* ---- stage-isolated entry points (host in / host out, one frame) ---------------------------- */
/* Descriptor::Descriptor (src/common_includes/elas/descriptor.cpp:30-39): out = 16*W*H bytes */
int svb_stage_descriptor(svb_context *ctx, const uint8_t *I, int stride, uint8_t *desc_out);
/* Elas::computeSupportMatches (elas.cpp:373-440): dcan_raw/dcan are ch*cw int16 (may be NULL),
 * support is cap x {u,v,d}; returns the number of support points via *n_out. */
int svb_stage_support(svb_context *ctx, const uint8_t *desc1, const uint8_t *desc2, int16_t *dcan_raw, int16_t *dcan,
                      int32_t *support, int cap, int *n_out);
/* Elas::computeDelaunayTriangulation (elas.cpp:442-501) -- the host stage on its own.  All four Delaunay entry points return
 * SVB_ERR_ARG for a point whose x (u, or u - d for the right image) lies outside [-8192, 16383] or whose v lies outside [0, 8191]:
 * the range the stage's integer predicates are exact for, and more than any supported frame (<= 8192 x 8192, disp_max <= 4095)
 * produces. */
int svb_stage_delaunay(const int32_t *support, int n, int right_image, int32_t *tri, int cap, int *n_tri_out);
/* Elas::computeDisparityPlanes (elas.cpp:503-575): planes = m x {t1a,t1b,t1c,t2a,t2b,t2c} */
/* The host half of the pipeline's stage: `order` = the n support indices in the order the divide-and-conquer meets them
 * (lexicographic (x, y) sort, then the alternating-axis median partition with subsets of <= 3 x-sorted; duplicate-free
 * input only), as csrc/k_order.cu computes it on the device. */
int svb_stage_delaunay_ordered(const int32_t *support, int n, int right_image, const int32_t *order, int32_t *tri, int cap, int *n_tri_out);
/* The device's share of the stage restated on the host (tests; no GPU needed): the recursion levels from the leaves up to depth
 * `host_levels` built level by level, every node independently, in 16-bit records -- what csrc/k_delaunay.cu does with one thread per
 * node -- and the levels above by the host recursion (0: the device's work is the whole triangulation).  n <= 4096. */
int svb_stage_delaunay_levels(const int32_t *support, int n, int right_image, const int32_t *order, int host_levels, int32_t *tri, int cap,
                              int *n_tri_out);
/* The same stage as the pipeline runs it: vertex order (sort + alternating cuts) and the divide-and-conquer on the device
 * (*used_device_order = 2), or -- SVB_DELAUNAY_DEVICE=0 -- the recursion on the host (1); 0 when the device flagged the list
 * (duplicate coordinates, > 4096 points) and the host did it all. */
int svb_stage_delaunay_pipeline(svb_context *ctx, const int32_t *support, int n, int right_image, int32_t *tri, int cap, int *n_tri_out,
                                int *used_device_order);
int svb_stage_planes(svb_context *ctx, const int32_t *support, int n, const int32_t *tri, int m, float *planes);
/* Elas::createGrid (elas.cpp:577-653): grid = gh*gw*(disp_max+2) int32, reference layout */
int svb_stage_grid(svb_context *ctx, const int32_t *support, int n, int right_image, int32_t *grid);
/* Elas::computeDisparity + findMatch (elas.cpp:688-944) */
int svb_stage_disparity(svb_context *ctx, const int32_t *support, int n, const int32_t *tri, int m, const uint8_t *desc1,
                        const uint8_t *desc2, int right_image, float *D);
/* Elas::leftRightConsistencyCheck (elas.cpp:946-1011), in place on both */
int svb_stage_lr_check(svb_context *ctx, float *D1, float *D2);
/* Elas::removeSmallSegments / gapInterpolation / adaptiveMean / median (elas.cpp:1013,1126,1297,1496) */
int svb_stage_remove_small_segments(svb_context *ctx, float *D);
int svb_stage_gap_interpolation(svb_context *ctx, float *D);
int svb_stage_adaptive_mean(svb_context *ctx, float *D);
int svb_stage_median(svb_context *ctx, float *D);
/* generateDisparityMap tail + projectParallel (stereo_vision.cu:324,188-212):
 * dmap = saturate_u8(rint(4*D)); points = XR * (Q*[x y d 1]^T)_{xyz/w} + XT.  dmap_out may be NULL. */
int svb_stage_reproject(svb_context *ctx, const float *D, const double *Q16, const double *XR9, const double *XT3,
                        uint8_t *dmap_out, double *points_out);

/* ---- frame-batch pipeline (BASELINE.json configs[1]: 1242x375 x 1024 frames) ------------------
 * Frames are independent: a batch is cut into chunks that flow through
 * descriptor+support+Delaunay (GPU; the host worker pool only triangulates lists the device hands back) -> planes..reproject (GPU),
 * with the lanes overlapping one another.  Replaces the per-frame loop imageLoop()/generatePointCloud()
 * (stereo_vision.cu:645-697,574-632). */
int svb_set_calibration(svb_context *ctx, const double *Q16, const double *XR9, const double *XT3);
/* copy n frames (tight W*H u8 each) into the device-resident input store */
int svb_batch_upload(svb_context *ctx, const uint8_t *left, const uint8_t *right, int n_frames);
/* the same from BGRA frames (height x width x 4 bytes each, what sv.py passes to generatePointCloud, sv.py:185-188): uploaded in groups
 * and converted on the device like cv::cvtColor(BGRA2GRAY) (stereo_vision.cu:346-347) */
int svb_batch_upload_bgra(svb_context *ctx, const uint8_t *left_bgra, const uint8_t *right_bgra, int n_frames);
/* run the whole path on the resident inputs; results stay resident.  flags: SVB_OUT_* */
/* SVB_OUT_POINTS: the drop-in point cloud (u8 disparity x4 as in generateDisparityMap, stereo_vision.cu:324, clips at 63.75 px);
 * SVB_OUT_POINTS_FLOATDISP (opt-in, instead of SVB_OUT_POINTS): the filtered FLOAT disparity enters Q, no quantisation and no clip --
 * what a 4K / disparity-range-512 caller needs; invalid pixels project like disparity 0. */
enum svb_out_flags { SVB_OUT_DISPARITY = 1, SVB_OUT_POINTS = 2, SVB_OUT_POINTS_FLOATDISP = 4 };
int svb_batch_run(svb_context *ctx, int n_frames, int flags);
/* Per-frame status of the last batch call: nsupport_out[f] = support points of frame f.  A frame with fewer than 3 has no
 * triangulation ("ERROR: Need at least 3 support points!", elas.cpp:64-69): its batch outputs are what generatePointCloud
 * delivers in that case -- disparity 0 everywhere (the driver's maps are zero-initialised, stereo_vision.cu:311-312) and
 * the point cloud of that map -- and it is counted in svb_stats::frames_failed. */
int svb_batch_frame_support(svb_context *ctx, int32_t *nsupport_out, int n_frames);
/* Zero-copy access to the resident results of the last batch call: DEVICE pointers, D1 = n_frames x H x W float,
 * points = n_frames x H x W x 3 double, on CUDA device *device; NULL for an output the call did not produce.  Valid until the next
 * batch call on this context. */
int svb_batch_device_ptrs(svb_context *ctx, float **D1_dev, double **points_dev, int *n_frames, int *device);
int svb_batch_download_disparity(svb_context *ctx, int frame, float *D1_out);
int svb_batch_download_points(svb_context *ctx, int frame, double *points_out);
/* end to end from pinned or pageable HOST buffers: H2D inputs, run, D2H outputs, all inside */
int svb_batch_run_host(svb_context *ctx, const uint8_t *left, const uint8_t *right, int n_frames, int flags, float *D1_out,
                       double *points_out);

/* Body of generatePointCloud() after its one-time init (stereo_vision.cu:596-618): BGRA -> gray (cvtColor), Elas::process
 * with the context's parameters, u8 conversion (convertTo(CV_8UC1, 4.0)) and projectParallel with the calibration set by
 * svb_set_calibration.  left/right: height*width*4 BGRA bytes (host); points_out: width*height*3 doubles (host, pinned
 * for full PCIe rate); dmap_out (u8) and D1_out (f32) may be NULL.  times_ms (may be NULL) receives {disparity part,
 * point-cloud part} in CUDA-event milliseconds (the reference's dmap_t / pc_t).  With < 3 support points the disparity
 * is 0 everywhere, like the reference's untouched zero-initialised map, and SVB_ERR_FEW_SUPPORT is returned. */
int svb_point_cloud_bgra(svb_context *ctx, const uint8_t *left_bgra, const uint8_t *right_bgra, double *points_out, uint8_t *dmap_out,
                         float *D1_out, double *times_ms);
/* cv::resize(src, dst, dsize) with the default INTER_LINEAR on 8-bit BGRA (stereo_vision.cu:599-600,665,676: frames are
 * resized to out_img_size = input size / scale_factor).  Host buffers; OpenCV's fixed-point arithmetic restated exactly. */
int svb_resize_bgra(const uint8_t *src, int src_width, int src_height, uint8_t *dst, int dst_width, int dst_height);
/* the same cv::resize arithmetic for a single-channel u8 image, and publishPointCloud (stereo_vision.cu:245-265) on a u8
 * disparity map of any size: what the driver's extrapolate_point_cloud (-e) option needs (resize the map, then project). */
int svb_resize_gray(const uint8_t *src, int src_width, int src_height, uint8_t *dst, int dst_width, int dst_height);
int svb_reproject_u8(const uint8_t *dmap, int width, int height, const double *Q16, const double *XR9, const double *XT3, double *points_out);
/* cv::cvtColor(BGRA2GRAY) on its own (stereo_vision.cu:346-347) */
int svb_stage_bgra_to_gray(svb_context *ctx, const uint8_t *bgra, uint8_t *gray_out);

/* ---- one frame split into row bands over several GPUs (BASELINE.json configs[3]: 3840x2160, disparity range 512) ------
 * New relative to the reference (single GPU).  Same semantics and bit-identical results as svb_process / Elas::process
 * (src/parallel_includes/elas/elas.h:151-160).  One process drives all devices; descriptor halo rows, the candidate
 * lattice and the L/R-checked band rows move between the GPUs' memories with peer-to-peer copies (NVLink 5 / NVSwitch
 * when peer access is available).  A device may be listed more than once. */
typedef struct svb_band_group svb_band_group;
typedef struct svb_band_stats {
    double gpu_ms;          /* CUDA-event time of the call on device 0 (first upload to last download) */
    double wall_ms;         /* host wall time of the call */
    double delaunay_ms;     /* host Delaunay stage inside the call */
    int64_t p2p_bytes;      /* bytes moved by peer-to-peer copies */
    int64_t p2p_copies;
    int32_t peer_links;     /* directed device pairs with peer access enabled */
    int32_t bands;
    int64_t support_points, triangles;
} svb_band_stats;
svb_band_group *svb_band_create(const svb_params *params, int width, int height, const int *devices, int n_devices);
void svb_band_destroy(svb_band_group *g);
int svb_band_process(svb_band_group *g, const uint8_t *I1, const uint8_t *I2, int stride, float *D1, float *D2);
int svb_band_get_stats(svb_band_group *g, svb_band_stats *out);

/* timing / accounting of the most recent batch or process call */
typedef struct svb_stats {
    double gpu_ms_total;       /* CUDA-event time from first to last kernel of the call */
    double delaunay_ms_total;  /* summed host time inside the Delaunay stage (all workers) */
    double delaunay_ms_wall;   /* wall time the pipeline waited on the host stage */
    int64_t kernel_launches;   /* kernels launched by this library during the call */
    int64_t support_points;    /* summed over frames */
    int64_t triangles;         /* summed over frames, both sides */
    int64_t frames;
    int64_t frames_failed;     /* frames with < 3 support points */
    double stage_ms[24];       /* per-stage CUDA-event time, index = svb_stage_id */
    int64_t delaunay_lists_device; /* triangulations the device made (csrc/k_delaunay.cu), both sides counted */
    int64_t delaunay_lists_host;   /* triangulations the host stage made (duplicate coordinates, > 4096 points, injected lists, SVB_DELAUNAY_DEVICE=0) */
} svb_stats;
int svb_get_stats(svb_context *ctx, svb_stats *out);
const char *svb_stage_name(int stage_id);
int svb_set_stage_timing(svb_context *ctx, int on);

/* ---- calibration without OpenCV (host only) -------------------------------------------------------------
 * svb_calib_load_yaml replaces the cv::FileStorage reads of externalInit()/main()
 * (src/parallel_includes/main/stereo_vision.cu:536-545,824-832): K1 K2 D1 D2 R T XR XT from an OpenCV-YAML file
 * (XR/XT default to identity/zero when absent).
 * svb_stereo_rectify replaces findRectificationMap() (stereo_vision.cu:368-447): K1/K2's first two rows are divided
 * by scale_factor, then cv::stereoRectify(K1, D1, K2, D2, calib size, R, T, ..., CALIB_ZERO_DISPARITY, alpha, new
 * size) is restated; the reference passes alpha = 0.  Outputs are row-major R1[9] R2[9] P1[12] P2[12] Q[16]; any may be
 * NULL. */
typedef struct svb_calibration {
    double K1[9], K2[9];
    double D1[14], D2[14];
    int32_t n_d1, n_d2;
    double R[9], T[3];
    double XR[9], XT[3];
} svb_calibration;
int svb_calib_load_yaml(const char *path, svb_calibration *out);
int svb_stereo_rectify(const svb_calibration *cal, int calib_width, int calib_height, int new_width, int new_height, double scale_factor,
                       double alpha, double *R1, double *R2, double *P1, double *P2, double *Q);

/* Image file input of the sequence driver without OpenCV / libpng (host only): 8/16-bit non-interlaced PNG (decoded
 * with zlib) or binary PGM, chosen by the file extension.  Replaces cv::imread (stereo_vision.cu:661-662) and loadPGM
 * (src/common_includes/image.h:134-161).  Colour PNGs are returned as BGRA (4 channels), gray images as 1 channel.
 * With out == NULL only the dimensions are returned. */
int svb_image_read(const char *path, uint8_t *out, int64_t capacity, int *width, int *height, int *channels);

/* Deterministic synthetic rectified stereo pair of known disparity (bench / parity INPUT generator, host only;
 * SURVEY.md 8d).  left/right: W*H u8 each.  slanted = 0: bands of disparity 8/24/48; 1: d = 10 + 0.03 u. */
int svb_synth_pair(int frame_index, int width, int height, int slanted, uint8_t *left, uint8_t *right);

/* pinned host memory helpers for callers that want the e2e path at full PCIe rate */
void *svb_host_alloc(size_t bytes);
void svb_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
