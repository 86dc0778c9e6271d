// The reference's own entry points on top of the svb_* C-ABI:
//   C symbols   generatePointCloud / clean / getColor   (src/parallel_includes/main/stereo_vision.cu:113-127,574-637)
//   C++ classes Elas / ElasGPU                           (src/parallel_includes/elas/elas.h:53-160, elas_gpu.h:26-33)
// Host C++ only; every computation happens in the CUDA library.  No OpenCV, popt or GL is needed.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include "../../include/elas.h"
#include "../../include/elas_b200.h"
#include "../../include/stereo_vision_c.h"

// ------------------------------------------------------------------------------------------------ Elas / ElasGPU
Elas::parameters::parameters(setting s) {
    // src/parallel_includes/elas/elas.h:86-142 through the one table the library keeps (svb_default_params)
    svb_params p;
    svb_default_params(s == ROBOTICS ? SVB_ROBOTICS : SVB_MIDDLEBURY, &p);
    disp_min = p.disp_min;
    disp_max = p.disp_max;
    support_threshold = p.support_threshold;
    support_texture = p.support_texture;
    candidate_stepsize = p.candidate_stepsize;
    incon_window_size = p.incon_window_size;
    incon_threshold = p.incon_threshold;
    incon_min_support = p.incon_min_support;
    add_corners = p.add_corners != 0;
    grid_size = p.grid_size;
    beta = p.beta;
    gamma = p.gamma;
    sigma = p.sigma;
    sradius = p.sradius;
    match_texture = p.match_texture;
    lr_threshold = p.lr_threshold;
    speckle_sim_threshold = p.speckle_sim_threshold;
    speckle_size = p.speckle_size;
    ipol_gap_width = p.ipol_gap_width;
    filter_median = p.filter_median != 0;
    filter_adaptive_mean = p.filter_adaptive_mean != 0;
    postprocess_only_left = p.postprocess_only_left != 0;
    subsampling = p.subsampling != 0;
}

static svb_params to_svb(const Elas::parameters &q) {
    svb_params p;
    memset(&p, 0, sizeof(p));
    p.disp_min = q.disp_min;
    p.disp_max = q.disp_max;
    p.support_threshold = q.support_threshold;
    p.support_texture = q.support_texture;
    p.candidate_stepsize = q.candidate_stepsize;
    p.incon_window_size = q.incon_window_size;
    p.incon_threshold = q.incon_threshold;
    p.incon_min_support = q.incon_min_support;
    p.add_corners = q.add_corners ? 1 : 0;
    p.grid_size = q.grid_size;
    p.beta = q.beta;
    p.gamma = q.gamma;
    p.sigma = q.sigma;
    p.sradius = q.sradius;
    p.match_texture = q.match_texture;
    p.lr_threshold = q.lr_threshold;
    p.speckle_sim_threshold = q.speckle_sim_threshold;
    p.speckle_size = q.speckle_size;
    p.ipol_gap_width = q.ipol_gap_width;
    p.filter_median = q.filter_median ? 1 : 0;
    p.filter_adaptive_mean = q.filter_adaptive_mean ? 1 : 0;
    p.postprocess_only_left = q.postprocess_only_left ? 1 : 0;
    p.subsampling = q.subsampling ? 1 : 0;
    return p;
}

static_assert(sizeof(svb_params) <= 128, "Elas::ctx_param_ too small");

Elas::Elas(parameters p) : param(p) { memset(ctx_param_, 0, sizeof(ctx_param_)); }

Elas::~Elas() {
    if (ctx_) svb_destroy(ctx_);
}

void Elas::process(uint8_t *I1, uint8_t *I2, float *D1, float *D2, const int32_t *dims) {
    const int width = dims[0], height = dims[1], bpl = dims[2];
    const svb_params p = to_svb(param);
    if (!ctx_ || ctx_w_ != width || ctx_h_ != height || memcmp(&p, ctx_param_, sizeof(p)) != 0) {
        if (ctx_) svb_destroy(ctx_);
        ctx_ = svb_create(&p, width, height, 1, -1);
        if (!ctx_) {
            fprintf(stderr, "Elas::process: cannot create the CUDA context: %s\n", svb_last_error());
            return;
        }
        ctx_w_ = width;
        ctx_h_ = height;
        memcpy(ctx_param_, &p, sizeof(p));
    }
    const int rc = svb_process(ctx_, I1, I2, bpl, D1, D2);
    if (rc == SVB_ERR_FEW_SUPPORT)
        printf("ERROR: Need at least 3 support points!\n");  // src/serial_includes/elas/elas.cpp:64-69, D1/D2 untouched
    else if (rc != SVB_OK)
        fprintf(stderr, "Elas::process failed: %s\n", svb_last_error());
}

// ------------------------------------------------------------------------------------------------ C symbols
namespace {

struct DropIn {
    bool initialised = false;
    bool failed = false;
    int width = 0, height = 0;
    svb_context *ctx = nullptr;
    double *points = nullptr;        // pinned, width*height*3
    unsigned char *colors = nullptr;  // width*height*4
    double t_t = 1, dmap_t = 0, pc_t = 0;
} g;

bool env_flag(const char *name) {
    const char *v = getenv(name);
    return v && v[0] && v[0] != '0';
}

// externalInit() (stereo_vision.cu:506-572) without OpenCV / YOLO / GL
void external_init(int width, int height, bool graphics, bool display, bool trackObjects, float scale, const char *yaml, bool subsampling) {
    g.initialised = true;
    g.width = width;
    g.height = height;
    if (trackObjects)
        fprintf(stderr, "generatePointCloud: object tracking (YOLO + Bayesian tracker) is outside this library's scope; ignored\n");
    else
        printf("\n** Object tracking disabled\n");
    if (graphics) fprintf(stderr, "generatePointCloud: the OpenGL viewer is outside this library's scope; ignored\n");
    if (display) fprintf(stderr, "generatePointCloud: imshow windows are outside this library's scope; ignored\n");
    printf("Using CAMERA_CALIBRATION_YAML : %s\n", yaml ? yaml : "(null)");
    const size_t n = (size_t)width * height;
    g.points = (double *)svb_host_alloc(n * 3 * sizeof(double));
    if (!g.points) g.points = (double *)calloc(n * 3, sizeof(double));
    else memset(g.points, 0, n * 3 * sizeof(double));
    g.colors = (unsigned char *)calloc(n, 4);
    svb_calibration cal;
    double Q[16];
    if (!yaml || svb_calib_load_yaml(yaml, &cal) != SVB_OK) {
        fprintf(stderr, "generatePointCloud: %s\n", svb_last_error());
        g.failed = true;
        return;
    }
    // findRectificationMap(calib_file, out_img_size) with calib_img_size = out_img_size = (width, height), stereo_vision.cu:521-535
    if (svb_stereo_rectify(&cal, width, height, width, height, scale, 0.0, nullptr, nullptr, nullptr, nullptr, Q) != SVB_OK) {
        fprintf(stderr, "generatePointCloud: %s\n", svb_last_error());
        g.failed = true;
        return;
    }
    // generateDisparityMap()'s preset (stereo_vision.cu:315-319)
    svb_params p;
    svb_default_params(SVB_PIPELINE, &p);
    // `static int res = printf(..., param.subsampling = subsample)` latches the flag of the first frame (stereo_vision.cu:316-317).
    // The half-size map then occupies the first quarter of the full-size float buffer that is converted and projected
    // (leftdpf is width x height, :312-324): reproduced as is.
    p.subsampling = subsampling ? 1 : 0;
    printf("Post Process only left = %d, Subsampling = %d\n", p.postprocess_only_left, p.subsampling);
    g.ctx = svb_create(&p, width, height, 1, -1);
    if (!g.ctx) {
        fprintf(stderr, "generatePointCloud: %s\n", svb_last_error());
        g.failed = true;
        return;
    }
    svb_set_calibration(g.ctx, Q, cal.XR, cal.XT);
    printf("CUDA Init done\n");
    printf("\n** 3D plotting disabled\n");
}

}  // namespace

extern "C" {

sv_double3 *generatePointCloud(unsigned char *left, unsigned char *right, char *CAMERA_CALIBRATION_YAML, int width, int height,
                               bool kittiCalibration, bool objectTracking, bool graphics, bool display, int scale, int pc_extrapolation,
                               const char *YOLO_CFG, const char *YOLO_WEIGHTS, const char *YOLO_CLASSES, bool removeSky, bool subsampling) {
    (void)kittiCalibration;
    (void)pc_extrapolation;  // the reference passes its global (= 1) to externalInit, not this argument (stereo_vision.cu:591)
    (void)YOLO_CFG;
    (void)YOLO_WEIGHTS;
    (void)YOLO_CLASSES;
    // sv.py passes 14 of the 16 arguments (sv.py:180): the last two are whatever the registers held
    if (!env_flag("SVB_TRUST_TAIL_ARGS")) {
        removeSky = false;
        subsampling = false;
    }
    (void)removeSky;  // dmapOLD.copyTo(dmapOLD, sky_mask) copies the map onto itself: no effect in the reference either
    if (!g.initialised) external_init(width, height, graphics, display, objectTracking, (float)scale, CAMERA_CALIBRATION_YAML, subsampling);
    const auto t0 = std::chrono::steady_clock::now();
    if (!g.failed && left && right) {
        if (width != g.width || height != g.height) {
            // the reference resizes to the latched size (stereo_vision.cu:599-600); this path keeps the latched size only
            fprintf(stderr, "generatePointCloud: %dx%d differs from the size latched at the first call (%dx%d); call ignored\n", width, height,
                    g.width, g.height);
        } else {
            memcpy(g.colors, left, (size_t)width * height * 4);  // grapher->setColorsArray(left image), stereo_vision.cu:603
            double times[2] = {0, 0};
            const int rc = svb_point_cloud_bgra(g.ctx, left, right, g.points, nullptr, nullptr, times);
            if (rc == SVB_ERR_FEW_SUPPORT)
                printf("ERROR: Need at least 3 support points!\n");
            else if (rc != SVB_OK)
                fprintf(stderr, "generatePointCloud: %s\n", svb_last_error());
            g.dmap_t = times[0] * 1e-3;
            g.pc_t = times[1] * 1e-3;
        }
    }
    g.t_t = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    // stereo_vision.cu:630 (same line format; scripts parse it)
    printf("(FPS=%f) (%d, %d) (t_t=%f, dmap_t=%f, pc_t=%f)\n", 1 / g.t_t, g.height, g.width, g.t_t, g.dmap_t, g.pc_t);
    return reinterpret_cast<sv_double3 *>(g.points);
}

sv_uchar4 *getColor(void) { return reinterpret_cast<sv_uchar4 *>(g.colors); }

void clean(void) {
    if (g.ctx) svb_destroy(g.ctx);
    g.ctx = nullptr;
    if (g.points) svb_host_free(g.points);
    g.points = nullptr;
    free(g.colors);
    g.colors = nullptr;
    g.initialised = false;
    g.failed = false;
    printf("\n\nProgram exitted successfully!\n\n");
    fflush(stdout);
    if (!env_flag("SVB_CLEAN_NO_EXIT")) exit(0);  // stereo_vision.cu:125
}

}  // extern "C"
