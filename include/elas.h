// elas.h -- the C++ entry points of the reference's ELAS classes, backed by the sm_100a library.
//
// Mirrors src/parallel_includes/elas/elas.h:53-160 (class Elas, enum setting, struct parameters with the same field
// order, types and presets, Elas(parameters), process(I1, I2, D1, D2, dims)) and
// src/parallel_includes/elas/elas_gpu.h:26-33 (class ElasGPU : public Elas, constructible from parameters).
// Callers such as generateDisparityMap() (src/parallel_includes/main/stereo_vision.cu:315-321) compile unchanged
// against this header and link libelas_b200.so.  The reference exposes its stage methods as public/virtual members
// for its own GPU subclass; those are implementation details and are not part of this boundary -- the stage-level
// entry points of the new path are the svb_stage_* functions of elas_b200.h.
//
// There is no CPU implementation behind this class: process() fails loudly (message on stderr, D1/D2 untouched)
// when no CUDA device is available.
#ifndef ELAS_B200_ELAS_H
#define ELAS_B200_ELAS_H

#include <stdint.h>

struct svb_context;

class Elas {
   public:
    enum setting { ROBOTICS, MIDDLEBURY };

    struct parameters {
        int32_t disp_min;
        int32_t disp_max;
        float support_threshold;
        int32_t support_texture;
        int32_t candidate_stepsize;
        int32_t incon_window_size;
        int32_t incon_threshold;
        int32_t incon_min_support;
        bool add_corners;
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
        bool filter_median;
        bool filter_adaptive_mean;
        bool postprocess_only_left;
        bool subsampling;

        parameters(setting s = ROBOTICS);
    };

    Elas(parameters param);
    virtual ~Elas();
    Elas(const Elas &) = delete;
    Elas &operator=(const Elas &) = delete;

    // dims[0] = width, dims[1] = height, dims[2] = bytes per line of I1 / I2; D1 / D2: width x height floats
    void process(uint8_t *I1, uint8_t *I2, float *D1, float *D2, const int32_t *dims);

    parameters param;  // the reference keeps it readable by subclasses; changes take effect at the next process()

   private:
    svb_context *ctx_ = nullptr;
    int ctx_w_ = 0, ctx_h_ = 0;
    unsigned char ctx_param_[128];  // the svb_params the context was created with
};

class ElasGPU : public Elas {
   public:
    ElasGPU(parameters param) : Elas(param) {}
    ~ElasGPU() override {}
};

#endif
