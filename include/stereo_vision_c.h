/* stereo_vision_c.h -- the C symbols the reference's `make shared_library` exports
 * (src/parallel_includes/main/stereo_vision.cu:113-127 clean, :574-632 generatePointCloud, :634-636 getColor;
 * Makefile:135-140), re-implemented on the sm_100a path.  stereo_vision/sv.py binds exactly these through ctypes
 * (sv.py:164-192), so the library built here can be handed to it through `so_lib_path=`.
 *
 * Same names, argument order and types as the reference.  double3 / uchar4 are CUDA's vector types in the
 * reference's signature; their layouts ({double x,y,z}, 24 bytes; {unsigned char x,y,z,w}) are spelled out here so
 * that a C compiler can include this file.
 */
#ifndef STEREO_VISION_C_H
#define STEREO_VISION_C_H

#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sv_double3 {
    double x, y, z;
} sv_double3;
typedef struct sv_uchar4 {
    unsigned char x, y, z, w;
} sv_uchar4;

/* stereo_vision.cu:574-632.  left/right: BGRA bytes, height x width x 4 (sv.py:185-188).  The first call latches
 * width, height, scale and the calibration file (function-local static, :591); the calibration YAML is read without
 * OpenCV and cv::stereoRectify is restated (csrc/calib.cpp).  Returns a library-owned buffer of width*height points,
 * overwritten by the next call, valid until clean().  Never throws; problems are logged to stderr and the buffer is
 * still returned (the reference has no error convention).
 * objectTracking / graphics / display select subsystems outside the hot path (YOLO, OpenGL viewer, imshow): they are
 * accepted and ignored with a one-time notice.  removeSky / subsampling are not passed by sv.py (14 of 16 arguments,
 * sv.py:180): they are read only when the environment variable SVB_TRUST_TAIL_ARGS=1 is set. */
sv_double3 *generatePointCloud(unsigned char *left, unsigned char *right, char *CAMERA_CALIBRATION_YAML, int width, int height,
                               bool kittiCalibration, bool objectTracking, bool graphics, bool display, int scale, int pc_extrapolation,
                               const char *YOLO_CFG, const char *YOLO_WEIGHTS, const char *YOLO_CLASSES, bool removeSky, bool subsampling);

/* stereo_vision.cu:113-127: releases everything, prints "Program exitted successfully!" and calls exit(0) -- yes,
 * it ends the process, exactly like the reference (sv.py:191-192 calls it from __del__).  Set SVB_CLEAN_NO_EXIT=1 to
 * release without exiting. */
void clean(void);

/* stereo_vision.cu:634-636: the BGRA colours of the most recent left image (library-owned copy). */
sv_uchar4 *getColor(void);

#ifdef __cplusplus
}
#endif
#endif
