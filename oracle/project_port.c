/* TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
 *
 * Row 19 oracle on the CPU: a C restatement of the reference's `projectParallel` kernel
 * (src/parallel_includes/main/stereo_vision.cu:188-212) AS THE REFERENCE'S BUILD COMPUTES IT.  The reference Makefile
 * compiles that kernel with nvcc's default -fmad=true, so its two sums of products are contracted; which products are
 * fused is read off the SASS of oracle/_ref/libproject_ref.so (the reference kernel itself, compiled by
 * oracle/build_ref.sh with the reference's flags):
 *
 *   pos[j]   = fma(d, Q[4j+2], fma(x, Q[4j+0], y * Q[4j+1])) + Q[4j+3]          DMUL, DFMA, DFMA, DADD
 *   X, Y, Z  = pos[0..2] / pos[3]                                                 IEEE division (div.rn.f64)
 *   point[j] = fma(XR[3j+2], Z, fma(XR[3j+0], X, XR[3j+1] * Y)) + XT[j]          DMUL, DFMA, DFMA, DADD
 *
 * fma() here is C99's correctly rounded fused multiply-add (one rounding, like DFMA); the file is compiled with
 * -ffp-contract=off so that nothing else is fused.  Pinned: tests/test_gpu_parity.py::test_reproject_against_the_reference_kernel
 * runs the reference kernel on the GPU box and requires this port (and the product) to equal it bit for bit.
 *
 * d is passed as double per pixel: the u8 map of the drop-in path (stereo_vision.cu:324) or, for the product's own
 * SVB_OUT_POINTS_FLOATDISP extension, the float disparity itself.
 */
#include <math.h>
#include <stddef.h>

/* d: rows*cols doubles; points: rows*cols*3 doubles; XT[3], XR[9], Q[16] row-major */
void port_project_parallel(const double *d, double *points, int rows, int cols, const double *XT, const double *XR, const double *Q) {
    for (int y = 0; y < rows; y++) {
        for (int x = 0; x < cols; x++) {
            const size_t p = (size_t)y * cols + x;
            const double fx = (double)x, fy = (double)y, fd = d[p];
            double pos[4];
            for (int j = 0; j < 4; j++) pos[j] = fma(fd, Q[4 * j + 2], fma(fx, Q[4 * j + 0], fy * Q[4 * j + 1])) + Q[4 * j + 3];
            const double X = pos[0] / pos[3], Y = pos[1] / pos[3], Z = pos[2] / pos[3];
            for (int j = 0; j < 3; j++) points[3 * p + j] = fma(XR[3 * j + 2], Z, fma(XR[3 * j + 0], X, XR[3 * j + 1] * Y)) + XT[j];
        }
    }
}
