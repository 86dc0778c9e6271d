// TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product.
//
// Row 19 oracle: the reference's OWN `projectParallel` kernel.  oracle/build_ref.sh cuts the kernel's text out
// of /root/reference/src/parallel_includes/main/stereo_vision.cu (lines 188-212, from `__global__ void projectParallel` to the
// closing brace) into oracle/_ref/project_parallel_extract.cuh at build time -- nothing of it is committed -- and compiles this
// harness around it with the reference Makefile's nvcc flags (-O2 -std=c++17 -w; FMA contraction on, nvcc's default) for sm_100a.
// The harness restates only publishPointCloud's launch (stereo_vision.cu:245-265: cudaMemcpy in, 32x32 blocks,
// grid (cols/32 + 1, rows/32 + 1), cudaMemcpy out).
#include <cuda_runtime.h>

typedef unsigned char uchar;

#include "project_parallel_extract.cuh"

extern "C" const char *ref_project_flags(void) { return ORACLE_REF_PROJECT_FLAGS; }

// dmap: rows*cols u8 (host); points: rows*cols double3 (host); XT[3], XR[9], Q[16] row-major doubles (host).
// Returns 0, or the cudaError_t that stopped it.
extern "C" int ref_project_parallel(const uchar *dmap, double *points, int rows, int cols, const double *XT, const double *XR, const double *Q) {
    const size_t n = (size_t)rows * cols;
    uchar *d_dmap = nullptr;
    double3 *d_points = nullptr;
    double *d_XT = nullptr, *d_XR = nullptr, *d_Q = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_dmap, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_points, n * sizeof(double3));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_XT, 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_XR, 9 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_Q, 16 * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_XT, XT, 3 * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_XR, XR, 9 * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_Q, Q, 16 * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_dmap, dmap, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const dim3 blockSize(32, 32, 1);
        const dim3 gridSize((cols / blockSize.x) + 1, (rows / blockSize.y) + 1, 1);
        projectParallel<<<gridSize, blockSize, 0>>>(d_dmap, d_points, rows, cols, d_XT, d_XR, d_Q);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(points, d_points, n * sizeof(double3), cudaMemcpyDeviceToHost);
    cudaFree(d_dmap);
    cudaFree(d_points);
    cudaFree(d_XT);
    cudaFree(d_XR);
    cudaFree(d_Q);
    return (int)e;
}
