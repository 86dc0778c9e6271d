// BGRA -> 8-bit gray, the input side of the drop-in boundary.
//
// Replaces cv::cvtColor(..., COLOR_BGRA2GRAY) in imgCallback_video() (src/parallel_includes/main/stereo_vision.cu:346-347).
// OpenCV's 8-bit path is fixed point: gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15 (BT.601 weights scaled by 2^15,
// round to nearest); pinned against python cv2 4.13 on random pixels (tests/golden/calib_golden.json: gray_probe).
// Memory-bound: 4 B read + 1 B written per pixel; a thread converts 4 pixels (one 16-byte load, one 4-byte store).
#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ uint32_t gray_of(uint32_t bgra) {
    const uint32_t b = bgra & 0xFFu, g = (bgra >> 8) & 0xFFu, r = (bgra >> 16) & 0xFFu;
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) k_bgra_to_gray(const uint8_t *__restrict__ bgra, uint8_t *__restrict__ gray, int n) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
    const int p = q * 4;
    if (p + 3 < n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(bgra) + q);
        const uint32_t out = gray_of(v.x) | (gray_of(v.y) << 8) | (gray_of(v.z) << 16) | (gray_of(v.w) << 24);
        reinterpret_cast<uint32_t *>(gray)[q] = out;
    } else {
        for (int i = p; i < n; i++) gray[i] = (uint8_t)gray_of(reinterpret_cast<const uint32_t *>(bgra)[i]);
    }
}

}  // namespace

int launch_bgra_to_gray(const uint8_t *bgra, uint8_t *gray, int n, cudaStream_t s) {
    if (n <= 0) return SVB_OK;
    const int groups = (n + 3) / 4;
    k_bgra_to_gray<<<(groups + 255) / 256, 256, 0, s>>>(bgra, gray, n);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
