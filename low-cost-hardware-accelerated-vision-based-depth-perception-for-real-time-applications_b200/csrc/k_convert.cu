// BGRA -> 8-bit gray, the input side of the drop-in boundary.
//
// Replaces cv::cvtColor(..., COLOR_BGRA2GRAY) in imgCallback_video() (src/parallel_includes/main/stereo_vision.cu:346-347).
// OpenCV's 8-bit path is fixed point: gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15 (BT.601 weights scaled by 2^15,
// round to nearest); pinned against python cv2 4.13 on random pixels (tests/golden/calib_golden.json: gray_probe).
// Memory-bound: 4 B read + 1 B written per pixel, into the pitched device image (Dims::bpl bytes per line).
#include <math.h>

#include <vector>

#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ uint32_t gray_of(uint32_t bgra) {
    const uint32_t b = bgra & 0xFFu, g = (bgra >> 8) & 0xFFu, r = (bgra >> 16) & 0xFFu;
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

// grid: (ceil(W / 256), rows).  One pixel per thread: a 4-byte coalesced load, a 1-byte store into the pitched gray image.
__global__ void __launch_bounds__(256) k_bgra_to_gray(const uint8_t *__restrict__ bgra, uint8_t *__restrict__ gray, int W, int pitch) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const size_t row = blockIdx.y;
    gray[row * pitch + x] = (uint8_t)gray_of(__ldg(reinterpret_cast<const uint32_t *>(bgra) + row * W + x));
}

}  // namespace

int launch_bgra_to_gray(const uint8_t *bgra, uint8_t *gray, int W, int rows, int pitch, cudaStream_t s) {
    if (W <= 0 || rows <= 0) return SVB_OK;
    for (int r0 = 0; r0 < rows; r0 += 65535) {  // grid.y limit
        const int nr = rows - r0 < 65535 ? rows - r0 : 65535;
        k_bgra_to_gray<<<dim3((W + 255) / 256, nr), 256, 0, s>>>(bgra + (size_t)r0 * W * 4, gray + (size_t)r0 * pitch, W, pitch);
        SVB_LAUNCH_CHECK();
    }
    return SVB_OK;
}

}  // namespace svb

// ---- cv::resize(INTER_LINEAR) for 8-bit BGRA ---------------------------------------------------------------------
// Replaces the resize() calls of the sequence driver (src/parallel_includes/main/stereo_vision.cu:599-600,665,676: input
// frames are shrunk / enlarged by scale_factor before matching).  OpenCV's 8-bit bilinear path is fixed point and is
// restated exactly (pinned against python cv2 4.13 for scale factors 0.5 .. 3.0, tests/test_gpu_dropin.py):
//   fx = float((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx; columns are clamped WITH their weight
//   (sx < 0 -> sx = 0, fx = 0; sx >= sw-1 -> sx = sw-1, fx = 0), rows only clamp the index and keep the weight;
//   weights = round(w * 2048) as int16; horizontal pass in int32, vertical pass
//   dst = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
namespace svb {
namespace {

struct ResizeTab {
    int idx;
    short w0, w1;
};

__global__ void __launch_bounds__(256) k_resize_linear_bgra(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const ResizeTab *__restrict__ xt,
                                                            const ResizeTab *__restrict__ yt, int sw, int sh, int dw, int dh) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y;
    if (dx >= dw) return;
    const ResizeTab tx = xt[dx], ty = yt[dy];
    const int x0 = tx.idx, x1 = min(tx.idx + 1, sw - 1);
    const int y0 = min(max(ty.idx, 0), sh - 1), y1 = min(max(ty.idx + 1, 0), sh - 1);
    const uint32_t *r0 = reinterpret_cast<const uint32_t *>(src) + (size_t)y0 * sw;
    const uint32_t *r1 = reinterpret_cast<const uint32_t *>(src) + (size_t)y1 * sw;
    const uint32_t p00 = r0[x0], p01 = r0[x1], p10 = r1[x0], p11 = r1[x1];
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int sft = 8 * c;
        const int S0 = (int)((p00 >> sft) & 0xFF) * tx.w0 + (int)((p01 >> sft) & 0xFF) * tx.w1;
        const int S1 = (int)((p10 >> sft) & 0xFF) * tx.w0 + (int)((p11 >> sft) & 0xFF) * tx.w1;
        int v = (((ty.w0 * (S0 >> 4)) >> 16) + ((ty.w1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v = min(max(v, 0), 255);
        out |= (uint32_t)v << sft;
    }
    reinterpret_cast<uint32_t *>(dst)[(size_t)dy * dw + dx] = out;
}

// The same arithmetic for a single-channel image (the u8 disparity map of publishPointCloud, stereo_vision.cu:249).
__global__ void __launch_bounds__(256) k_resize_linear_gray(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, const ResizeTab *__restrict__ xt,
                                                            const ResizeTab *__restrict__ yt, int sw, int sh, int dw, int dh) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y;
    if (dx >= dw) return;
    const ResizeTab tx = xt[dx], ty = yt[dy];
    const int x0 = tx.idx, x1 = min(tx.idx + 1, sw - 1);
    const int y0 = min(max(ty.idx, 0), sh - 1), y1 = min(max(ty.idx + 1, 0), sh - 1);
    const uint8_t *r0 = src + (size_t)y0 * sw, *r1 = src + (size_t)y1 * sw;
    const int S0 = (int)r0[x0] * tx.w0 + (int)r0[x1] * tx.w1;
    const int S1 = (int)r1[x0] * tx.w0 + (int)r1[x1] * tx.w1;
    const int v = (((ty.w0 * (S0 >> 4)) >> 16) + ((ty.w1 * (S1 >> 4)) >> 16) + 2) >> 2;
    dst[(size_t)dy * dw + dx] = (uint8_t)min(max(v, 0), 255);
}

void make_table(std::vector<ResizeTab> &t, int dn, int sn, bool clamp_weight) {
    const double scale = 1.0 / ((double)dn / sn);
    t.resize(dn);
    for (int d = 0; d < dn; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (clamp_weight) {
            if (s < 0) {
                f = 0;
                s = 0;
            }
            if (s >= sn - 1) {
                f = 0;
                s = sn - 1;
            }
        }
        t[d].idx = s;
        t[d].w0 = (short)lrintf((1.f - f) * 2048.f);
        t[d].w1 = (short)lrintf(f * 2048.f);
    }
}

}  // namespace
}  // namespace svb

static int resize_u8(const char *who, int channels, const uint8_t *src, int src_width, int src_height, uint8_t *dst, int dst_width, int dst_height) {
    using namespace svb;
    if (!src || !dst || src_width < 1 || src_height < 1 || dst_width < 1 || dst_height < 1) return SVB_ERR_ARG;
    std::vector<ResizeTab> xt, yt;
    make_table(xt, dst_width, src_width, true);
    make_table(yt, dst_height, src_height, false);
    uint8_t *d_src = nullptr, *d_dst = nullptr;
    ResizeTab *d_xt = nullptr, *d_yt = nullptr;
    const size_t sb = (size_t)src_width * src_height * channels, db = (size_t)dst_width * dst_height * channels;
    int rc = SVB_OK;
    cudaError_t e = cudaMalloc(&d_src, sb);
    if (e == cudaSuccess) e = cudaMalloc(&d_dst, db);
    if (e == cudaSuccess) e = cudaMalloc(&d_xt, xt.size() * sizeof(ResizeTab));
    if (e == cudaSuccess) e = cudaMalloc(&d_yt, yt.size() * sizeof(ResizeTab));
    if (e == cudaSuccess) e = cudaMemcpy(d_src, src, sb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_xt, xt.data(), xt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_yt, yt.data(), yt.size() * sizeof(ResizeTab), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dim3 grid((dst_width + 255) / 256, dst_height);
        if (channels == 4)
            k_resize_linear_bgra<<<grid, 256>>>(d_src, d_dst, d_xt, d_yt, src_width, src_height, dst_width, dst_height);
        else
            k_resize_linear_gray<<<grid, 256>>>(d_src, d_dst, d_xt, d_yt, src_width, src_height, dst_width, dst_height);
        g_launch_counter++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(dst, d_dst, db, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        set_error("%s: %s", who, cudaGetErrorString(e));
        rc = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SVB_ERR_NO_DEVICE : SVB_ERR_CUDA;
    }
    cudaFree(d_src);
    cudaFree(d_dst);
    cudaFree(d_xt);
    cudaFree(d_yt);
    return rc;
}

extern "C" int svb_resize_bgra(const uint8_t *src, int src_width, int src_height, uint8_t *dst, int dst_width, int dst_height) {
    return resize_u8("svb_resize_bgra", 4, src, src_width, src_height, dst, dst_width, dst_height);
}

extern "C" int svb_resize_gray(const uint8_t *src, int src_width, int src_height, uint8_t *dst, int dst_width, int dst_height) {
    return resize_u8("svb_resize_gray", 1, src, src_width, src_height, dst, dst_width, dst_height);
}
