// Sobel 3x3 + 16-byte descriptor, fused.
//
// Replaces Descriptor::Descriptor -> filter::sobel3x3 -> Descriptor::createDescriptor
// (src/common_includes/elas/descriptor.cpp:30-39,98-124; src/common_includes/elas/filter.cpp:380-424).
//
// Semantics reproduced:
//   tv(r,c) = I(r-1,c) + 2 I(r,c) + I(r+1,c)          th(r,c) = I(r-1,c) - I(r+1,c)
//   du(r,c) = sat_u8(((tv(r,c-1) - tv(r,c+1)) >> 2) + 128)
//   dv(r,c) = sat_u8(((th(r,c-1) + 2 th(r,c) + th(r,c+1)) >> 2) + 128)
//   desc(v,u), 3<=v<H-3, 3<=u<W-3 = { du(v-2,u), du(v-1,u-2), du(v-1,u), du(v-1,u+2), du(v,u-1), du(v,u), du(v,u),
//                                     du(v,u+1), du(v+1,u-2), du(v+1,u), du(v+1,u+2), du(v+2,u),
//                                     dv(v-1,u), dv(v,u-1), dv(v,u+1), dv(v+1,u) }
//   every other descriptor is 0 (the reference leaves them unwritten; SURVEY.md finding 5 defines them as 0).
// The reference runs the row convolutions over the flat bpl*H array, but every du/dv value that a descriptor
// reads lies in rows 1..H-2, cols 1..W-2, where the flat and the 2-D formulations coincide.
//
// Layout: one CTA produces a 64x16 tile of descriptors.  The (64+6)x(16+6) input halo tile is staged in shared
// memory once, du/dv tiles are derived in shared memory, and every thread emits whole 16-byte descriptors with
// one 128-bit store; a warp writes 512 contiguous bytes.  HBM traffic per image: W*H read + 16*W*H written.
#include "svb_internal.h"

namespace svb {

namespace {

constexpr int TW = 64;
constexpr int TH = 16;
constexpr int NT = 256;

__device__ __forceinline__ int sat_u8(int x) { return min(max(x, 0), 255); }

__global__ void __launch_bounds__(NT) k_descriptor(const uint8_t *__restrict__ img, uint8_t *__restrict__ desc, int W, int H, int row0, int row1, int half) {
    __shared__ uint8_t sI[TH + 6][TW + 6 + 2];
    __shared__ uint8_t sDu[TH + 4][TW + 4];
    __shared__ uint8_t sDv[TH + 2][TW + 2 + 2];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    const int y0 = row0 + blockIdx.y * TH;  // only rows row0 .. row1-1 are produced (row-band split; the whole image otherwise)
    const size_t N = (size_t)W * H;
    const uint8_t *I = img + (size_t)blockIdx.z * N;
    uint4 *out = reinterpret_cast<uint4 *>(desc + (size_t)blockIdx.z * N * 16);

    // input tile: rows y0-3 .. y0+TH+2, cols x0-3 .. x0+TW+2 (zero outside the image: never reaches a valid descriptor)
    for (int i = tid; i < (TH + 6) * (TW + 6); i += NT) {
        int r = i / (TW + 6), c = i - r * (TW + 6);
        int y = y0 - 3 + r, x = x0 - 3 + c;
        uint8_t val = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) val = __ldg(I + (size_t)y * W + x);
        sI[r][c] = val;
    }
    __syncthreads();

    // du tile: rows y0-2 .. y0+TH+1, cols x0-2 .. x0+TW+1   (image (y,x) lives at sI[y-y0+3][x-x0+3])
    for (int i = tid; i < (TH + 4) * (TW + 4); i += NT) {
        int r = i / (TW + 4), c = i - r * (TW + 4);
        int rr = r + 1, cc = c + 1;
        int tl = sI[rr - 1][cc - 1] + 2 * sI[rr][cc - 1] + sI[rr + 1][cc - 1];
        int tr = sI[rr - 1][cc + 1] + 2 * sI[rr][cc + 1] + sI[rr + 1][cc + 1];
        sDu[r][c] = (uint8_t)sat_u8(((tl - tr) >> 2) + 128);
    }
    // dv tile: rows y0-1 .. y0+TH, cols x0-1 .. x0+TW
    for (int i = tid; i < (TH + 2) * (TW + 2); i += NT) {
        int r = i / (TW + 2), c = i - r * (TW + 2);
        int rr = r + 2, cc = c + 2;
        int h0 = sI[rr - 1][cc - 1] - sI[rr + 1][cc - 1];
        int h1 = sI[rr - 1][cc] - sI[rr + 1][cc];
        int h2 = sI[rr - 1][cc + 1] - sI[rr + 1][cc + 1];
        sDv[r][c] = (uint8_t)sat_u8(((h0 + 2 * h1 + h2) >> 2) + 128);
    }
    __syncthreads();

    for (int i = tid; i < TH * TW; i += NT) {
        int ty = i / TW, tx = i - ty * TW;
        int u = x0 + tx, v = y0 + ty;
        if (u >= W || v >= row1) continue;
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        // half resolution (descriptor.cpp:50-93): only rows v = 4, 6, ... < H-3 are produced, every other row is zero
        if (u >= 3 && u < W - 3 && v >= 3 && v < H - 3 && (!half || (v >= 4 && (v & 1) == 0))) {
            const int a = ty + 2, b = tx + 2;  // du(v,u) = sDu[a][b]
            const int e = ty + 1, f = tx + 1;  // dv(v,u) = sDv[e][f]
            uint32_t c0 = sDu[a][b];
            q.x = (uint32_t)sDu[a - 2][b] | ((uint32_t)sDu[a - 1][b - 2] << 8) | ((uint32_t)sDu[a - 1][b] << 16) |
                  ((uint32_t)sDu[a - 1][b + 2] << 24);
            q.y = (uint32_t)sDu[a][b - 1] | (c0 << 8) | (c0 << 16) | ((uint32_t)sDu[a][b + 1] << 24);
            q.z = (uint32_t)sDu[a + 1][b - 2] | ((uint32_t)sDu[a + 1][b] << 8) | ((uint32_t)sDu[a + 1][b + 2] << 16) |
                  ((uint32_t)sDu[a + 2][b] << 24);
            q.w = (uint32_t)sDv[e - 1][f] | ((uint32_t)sDv[e][f - 1] << 8) | ((uint32_t)sDv[e][f + 1] << 16) |
                  ((uint32_t)sDv[e + 1][f] << 24);
        }
        out[(size_t)v * W + u] = q;
    }
}

}  // namespace

int launch_descriptor(const Dims &d, const uint8_t *img, uint8_t *desc, int nimg, cudaStream_t s) {
    return launch_descriptor_rows(d, img, desc, nimg, 0, d.H, s);
}

int launch_descriptor_rows(const Dims &d, const uint8_t *img, uint8_t *desc, int nimg, int row0, int row1, cudaStream_t s) {
    if (nimg <= 0 || row1 <= row0) return SVB_OK;
    dim3 grid((d.W + TW - 1) / TW, (row1 - row0 + TH - 1) / TH, nimg);
    k_descriptor<<<grid, NT, 0, s>>>(img, desc, d.W, d.H, row0, row1, d.sub);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
