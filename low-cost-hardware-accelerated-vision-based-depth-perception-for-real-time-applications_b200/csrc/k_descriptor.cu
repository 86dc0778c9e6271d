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
// Layout.  One CTA produces a 64 x 32 tile of descriptors; the kernel is bound by shared-memory / L1 wavefronts, not
// by HBM (ncu on the first version: 77 % L1 wavefronts, 25 % DRAM), so everything is arranged to need few of them:
//   phase 1  the (64+12) x (32+6) input tile goes to shared memory (zero outside the image);
//   phase 2  a thread owns one COLUMN of the du/dv tiles and slides down it with the 3 x 3 input window in registers:
//            3 byte loads per (du, dv) pair instead of 12;
//   phase 3  a thread owns one column and 4 consecutive ROWS of descriptors.  Per du / dv row it reads the 5 (3)
//            neighbouring bytes with two 32-bit shared loads (one wavefront each: four neighbouring lanes share a word)
//            and a funnel shift, the 8 du rows and 6 dv rows are shared by its 4 descriptors, and the descriptors are
//            assembled with byte permutes.  Neighbouring lanes own neighbouring columns, so every 128-bit store
//            instruction of a warp covers 512 contiguous bytes.
// HBM traffic per image: W*H read + 16*W*H written.
#include "svb_internal.h"

namespace svb {

namespace {

constexpr int TW = 64;            // tile width  (descriptors)
constexpr int TH = 32;            // tile height (descriptors)
constexpr int NT = 256;           // threads: 64 columns x 4 row groups in phase 3
constexpr int SW = TW + 12;       // shared row stride in bytes (multiple of 4); local column = x - (x0 - 4)
constexpr int IN_ROWS = TH + 6;   // input rows y0-3 .. y0+TH+2
constexpr int DU_ROWS = TH + 4;   // du rows    y0-2 .. y0+TH+1
constexpr int DV_ROWS = TH + 2;   // dv rows    y0-1 .. y0+TH
constexpr int DCOLS = TW + 4;     // du / dv columns computed: local 2 .. TW+5  (x0-2 .. x0+TW+1)
constexpr int SEGS = 3;           // phase 2: the DU_ROWS rows of a column are split into 3 segments

__device__ __forceinline__ int sat_u8(int x) { return min(max(x, 0), 255); }

__global__ void __launch_bounds__(NT) k_descriptor(const uint8_t *__restrict__ img, uint8_t *__restrict__ desc, int W, int H, int row0, int row1,
                                                   int half) {
    __shared__ __align__(16) uint8_t sI[IN_ROWS][SW];
    __shared__ __align__(16) uint8_t sDu[DU_ROWS][SW];
    __shared__ __align__(16) uint8_t sDv[DV_ROWS][SW];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    const int y0 = row0 + blockIdx.y * TH;  // only rows row0 .. row1-1 are produced (row-band split; the whole image otherwise)
    const size_t N = (size_t)W * H;
    const uint8_t *I = img + (size_t)blockIdx.z * N;
    uint4 *out = reinterpret_cast<uint4 *>(desc + (size_t)blockIdx.z * N * 16);

    // ---- phase 1: input tile, rows y0-3 .. y0+TH+2, local columns 0 .. SW-1 (image column x0-4+c) -------------------
    // all loads of a thread are issued before the first store, so their latencies overlap
    constexpr int P1_ITERS = (IN_ROWS * SW + NT - 1) / NT;
    uint8_t vals[P1_ITERS];
#pragma unroll
    for (int k = 0; k < P1_ITERS; k++) {
        const int i = tid + k * NT;
        const int r = i / SW, c = i - r * SW;
        const int y = y0 - 3 + r, x = x0 - 4 + c;
        vals[k] = 0;
        if (i < IN_ROWS * SW && y >= 0 && y < H && x >= 0 && x < W) vals[k] = __ldg(I + (unsigned)(y * W) + x);
    }
#pragma unroll
    for (int k = 0; k < P1_ITERS; k++) {
        const int i = tid + k * NT;
        if (i < IN_ROWS * SW) (&sI[0][0])[i] = vals[k];
    }
    __syncthreads();

    // ---- phase 2: du / dv columns, sliding 3 x 3 window ------------------------------------------------------------
    if (tid < DCOLS * SEGS) {
        const int seg = tid / DCOLS, c = 2 + (tid - seg * DCOLS);  // local column of this thread
        const int rows_per = (DU_ROWS + SEGS - 1) / SEGS;
        const int ja = seg * rows_per, jb = min(ja + rows_per, DU_ROWS);  // du rows [ja, jb): image row y0-2+j, input rows j..j+2
        int a0 = sI[ja][c - 1], a1 = sI[ja][c], a2 = sI[ja][c + 1];              // input row j
        int b0 = sI[ja + 1][c - 1], b1 = sI[ja + 1][c], b2 = sI[ja + 1][c + 1];  // input row j+1
        for (int j = ja; j < jb; j++) {
            const int c0 = sI[j + 2][c - 1], c1 = sI[j + 2][c], c2 = sI[j + 2][c + 1];  // input row j+2
            const int tl = a0 + 2 * b0 + c0, tr = a2 + 2 * b2 + c2;
            sDu[j][c] = (uint8_t)sat_u8(((tl - tr) >> 2) + 128);
            if (j >= 1 && j <= DV_ROWS) {  // dv rows y0-1 .. y0+TH  <->  du row index j = 1 .. TH+2
                const int h0 = a0 - c0, h1 = a1 - c1, h2 = a2 - c2;
                sDv[j - 1][c] = (uint8_t)sat_u8(((h0 + 2 * h1 + h2) >> 2) + 128);
            }
            a0 = b0;
            a1 = b1;
            a2 = b2;
            b0 = c0;
            b1 = c1;
            b2 = c2;
        }
    }
    __syncthreads();

    // ---- phase 3: thread = one column x 4 rows -------------------------------------------------------------------
    const int tx = tid & (TW - 1);  // column in the tile
    const int u = x0 + tx;
    const int lc = tx + 4;          // local column of u
    if (u >= W) return;
    // du bytes u-2 .. u+2 of du row j: words (lc-2)>>2 and the next one, shifted by ((lc-2)&3) bytes
    const int wdu = (lc - 2) >> 2, sdu = ((lc - 2) & 3) * 8;
    const int wdv = (lc - 1) >> 2, sdv = ((lc - 1) & 3) * 8;
    const bool col_ok = u >= 3 && u < W - 3;
    for (int tg = tid / TW; tg < TH / 4; tg += NT / TW) {  // row groups of 4
        const int rbase = tg * 4;                            // first output row of the group inside the tile
        if (y0 + rbase >= row1 || y0 + rbase >= H) break;
        uint32_t lo[8], hi[8], dw[6];
#pragma unroll
        for (int j = 0; j < 8; j++) {  // du rows rbase+j (image row y0 + rbase - 2 + j)
            const uint32_t *row = reinterpret_cast<const uint32_t *>(sDu[rbase + j]);
            const uint32_t w0 = row[wdu], w1 = row[wdu + 1];
            lo[j] = __funnelshift_rc(w0, w1, sdu);      // du(u-2), du(u-1), du(u), du(u+1)
            hi[j] = __funnelshift_rc(w0, w1, sdu + 8);  // du(u-1), du(u), du(u+1), du(u+2)
        }
#pragma unroll
        for (int j = 0; j < 6; j++) {  // dv rows rbase+j (image row y0 + rbase - 1 + j)
            const uint32_t *row = reinterpret_cast<const uint32_t *>(sDv[rbase + j]);
            dw[j] = __funnelshift_rc(row[wdv], row[wdv + 1], sdv);  // dv(u-1), dv(u), dv(u+1), .
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int v = y0 + rbase + k;
            if (v >= row1 || v >= H) break;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            // half resolution (descriptor.cpp:50-93): only rows v = 4, 6, ... < H-3 are produced, every other row is zero
            if (col_ok && v >= 3 && v < H - 3 && (!half || (v >= 4 && (v & 1) == 0))) {
                // du rows: k = v-2, k+1 = v-1, k+2 = v, k+3 = v+1, k+4 = v+2;   dv rows: k = v-1, k+1 = v, k+2 = v+1
                // q.x = du(v-2,u) | du(v-1,u-2) | du(v-1,u) | du(v-1,u+2)
                q.x = __byte_perm(__byte_perm(lo[k], lo[k + 1], 0x0642), hi[k + 1], 0x7210);
                // q.y = du(v,u-1) | du(v,u) | du(v,u) | du(v,u+1)
                q.y = __byte_perm(lo[k + 2], 0u, 0x3221);
                // q.z = du(v+1,u-2) | du(v+1,u) | du(v+1,u+2) | du(v+2,u)
                q.z = __byte_perm(__byte_perm(lo[k + 3], hi[k + 3], 0x0720), lo[k + 4], 0x6210);
                // q.w = dv(v-1,u) | dv(v,u-1) | dv(v,u+1) | dv(v+1,u)
                q.w = __byte_perm(__byte_perm(dw[k], dw[k + 1], 0x0641), dw[k + 2], 0x5210);
            }
            out[(unsigned)(v * W + u)] = q;
        }
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
