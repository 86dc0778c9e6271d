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
//   phase 1  the (64+32) x (32+6) input tile goes to shared memory (zero outside the image): ONE TMA tile load per CTA
//            (cp.async.bulk.tensor.3d over a {W, H, images} u8 tensor map with Dims::bpl bytes per line; coordinates outside the
//            image -- negative ones included -- are zero-filled by the hardware, which IS the border semantics; an elected thread
//            arms an mbarrier with the tile's byte count, the CTA waits on its phase bit).  The first version's per-byte __ldg loop
//            with four range tests per element cost more ALU-pipe time than everything else in the kernel; it remains as the
//            fallback for partial-frame launches (row bands) and half-resolution descriptors;
//   phase 2  a thread owns one COLUMN of the du/dv tiles and slides down it with the 3 x 3 input window in registers:
//            3 byte loads per (du, dv) pair instead of 12;
//   phase 3  a thread owns one column and 4 consecutive ROWS of descriptors.  Per du / dv row it reads the 5 (3)
//            neighbouring bytes with two 32-bit shared loads (one wavefront each: four neighbouring lanes share a word)
//            and a funnel shift, the 8 du rows and 6 dv rows are shared by its 4 descriptors, and the descriptors are
//            assembled with byte permutes.  Neighbouring lanes own neighbouring columns, so every 128-bit store
//            instruction of a warp covers 512 contiguous bytes.
// HBM traffic per image: W*H read + 16*W*H written.
#include <cuda.h>
#include <string.h>

#include "svb_internal.h"

namespace svb {

namespace {

constexpr int TW = 64;            // tile width  (descriptors)
constexpr int TH = 32;            // tile height (descriptors)
constexpr int NT = 256;           // threads: 64 columns x 4 row groups in phase 3
constexpr int XO = 16;            // the tile starts XO columns left of x0: a TMA box must start at a 16-byte multiple of the innermost
                                  // coordinate (an unaligned start traps as an illegal instruction), and 4 columns of halo are needed
constexpr int SW = TW + 2 * XO;   // shared row stride in bytes = TMA box width (a multiple of 16); local column = x - (x0 - XO)
constexpr int IN_ROWS = TH + 6;   // input rows y0-3 .. y0+TH+2
constexpr int DU_ROWS = TH + 4;   // du rows    y0-2 .. y0+TH+1
constexpr int DV_ROWS = TH + 2;   // dv rows    y0-1 .. y0+TH
constexpr int DCOLS = TW + 4;     // du / dv columns computed: local XO-2 .. XO+TW+1  (x0-2 .. x0+TW+1)
constexpr int SEGS = 3;           // phase 2: the DU_ROWS rows of a column are split into 3 segments

__device__ __forceinline__ int sat_u8(int x) { return min(max(x, 0), 255); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// USE_TMA: whole-frame launches (row0 = 0); tmap describes the images of this launch as a {W, H, nimg} u8 tensor, box {SW, IN_ROWS, 1}
template <bool USE_TMA>
__global__ void __launch_bounds__(NT) k_descriptor(const __grid_constant__ CUtensorMap tmap, const uint8_t *__restrict__ img,
                                                   uint8_t *__restrict__ desc, int W, int H, int pitch, int row0, int row1, int half) {
    __shared__ __align__(128) uint8_t sI[IN_ROWS][SW];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(16) uint8_t sDu[DU_ROWS][SW];
    __shared__ __align__(16) uint8_t sDv[DV_ROWS][SW];

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TW;
    const int y0 = row0 + blockIdx.y * TH;  // only rows row0 .. row1-1 are produced (row-band split; the whole image otherwise)
    const size_t N = (size_t)W * H;
    uint4 *out = reinterpret_cast<uint4 *>(desc + (size_t)blockIdx.z * N * 16);

    // ---- phase 1: input tile, rows y0-3 .. y0+TH+2, local columns 0 .. SW-1 (image column x0-XO+c) -------------------
    if (USE_TMA) {
        const uint32_t bar = smem_u32(&s_bar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the initialised barrier visible to the async proxy
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(IN_ROWS * SW)) : "memory");
            // (.L2::cache_hint with policy 0 = none: the form CUTLASS issues)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
                    smem_u32(&sI[0][0])),
                "l"(&tmap), "r"(x0 - XO), "r"(y0 - 3), "r"((int)blockIdx.z), "r"(bar), "l"(0ull)
                : "memory");
        }
        // every thread waits for phase 0 of the barrier to complete (the TMA's bytes have landed)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_TILE:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
            "@p bra DONE_TILE;\n"
            "bra WAIT_TILE;\n"
            "DONE_TILE:\n"
            "}\n" ::"r"(bar)
            : "memory");
    } else {
        const uint8_t *I = img + (size_t)blockIdx.z * ((size_t)pitch * H);
        // all loads of a thread are issued before the first store, so their latencies overlap
        constexpr int P1_ITERS = (IN_ROWS * SW + NT - 1) / NT;
        uint8_t vals[P1_ITERS];
#pragma unroll
        for (int k = 0; k < P1_ITERS; k++) {
            const int i = tid + k * NT;
            const int r = i / SW, c = i - r * SW;
            const int y = y0 - 3 + r, x = x0 - XO + c;
            vals[k] = 0;
            if (i < IN_ROWS * SW && y >= 0 && y < H && x >= 0 && x < W) vals[k] = __ldg(I + (size_t)y * pitch + x);
        }
#pragma unroll
        for (int k = 0; k < P1_ITERS; k++) {
            const int i = tid + k * NT;
            if (i < IN_ROWS * SW) (&sI[0][0])[i] = vals[k];
        }
        __syncthreads();
    }

    // ---- phase 2: du / dv columns, sliding 3 x 3 window ------------------------------------------------------------
    if (tid < DCOLS * SEGS) {
        const int seg = tid / DCOLS, c = XO - 2 + (tid - seg * DCOLS);  // local column of this thread
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
    const int lc = tx + XO;         // local column of u
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

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime (no link against libcuda); nullptr if the driver does not provide it
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

}  // namespace

int launch_descriptor(const Dims &d, const uint8_t *img, uint8_t *desc, int nimg, cudaStream_t s) {
    if (nimg <= 0) return SVB_OK;
    EncodeTiledFn enc = d.sub ? nullptr : encode_tiled();
    if (!enc || ((uintptr_t)img & 15) != 0) return launch_descriptor_rows(d, img, desc, nimg, 0, d.H, s);
    // the images of this launch as a {W, H, nimg} u8 tensor with bpl bytes per line; a CTA fetches a {SW, IN_ROWS, 1} box
    CUtensorMap tmap;
    const cuuint64_t dims[3] = {(cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)nimg};
    const cuuint64_t strides[2] = {(cuuint64_t)d.bpl, (cuuint64_t)d.IN};  // bytes, dimensions 1 and 2; both multiples of 16
    const cuuint32_t box[3] = {(cuuint32_t)SW, (cuuint32_t)IN_ROWS, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(img), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%d x %d x %d, pitch %d) failed: %d", d.W, d.H, nimg, d.bpl, (int)r);
        return SVB_ERR_CUDA;
    }
    dim3 grid((d.W + TW - 1) / TW, (d.H + TH - 1) / TH, nimg);
    k_descriptor<true><<<grid, NT, 0, s>>>(tmap, img, desc, d.W, d.H, d.bpl, 0, d.H, 0);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

int launch_descriptor_rows(const Dims &d, const uint8_t *img, uint8_t *desc, int nimg, int row0, int row1, cudaStream_t s) {
    if (nimg <= 0 || row1 <= row0) return SVB_OK;
    dim3 grid((d.W + TW - 1) / TW, (row1 - row0 + TH - 1) / TH, nimg);
    CUtensorMap unused;
    memset(&unused, 0, sizeof(unused));
    k_descriptor<false><<<grid, NT, 0, s>>>(unused, img, desc, d.W, d.H, d.bpl, row0, row1, d.sub);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
