// Dense matching: per pixel MAP disparity over (grid candidates outside the plane band) then (the plane band
// with the Gaussian prior), integer SAD over 16-byte descriptors.
//
// Replaces Elas::findMatch + updatePosteriorMinimum (src/serial_includes/elas/elas.cpp:655-802) as driven by
// Elas::computeDisparity (elas.cpp:804-944); the triangle -> pixel assignment comes from the owner map
// (k_prior.cu).  Evaluation order and the strict "<" (first evaluated wins ties) are the reference's.
#include "svb_internal.h"

namespace svb {

namespace {

__device__ __forceinline__ uint32_t sad16(const uint4 &a, const uint4 &b) {
    return __vsadu4(a.x, b.x) + __vsadu4(a.y, b.y) + __vsadu4(a.z, b.z) + __vsadu4(a.w, b.w);
}

__device__ __forceinline__ int f2i_trunc_x86(float x) {
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return (int)0x80000000;
    return __float2int_rz(x);
}

struct DenseArgs {
    const uint8_t *desc[2];
    const int32_t *owner[2];
    const PlaneRec *rec[2];
    const uint32_t *grid[2];
    float *D[2];
    int W, H, maxT, gw, gh, gwords, grid_size, disp_max, match_texture, plane_radius;
    int P[8];
};

// grid: (ceil(W/128), H, nf*2); blockIdx.z = 2*frame + side
__global__ void __launch_bounds__(128) k_dense(const DenseArgs a) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    const int side = blockIdx.z & 1;
    const int f = blockIdx.z >> 1;
    if (u >= a.W) return;
    const int W = a.W, H = a.H;
    const size_t N = (size_t)W * H;
    const size_t pix = (size_t)v * W + u;
    float *D = a.D[side] + (size_t)f * N;

    float out = -10.f;  // elas.cpp:820-826: pixels nobody writes keep -10
    const int o = a.owner[side][(size_t)f * N + pix];
    if (o >= 0 && u >= 2 && u < W - 2) {  // elas.cpp:714
        const int row = max(min(v, H - 3), 2);  // elas.cpp:718
        const uint4 *own = reinterpret_cast<const uint4 *>(a.desc[side]) + (size_t)f * N + (size_t)row * W;
        const uint4 *oth = reinterpret_cast<const uint4 *>(a.desc[side ^ 1]) + (size_t)f * N + (size_t)row * W;
        const uint4 c = __ldg(own + u);
        const uint4 k128 = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
        if ((int)sad16(c, k128) >= a.match_texture) {  // elas.cpp:731-736
            const PlaneRec pr = a.rec[side][(size_t)f * a.maxT + o];
            // elas.cpp:739: (a*u + b*v) + c in f32, separate roundings, truncation like cvttss2si
            const float fp = __fadd_rn(__fadd_rn(__fmul_rn(pr.a, (float)u), __fmul_rn(pr.b, (float)v)), pr.c);
            const int d_plane = f2i_trunc_x86(fp);
            const int d_plane_min = max((int)((unsigned)d_plane - (unsigned)a.plane_radius), 0);
            const int d_plane_max = min((int)((unsigned)d_plane + (unsigned)a.plane_radius), a.disp_max);

            const int gx = u / a.grid_size, gy = v / a.grid_size;  // u, v >= 0 so this equals the float floor
            const uint32_t *cell = a.grid[side] + ((size_t)f * a.gw * a.gh + (size_t)gy * a.gw + gx) * a.gwords;

            int min_val = 10000, min_d = -1;
            // (i) grid candidates outside the band, ascending (elas.cpp:759-767 / 778-786)
            for (int w = 0; w < a.gwords; w++) {
                uint32_t bits = __ldg(cell + w);
                while (bits) {
                    const int b = __ffs(bits) - 1;
                    bits &= bits - 1;
                    const int d = (w << 5) + b;
                    if (d >= d_plane_min && d <= d_plane_max) continue;
                    const int uw = side ? u + d : u - d;
                    if (uw < 2 || uw >= W - 2) continue;
                    const int val = (int)sad16(c, __ldg(oth + uw));
                    if (val < min_val) {
                        min_val = val;
                        min_d = d;
                    }
                }
            }
            // (ii) the plane band with the prior (elas.cpp:768-774 / 787-793)
            for (int d = d_plane_min; d <= d_plane_max; d++) {
                const int uw = side ? u + d : u - d;
                if (uw < 2 || uw >= W - 2) continue;
                const int val = (int)sad16(c, __ldg(oth + uw)) + (pr.valid ? a.P[abs(d - d_plane)] : 0);
                if (val < min_val) {
                    min_val = val;
                    min_d = d;
                }
            }
            out = min_d >= 0 ? (float)min_d : -1.f;  // elas.cpp:797-800
        }
    }
    D[pix] = out;
}

}  // namespace

int launch_dense(const Dims &d, const svb_params &p, const uint8_t *desc1, const uint8_t *desc2, const int32_t *owner1, const int32_t *owner2,
                 const PlaneRec *rec1, const PlaneRec *rec2, const uint32_t *grid1, const uint32_t *grid2, float *D1, float *D2, int nf,
                 cudaStream_t s) {
    if (nf <= 0) return SVB_OK;
    DenseArgs a;
    a.desc[0] = desc1;
    a.desc[1] = desc2;
    a.owner[0] = owner1;
    a.owner[1] = owner2;
    a.rec[0] = rec1;
    a.rec[1] = rec2;
    a.grid[0] = grid1;
    a.grid[1] = grid2;
    a.D[0] = D1;
    a.D[1] = D2;
    a.W = d.W;
    a.H = d.H;
    a.maxT = d.maxT;
    a.gw = d.gw;
    a.gh = d.gh;
    a.gwords = d.gwords;
    a.grid_size = p.grid_size;
    a.disp_max = p.disp_max;
    a.match_texture = p.match_texture;
    a.plane_radius = d.plane_radius;
    for (int i = 0; i < 8; i++) a.P[i] = d.P[i];
    dim3 grid((d.W + 127) / 128, d.H, nf * 2);
    k_dense<<<grid, 128, 0, s>>>(a);
    SVB_LAUNCH_CHECK();
    return SVB_OK;
}

}  // namespace svb
