// Deterministic synthetic rectified stereo pairs of known disparity (bench + parity inputs).
// Host C++ only; it produces INPUTS, it is not part of the matching path.
//
//   seed    = 0x5EED0000 + frame_index into SplitMix64
//   texture = 3 octaves of value noise (periods 4, 16, 64 px; amplitudes 48, 32, 24) around 128, plus uniform
//             noise +-6, clamped to u8, defined on an extended canvas (W + 96 columns)
//   scene   = three fronto-parallel bands with integer disparities {8, 24, 48} (rows [0,H/3), [H/3,2H/3), rest),
//             or, with `slanted`, d(u) = 10 + 0.03 u
//   left(u,v) = canvas(u,v);  right(x,v) = canvas(x + d, v) with independent +-2 noise (for the slanted plane the
//             canvas is sampled at the real-valued left column that maps to x)
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/elas_b200.h"

namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double sym() { return 2.0 * unit() - 1.0; }                                    // [-1,1)
};

inline uint8_t clamp_u8(double x) {
    long v = lrint(x);
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

}  // namespace

extern "C" int svb_synth_pair(int frame_index, int W, int H, int slanted, uint8_t *left, uint8_t *right) {
    if (!left || !right || W < 16 || H < 16) return SVB_ERR_ARG;
    const int EXT = 96;
    const int CW = W + EXT;
    SplitMix64 rng(0x5EED0000ull + (uint64_t)(uint32_t)frame_index);
    std::vector<float> canvas((size_t)CW * H, 128.f);
    const int periods[3] = {4, 16, 64};
    const double amps[3] = {48, 32, 24};
    for (int o = 0; o < 3; o++) {
        const int p = periods[o];
        const int lw = CW / p + 2, lh = H / p + 2;
        std::vector<float> lat((size_t)lw * lh);
        for (auto &x : lat) x = (float)rng.sym();
        for (int v = 0; v < H; v++) {
            const int y0 = v / p;
            const float fy = (float)(v % p) / p;
            const float sy = fy * fy * (3 - 2 * fy);
            for (int u = 0; u < CW; u++) {
                const int x0 = u / p;
                const float fx = (float)(u % p) / p;
                const float sx = fx * fx * (3 - 2 * fx);
                const float a = lat[(size_t)y0 * lw + x0], b = lat[(size_t)y0 * lw + x0 + 1];
                const float c = lat[(size_t)(y0 + 1) * lw + x0], d = lat[(size_t)(y0 + 1) * lw + x0 + 1];
                const float top = a + (b - a) * sx, bot = c + (d - c) * sx;
                canvas[(size_t)v * CW + u] += (float)amps[o] * (top + (bot - top) * sy);
            }
        }
    }
    std::vector<uint8_t> cv8((size_t)CW * H);
    for (size_t i = 0; i < cv8.size(); i++) cv8[i] = clamp_u8(canvas[i] + 6.0 * rng.sym());
    for (int v = 0; v < H; v++)
        for (int u = 0; u < W; u++) left[(size_t)v * W + u] = cv8[(size_t)v * CW + u];
    for (int v = 0; v < H; v++) {
        const int band = v < H / 3 ? 0 : (v < 2 * H / 3 ? 1 : 2);
        const int dband = band == 0 ? 8 : (band == 1 ? 24 : 48);
        for (int x = 0; x < W; x++) {
            double val;
            if (!slanted) {
                val = cv8[(size_t)v * CW + x + dband];
            } else {
                // left column u maps to x = u - (10 + 0.03 u)  =>  u = (x + 10) / 0.97
                const double u = (x + 10.0) / 0.97;
                int u0 = (int)floor(u);
                if (u0 > CW - 2) u0 = CW - 2;
                const double f = u - u0;
                val = (1 - f) * cv8[(size_t)v * CW + u0] + f * cv8[(size_t)v * CW + u0 + 1];
            }
            right[(size_t)v * W + x] = clamp_u8(val + 2.0 * rng.sym());
        }
    }
    return SVB_OK;
}
