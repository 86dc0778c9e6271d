// TEST INFRASTRUCTURE ONLY.  C-callable stage taps around the UNMODIFIED reference serial ELAS
// (src/serial_includes/elas/elas.cpp + src/common_includes/elas/*.cpp), compiled where the sources
// lie under /root/reference by oracle/build_ref.sh into oracle/_ref/libelas_ref.so.
// No reference source is copied into this repository: the reference translation unit is
// #included from its original location so that its private/inline stage functions can be
// called one by one (SURVEY.md 8c "Stage taps").
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the resulting library.  The product (csrc/) never links or dlopens it.

#ifdef ORACLE_REF_OMP
// the reference's OpenMP variant (`make omp=1`): timing baseline only (racy, not a parity target); only process() is wrapped,
// and `private` must stay a keyword for its `#pragma omp ... private(...)` clauses
#include "omp_includes/elas/elas.cpp"
#else
#define private public
#include "serial_includes/elas/elas.cpp"   // -I /root/reference/src
#undef private
#endif

#include <chrono>

extern "C" {

// POD mirror of Elas::parameters (src/parallel_includes/elas/elas.h:58-83), all 4-byte fields so
// that ctypes, the oracle port and the product C-ABI (include/elas_b200.h: svb_params) share it.
struct ref_params {
    int32_t disp_min, disp_max;
    float support_threshold;
    int32_t support_texture, candidate_stepsize, incon_window_size, incon_threshold, incon_min_support;
    int32_t add_corners, grid_size;
    float beta, gamma, sigma, sradius;
    int32_t match_texture, lr_threshold;
    float speckle_sim_threshold;
    int32_t speckle_size, ipol_gap_width;
    int32_t filter_median, filter_adaptive_mean, postprocess_only_left, subsampling;
};

}  // extern "C"

static Elas::parameters to_ref(const ref_params *p) {
    Elas::parameters q;
    q.disp_min = p->disp_min;
    q.disp_max = p->disp_max;
    q.support_threshold = p->support_threshold;
    q.support_texture = p->support_texture;
    q.candidate_stepsize = p->candidate_stepsize;
    q.incon_window_size = p->incon_window_size;
    q.incon_threshold = p->incon_threshold;
    q.incon_min_support = p->incon_min_support;
    q.add_corners = p->add_corners != 0;
    q.grid_size = p->grid_size;
    q.beta = p->beta;
    q.gamma = p->gamma;
    q.sigma = p->sigma;
    q.sradius = p->sradius;
    q.match_texture = p->match_texture;
    q.lr_threshold = p->lr_threshold;
    q.speckle_sim_threshold = p->speckle_sim_threshold;
    q.speckle_size = p->speckle_size;
    q.ipol_gap_width = p->ipol_gap_width;
    q.filter_median = p->filter_median != 0;
    q.filter_adaptive_mean = p->filter_adaptive_mean != 0;
    q.postprocess_only_left = p->postprocess_only_left != 0;
    q.subsampling = p->subsampling != 0;
    return q;
}

#ifndef ORACLE_REF_OMP  // ---- helpers of the stage taps (serial oracle only) ----
static Elas make_elas(const ref_params *p, int W, int H) {
    Elas e(to_ref(p));
    e.width = W;
    e.height = H;
    e.bpl = W + 15 - (W - 1) % 16;
    e.I1 = e.I2 = nullptr;
    return e;
}

static std::vector<Elas::support_pt> to_pts(const int32_t *pts, int n) {
    std::vector<Elas::support_pt> v;
    v.reserve(n);
    for (int i = 0; i < n; i++) v.push_back(Elas::support_pt(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    return v;
}

static std::vector<Elas::triangle> to_tris(const int32_t *tri, const float *planes, int m) {
    std::vector<Elas::triangle> v;
    v.reserve(m);
    for (int i = 0; i < m; i++) {
        Elas::triangle t(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
        if (planes) {
            t.t1a = planes[6 * i + 0];
            t.t1b = planes[6 * i + 1];
            t.t1c = planes[6 * i + 2];
            t.t2a = planes[6 * i + 3];
            t.t2b = planes[6 * i + 4];
            t.t2c = planes[6 * i + 5];
        } else {
            t.t1a = t.t1b = t.t1c = t.t2a = t.t2b = t.t2c = 0;
        }
        v.push_back(t);
    }
    return v;
}

#endif  // ORACLE_REF_OMP

extern "C" {

void ref_default_params(int setting, ref_params *out) {
    Elas::parameters q(setting == 0 ? Elas::ROBOTICS : Elas::MIDDLEBURY);
    out->disp_min = q.disp_min;
    out->disp_max = q.disp_max;
    out->support_threshold = q.support_threshold;
    out->support_texture = q.support_texture;
    out->candidate_stepsize = q.candidate_stepsize;
    out->incon_window_size = q.incon_window_size;
    out->incon_threshold = q.incon_threshold;
    out->incon_min_support = q.incon_min_support;
    out->add_corners = q.add_corners;
    out->grid_size = q.grid_size;
    out->beta = q.beta;
    out->gamma = q.gamma;
    out->sigma = q.sigma;
    out->sradius = q.sradius;
    out->match_texture = q.match_texture;
    out->lr_threshold = q.lr_threshold;
    out->speckle_sim_threshold = q.speckle_sim_threshold;
    out->speckle_size = q.speckle_size;
    out->ipol_gap_width = q.ipol_gap_width;
    out->filter_median = q.filter_median;
    out->filter_adaptive_mean = q.filter_adaptive_mean;
    out->postprocess_only_left = q.postprocess_only_left;
    out->subsampling = q.subsampling;
}

// Elas::process (src/serial_includes/elas/elas.cpp:31).  Returns seconds spent inside process().
double ref_process(const ref_params *p, const uint8_t *I1, const uint8_t *I2, int W, int H, int stride, float *D1, float *D2) {
    Elas e(to_ref(p));
    const int32_t dims[3] = {W, H, stride};
    auto t0 = std::chrono::steady_clock::now();
    e.process((uint8_t *)I1, (uint8_t *)I2, D1, D2, dims);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

#ifndef ORACLE_REF_OMP  // ---- stage taps (serial oracle only) ----
// Descriptor (src/common_includes/elas/descriptor.cpp:30) on the bpl-padded copy made like
// process() does (elas.cpp:33-50).  out: 16*W*H bytes.
void ref_descriptor(const uint8_t *I, int W, int H, int stride, int subsampling, uint8_t *out) {
    int bpl = W + 15 - (W - 1) % 16;
    uint8_t *Ip = (uint8_t *)_mm_malloc(bpl * H, 16);
    for (int v = 0; v < H; v++) memcpy(Ip + v * bpl, I + v * stride, W);
    {
        Descriptor d(Ip, W, H, bpl, subsampling != 0);
        memcpy(out, d.I_desc, (size_t)16 * W * H);
    }
    _mm_free(Ip);
}

// candidate grid size as computed at elas.cpp:376-386
void ref_dcan_dims(const ref_params *p, int W, int H, int *cw, int *ch, int *step) {
    int s = p->candidate_stepsize;
    if (p->subsampling) s += s % 2;
    int w = 0, h = 0;
    for (int u = 0; u < W; u += s) w++;
    for (int v = 0; v < H; v += s) h++;
    *cw = w;
    *ch = h;
    *step = s;
}

// D_can right after the candidate loop (elas.cpp:394-411), before the three filters.
void ref_support_raw(const ref_params *p, const uint8_t *desc1, const uint8_t *desc2, int W, int H, int16_t *dcan) {
    Elas e = make_elas(p, W, H);
    int cw, ch, s;
    ref_dcan_dims(p, W, H, &cw, &ch, &s);
    memset(dcan, 0, sizeof(int16_t) * cw * ch);
    for (int uc = 1; uc < cw; uc++)
        for (int vc = 1; vc < ch; vc++) {
            int u = uc * s, v = vc * s;
            int16_t r = -1;
            int16_t d = e.computeMatchingDisparity(u, v, (uint8_t *)desc1, (uint8_t *)desc2, false);
            if (d >= 0) {
                int u2 = u - d;
                int16_t d2 = e.computeMatchingDisparity(u2, v, (uint8_t *)desc1, (uint8_t *)desc2, true);
                if (d2 >= 0 && abs(d - d2) <= p->lr_threshold) r = d;
            }
            dcan[vc * cw + uc] = r;
        }
}

// the three in-place filters (elas.cpp:414-420); which: 0 inconsistent, 1 redundant vertical, 2 redundant horizontal
void ref_support_filter(const ref_params *p, int16_t *dcan, int cw, int ch, int which) {
    Elas e = make_elas(p, 16, 16);
    if (which == 0) e.removeInconsistentSupportPoints(dcan, cw, ch);
    if (which == 1) e.removeRedundantSupportPoints(dcan, cw, ch, 5, 1, true);
    if (which == 2) e.removeRedundantSupportPoints(dcan, cw, ch, 5, 1, false);
}

// computeSupportMatches (elas.cpp:373).  pts: cap x {u,v,d}.  Returns the number of support points.
int ref_support(const ref_params *p, const uint8_t *desc1, const uint8_t *desc2, int W, int H, int32_t *pts, int cap) {
    Elas e = make_elas(p, W, H);
    std::vector<Elas::support_pt> s = e.computeSupportMatches((uint8_t *)desc1, (uint8_t *)desc2);
    int n = (int)s.size();
    for (int i = 0; i < n && i < cap; i++) {
        pts[3 * i] = s[i].u;
        pts[3 * i + 1] = s[i].v;
        pts[3 * i + 2] = s[i].d;
    }
    return n;
}

// computeDelaunayTriangulation (elas.cpp:442).  tri: cap x {c1,c2,c3}.  Returns the triangle count.
int ref_delaunay(const int32_t *pts, int n, int right_image, int32_t *tri, int cap) {
    ref_params dummy;
    ref_default_params(0, &dummy);
    Elas e = make_elas(&dummy, 16, 16);
    std::vector<Elas::triangle> t = e.computeDelaunayTriangulation(to_pts(pts, n), right_image);
    int m = (int)t.size();
    for (int i = 0; i < m && i < cap; i++) {
        tri[3 * i] = t[i].c1;
        tri[3 * i + 1] = t[i].c2;
        tri[3 * i + 2] = t[i].c3;
    }
    return m;
}

// triangulate("zQB") (src/common_includes/elas/triangle.cpp:8116) on arbitrary float points.
int ref_triangulate_xy(const float *xy, int n, int32_t *tri, int cap) {
    struct triangulateio in, out;
    memset(&in, 0, sizeof(in));
    memset(&out, 0, sizeof(out));
    in.numberofpoints = n;
    in.pointlist = (float *)malloc(sizeof(float) * 2 * n);
    memcpy(in.pointlist, xy, sizeof(float) * 2 * n);
    char sw[] = "zQB";
    triangulate(sw, &in, &out, NULL);
    int m = out.numberoftriangles;
    for (int i = 0; i < m && i < cap; i++) {
        tri[3 * i] = out.trianglelist[3 * i];
        tri[3 * i + 1] = out.trianglelist[3 * i + 1];
        tri[3 * i + 2] = out.trianglelist[3 * i + 2];
    }
    free(in.pointlist);
    free(out.pointlist);
    free(out.trianglelist);
    return m;
}

// computeDisparityPlanes (elas.cpp:503).  planes: m x {t1a,t1b,t1c,t2a,t2b,t2c}
void ref_planes(const int32_t *pts, int n, const int32_t *tri, int m, float *planes) {
    ref_params dummy;
    ref_default_params(0, &dummy);
    Elas e = make_elas(&dummy, 16, 16);
    std::vector<Elas::triangle> t = to_tris(tri, nullptr, m);
    e.computeDisparityPlanes(to_pts(pts, n), t, 0);  // the right_image argument is unused by the reference
    for (int i = 0; i < m; i++) {
        planes[6 * i + 0] = t[i].t1a;
        planes[6 * i + 1] = t[i].t1b;
        planes[6 * i + 2] = t[i].t1c;
        planes[6 * i + 3] = t[i].t2a;
        planes[6 * i + 4] = t[i].t2b;
        planes[6 * i + 5] = t[i].t2c;
    }
}

void ref_grid_dims(const ref_params *p, int W, int H, int *gw, int *gh) {
    *gw = (int32_t)ceil((float)W / (float)p->grid_size);
    *gh = (int32_t)ceil((float)H / (float)p->grid_size);
}

// createGrid (elas.cpp:577).  grid: gh*gw*(disp_max+2) int32, zero-filled here like the calloc at :91.
void ref_grid(const ref_params *p, const int32_t *pts, int n, int W, int H, int right_image, int32_t *grid) {
    Elas e = make_elas(p, W, H);
    int gw, gh;
    ref_grid_dims(p, W, H, &gw, &gh);
    int32_t dims[3] = {p->disp_max + 2, gw, gh};
    memset(grid, 0, sizeof(int32_t) * (size_t)(p->disp_max + 2) * gw * gh);
    e.createGrid(to_pts(pts, n), grid, dims, right_image != 0);
}

// computeDisparity (elas.cpp:804)
void ref_disparity(const ref_params *p, const int32_t *pts, int n, const int32_t *tri, const float *planes, int m, const int32_t *grid,
                   const uint8_t *desc1, const uint8_t *desc2, int W, int H, int right_image, float *D) {
    Elas e = make_elas(p, W, H);
    int gw, gh;
    ref_grid_dims(p, W, H, &gw, &gh);
    int32_t dims[3] = {p->disp_max + 2, gw, gh};
    e.computeDisparity(to_pts(pts, n), to_tris(tri, planes, m), (int32_t *)grid, dims, (uint8_t *)desc1, (uint8_t *)desc2, right_image != 0, D);
}

// post-processing stages, in place (elas.cpp:946, 1013, 1126, 1297, 1496)
void ref_lr_check(const ref_params *p, int W, int H, float *D1, float *D2) {
    Elas e = make_elas(p, W, H);
    e.leftRightConsistencyCheck(D1, D2);
}
void ref_remove_small_segments(const ref_params *p, int W, int H, float *D) {
    Elas e = make_elas(p, W, H);
    e.removeSmallSegments(D);
}
void ref_gap_interpolation(const ref_params *p, int W, int H, float *D) {
    Elas e = make_elas(p, W, H);
    e.gapInterpolation(D);
}
void ref_adaptive_mean(const ref_params *p, int W, int H, float *D) {
    Elas e = make_elas(p, W, H);
    e.adaptiveMean(D);
}
void ref_median(const ref_params *p, int W, int H, float *D) {
    Elas e = make_elas(p, W, H);
    e.median(D);
}

#endif  // ORACLE_REF_OMP

const char *ref_build_flags() {
#ifdef ORACLE_REF_FLAGS
    return ORACLE_REF_FLAGS;
#else
    return "unknown";
#endif
}

}  // extern "C"
