// Calibration without OpenCV: reader for the OpenCV-FileStorage YAML subset the reference's calibration files use,
// and a restatement of cv::stereoRectify (CALIB_ZERO_DISPARITY, any alpha) that yields R1, R2, P1, P2 and the
// disparity-to-depth matrix Q.
//
// Replaces, in the reference driver (src/parallel_includes/main/stereo_vision.cu):
//   :536-545  calib_file["K1"] >> K1 ... calib_file["XT"] >> XT          (cv::FileStorage)
//   :368-447  findRectificationMap(): K1/K2 divided by scale_factor, cv::stereoRectify(..., alpha = 0, finalSize)
// cv::stereoRectify is third-party arithmetic (OpenCV, not under the reference tree; the reference does not pin a
// version, its CI installs 4.4.0).  Its published algorithm (calib3d: cvStereoRectify, icvGetRectangles,
// cvUndistortPoints with 5 fixed iterations, cvRodrigues2) is restated here in the form OpenCV 4.13 has it (float32
// corner arrays, a double-precision 9 x 9 grid over [0, w-1] x [0, h-1] in getRectangles), and pinned against python cv2 4.13 outputs committed under tests/golden/
// (tests/golden/make_calib_golden.py).  Host code, double precision.
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/elas_b200.h"

namespace svb {
void set_error(const char *fmt, ...);
}

namespace {

// ---------------------------------------------------------------------------------------------- YAML subset
struct Node {
    std::string key;
    std::string body;  // everything after "key:" up to the next top-level key
};

bool read_file(const char *path, std::string *out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out->append(buf, n);
    fclose(f);
    return true;
}

std::vector<Node> split_top_level(const std::string &text) {
    std::vector<Node> nodes;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        const bool top = !line.empty() && line[0] != ' ' && line[0] != '\t' && line[0] != '%' && line[0] != '#' && line[0] != '-';
        size_t colon = line.find(':');
        if (top && colon != std::string::npos) {
            Node n;
            n.key = line.substr(0, colon);
            n.body = line.substr(colon + 1) + "\n";
            nodes.push_back(n);
        } else if (!nodes.empty() && !line.empty() && line[0] != '%' && line.compare(0, 3, "---") != 0) {
            nodes.back().body += line + "\n";
        }
        pos = eol + 1;
    }
    return nodes;
}

// numbers between the first '[' and the matching ']' (or the whole string for a scalar)
std::vector<double> parse_numbers(const std::string &s) {
    std::vector<double> v;
    size_t a = s.find('['), b = s.rfind(']');
    std::string t = (a != std::string::npos && b != std::string::npos && b > a) ? s.substr(a + 1, b - a - 1) : s;
    const char *p = t.c_str();
    while (*p) {
        while (*p && !(isdigit((unsigned char)*p) || *p == '-' || *p == '+' || *p == '.')) p++;
        if (!*p) break;
        char *end = nullptr;
        double x = strtod(p, &end);
        if (end == p) {
            p++;
            continue;
        }
        v.push_back(x);
        p = end;
    }
    return v;
}

int field_int(const std::string &body, const char *name, int dflt) {
    size_t p = body.find(name);
    if (p == std::string::npos) return dflt;
    p = body.find(':', p);
    if (p == std::string::npos) return dflt;
    return atoi(body.c_str() + p + 1);
}

// returns the element count, rows/cols through the pointers (a plain sequence is 1 x n)
int node_matrix(const Node &n, std::vector<double> *data, int *rows, int *cols) {
    if (n.body.find("!!opencv-matrix") != std::string::npos) {
        *rows = field_int(n.body, "rows", 0);
        *cols = field_int(n.body, "cols", 0);
        size_t d = n.body.find("data");
        *data = parse_numbers(d == std::string::npos ? std::string() : n.body.substr(d));
    } else {
        *data = parse_numbers(n.body);
        *rows = 1;
        *cols = (int)data->size();
    }
    return (int)data->size();
}

// ---------------------------------------------------------------------------------------------- small linear algebra
struct M3 {
    double m[3][3];
};

M3 mul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
M3 transpose(const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
void mulv(const M3 &a, const double v[3], double out[3]) {
    double t[3];
    for (int i = 0; i < 3; i++) t[i] = a.m[i][0] * v[0] + a.m[i][1] * v[1] + a.m[i][2] * v[2];
    out[0] = t[0];
    out[1] = t[1];
    out[2] = t[2];
}
double det(const M3 &a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
           a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
M3 inverse(const M3 &a) {
    const double d = det(a), id = 1.0 / d;
    M3 r;
    r.m[0][0] = (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) * id;
    r.m[0][1] = (a.m[0][2] * a.m[2][1] - a.m[0][1] * a.m[2][2]) * id;
    r.m[0][2] = (a.m[0][1] * a.m[1][2] - a.m[0][2] * a.m[1][1]) * id;
    r.m[1][0] = (a.m[1][2] * a.m[2][0] - a.m[1][0] * a.m[2][2]) * id;
    r.m[1][1] = (a.m[0][0] * a.m[2][2] - a.m[0][2] * a.m[2][0]) * id;
    r.m[1][2] = (a.m[0][2] * a.m[1][0] - a.m[0][0] * a.m[1][2]) * id;
    r.m[2][0] = (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]) * id;
    r.m[2][1] = (a.m[0][1] * a.m[2][0] - a.m[0][0] * a.m[2][1]) * id;
    r.m[2][2] = (a.m[0][0] * a.m[1][1] - a.m[0][1] * a.m[1][0]) * id;
    return r;
}

// nearest rotation (the U * V^T of cvRodrigues2's SVD step) by Newton iteration on the polar decomposition
M3 orthonormalize(const M3 &a) {
    M3 x = a;
    for (int it = 0; it < 20; it++) {
        const M3 xit = transpose(inverse(x));
        M3 n;
        double diff = 0;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                n.m[i][j] = 0.5 * (x.m[i][j] + xit.m[i][j]);
                diff = std::max(diff, fabs(n.m[i][j] - x.m[i][j]));
            }
        x = n;
        if (diff < 1e-16) break;
    }
    return x;
}

// cvRodrigues2, vector -> matrix
M3 rodrigues_vec(const double r_in[3]) {
    M3 R;
    const double theta = sqrt(r_in[0] * r_in[0] + r_in[1] * r_in[1] + r_in[2] * r_in[2]);
    if (theta < DBL_EPSILON) {
        memset(&R, 0, sizeof(R));
        R.m[0][0] = R.m[1][1] = R.m[2][2] = 1;
        return R;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1. - c, itheta = 1. / theta;
    const double r[3] = {r_in[0] * itheta, r_in[1] * itheta, r_in[2] * itheta};
    const double rrt[3][3] = {{r[0] * r[0], r[0] * r[1], r[0] * r[2]}, {r[0] * r[1], r[1] * r[1], r[1] * r[2]}, {r[0] * r[2], r[1] * r[2], r[2] * r[2]}};
    const double rx[3][3] = {{0, -r[2], r[1]}, {r[2], 0, -r[0]}, {-r[1], r[0], 0}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R.m[i][j] = c * (i == j ? 1.0 : 0.0) + c1 * rrt[i][j] + s * rx[i][j];
    return R;
}

// cvRodrigues2, matrix -> vector
void rodrigues_mat(const M3 &Rin, double out[3]) {
    const M3 R = orthonormalize(Rin);
    double r[3] = {R.m[2][1] - R.m[1][2], R.m[0][2] - R.m[2][0], R.m[1][0] - R.m[0][1]};
    const double s = sqrt((r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * 0.25);
    double c = (R.m[0][0] + R.m[1][1] + R.m[2][2] - 1) * 0.5;
    c = c > 1. ? 1. : c < -1. ? -1. : c;
    const double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) {
            out[0] = out[1] = out[2] = 0;
        } else {
            double t;
            t = (R.m[0][0] + 1) * 0.5;
            r[0] = sqrt(std::max(t, 0.));
            t = (R.m[1][1] + 1) * 0.5;
            r[1] = sqrt(std::max(t, 0.)) * (R.m[0][1] < 0 ? -1. : 1.);
            t = (R.m[2][2] + 1) * 0.5;
            r[2] = sqrt(std::max(t, 0.)) * (R.m[0][2] < 0 ? -1. : 1.);
            if (fabs(r[0]) < fabs(r[1]) && fabs(r[0]) < fabs(r[2]) && (R.m[1][2] > 0) != (r[1] * r[2] > 0)) r[2] = -r[2];
            const double nrm = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
            const double k = theta / nrm;
            out[0] = r[0] * k;
            out[1] = r[1] * k;
            out[2] = r[2] * k;
        }
    } else {
        const double vth = 1 / (2 * s) * theta;
        out[0] = r[0] * vth;
        out[1] = r[1] * vth;
        out[2] = r[2] * vth;
    }
}

// cvUndistortPoints as cvStereoRectify calls it: 5 fixed iterations, result through (R, P[:, :3]) and stored as float32
struct Pt2f {
    float x, y;
};
struct Pt2d {
    double x, y;
};

// T = float mirrors the CV_32FC2 point arrays of the corner step, T = double the CV_64FC2 grid of getRectangles
template <typename PT>
void undistort_points(PT *pts, int n, const double K[9], const double *D, int nd, const M3 *R, const double *P, int p_cols) {
    double k[14] = {0};
    for (int i = 0; i < nd && i < 14; i++) k[i] = D[i];
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double ifx = 1. / fx, ify = 1. / fy;
    M3 RR;
    memset(&RR, 0, sizeof(RR));
    RR.m[0][0] = RR.m[1][1] = RR.m[2][2] = 1;
    if (R) RR = *R;
    if (P) {
        M3 PP;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) PP.m[i][j] = P[i * p_cols + j];
        RR = mul(PP, RR);
    }
    for (int i = 0; i < n; i++) {
        double x = pts[i].x, y = pts[i].y;
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        if (nd > 0) {
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; j++) {
                const double r2 = x * x + y * y;
                const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) {  // test: undistortPoints.regression_14583
                    x = (u - cx) * ifx;
                    y = (v - cy) * ify;
                    break;
                }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        const double xx = RR.m[0][0] * x + RR.m[0][1] * y + RR.m[0][2];
        const double yy = RR.m[1][0] * x + RR.m[1][1] * y + RR.m[1][2];
        const double ww = 1. / (RR.m[2][0] * x + RR.m[2][1] * y + RR.m[2][2]);
        pts[i].x = (decltype(pts[i].x))(xx * ww);
        pts[i].y = (decltype(pts[i].y))(yy * ww);
    }
}

struct Rectd {
    double x, y, width, height;
};

// getRectangles (OpenCV >= 4.5.x form: a 9 x 9 grid over [0, w-1] x [0, h-1] in double precision)
void get_rectangles(const double K[9], const double *D, int nd, const M3 &R, const double P[12], int w, int h, Rectd *inner, Rectd *outer) {
    const int N = 9;
    Pt2d pts[N * N];
    int k = 0;
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++) {
            pts[k].x = (double)x * (w - 1) / (N - 1);
            pts[k].y = (double)y * (h - 1) / (N - 1);
            k++;
        }
    undistort_points(pts, N * N, K, D, nd, &R, P, 4);
    double iX0 = -FLT_MAX, iX1 = FLT_MAX, iY0 = -FLT_MAX, iY1 = FLT_MAX;
    double oX0 = FLT_MAX, oX1 = -FLT_MAX, oY0 = FLT_MAX, oY1 = -FLT_MAX;
    k = 0;
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++) {
            const Pt2d p = pts[k++];
            oX0 = std::min(oX0, p.x);
            oX1 = std::max(oX1, p.x);
            oY0 = std::min(oY0, p.y);
            oY1 = std::max(oY1, p.y);
            if (x == 0) iX0 = std::max(iX0, p.x);
            if (x == N - 1) iX1 = std::min(iX1, p.x);
            if (y == 0) iY0 = std::max(iY0, p.y);
            if (y == N - 1) iY1 = std::min(iY1, p.y);
        }
    *inner = Rectd{iX0, iY0, iX1 - iX0, iY1 - iY0};
    *outer = Rectd{oX0, oY0, oX1 - oX0, oY1 - oY0};
}

}  // namespace

extern "C" {

int svb_calib_load_yaml(const char *path, svb_calibration *out) {
    if (!path || !out) return SVB_ERR_ARG;
    std::string text;
    if (!read_file(path, &text)) {
        svb::set_error("cannot read calibration file %s", path);
        return SVB_ERR_ARG;
    }
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 3; i++) out->XR[4 * i] = 1.0;  // a file without XR/XT means "camera frame"
    bool have[6] = {false, false, false, false, false, false};
    for (const Node &n : split_top_level(text)) {
        std::vector<double> v;
        int rows = 0, cols = 0;
        node_matrix(n, &v, &rows, &cols);
        auto take = [&](double *dst, size_t want) {
            if (v.size() != want) return false;
            for (size_t i = 0; i < want; i++) dst[i] = v[i];
            return true;
        };
        if (n.key == "K1") have[0] = take(out->K1, 9);
        else if (n.key == "K2") have[1] = take(out->K2, 9);
        else if (n.key == "R") have[4] = take(out->R, 9);
        else if (n.key == "T") have[5] = take(out->T, 3);
        else if (n.key == "XR") take(out->XR, 9);
        else if (n.key == "XT") take(out->XT, 3);
        else if (n.key == "D1" || n.key == "D2") {
            if (v.size() > 14) {
                svb::set_error("%s: %zu distortion coefficients (at most 14)", n.key.c_str(), v.size());
                return SVB_ERR_ARG;
            }
            double *dst = n.key == "D1" ? out->D1 : out->D2;
            for (size_t i = 0; i < v.size(); i++) dst[i] = v[i];
            (n.key == "D1" ? out->n_d1 : out->n_d2) = (int)v.size();
            have[n.key == "D1" ? 2 : 3] = true;
        }
    }
    static const char *names[6] = {"K1", "K2", "D1", "D2", "R", "T"};
    for (int i = 0; i < 6; i++)
        if (!have[i]) {
            svb::set_error("%s: missing or malformed entry %s", path, names[i]);
            return SVB_ERR_ARG;
        }
    return SVB_OK;
}

int svb_stereo_rectify(const svb_calibration *cal, int calib_w, int calib_h, int new_w, int new_h, double scale_factor, double alpha, double *R1o,
                       double *R2o, double *P1o, double *P2o, double *Qo) {
    if (!cal || calib_w <= 0 || calib_h <= 0 || !(scale_factor > 0)) return SVB_ERR_ARG;
    if (new_w * new_h == 0) {
        new_w = calib_w;
        new_h = calib_h;
    }
    double K1[9], K2[9];
    memcpy(K1, cal->K1, sizeof(K1));
    memcpy(K2, cal->K2, sizeof(K2));
    // stereo_vision.cu:372-384: the first two rows of K1/K2 are divided by scale_factor
    for (int i = 0; i < 6; i++) {
        K1[i] /= scale_factor;
        K2[i] /= scale_factor;
    }
    M3 R;
    memcpy(R.m, cal->R, sizeof(R.m));
    double om[3];
    rodrigues_mat(R, om);
    for (int i = 0; i < 3; i++) om[i] *= -0.5;  // average rotation
    const M3 r_r = rodrigues_vec(om);
    double t[3];
    mulv(r_r, cal->T, t);
    const int idx = fabs(t[0]) > fabs(t[1]) ? 0 : 1;
    const double c = t[idx], nt = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
    if (!(nt > 0.0)) {
        svb::set_error("stereo_rectify: zero baseline");
        return SVB_ERR_ARG;
    }
    double uu[3] = {0, 0, 0};
    uu[idx] = c > 0 ? 1 : -1;
    // global Z rotation
    double ww[3] = {t[1] * uu[2] - t[2] * uu[1], t[2] * uu[0] - t[0] * uu[2], t[0] * uu[1] - t[1] * uu[0]};
    const double nw = sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
    if (nw > 0.0) {
        const double sc = acos(fabs(c) / nt) / nw;
        for (int i = 0; i < 3; i++) ww[i] *= sc;
    }
    const M3 wR = rodrigues_vec(ww);
    const M3 R1 = mul(wR, transpose(r_r));
    const M3 R2 = mul(wR, r_r);
    mulv(R2, cal->T, t);

    const int nx = calib_w, ny = calib_h;
    const double ratio_x = (double)new_w / calib_w / 2;
    const double ratio_y = (double)new_h / calib_h / 2;
    const double ratio = idx == 1 ? ratio_x : ratio_y;
    double fc_new = (K1[(idx ^ 1) * 4] + K2[(idx ^ 1) * 4]) * ratio;

    double cc_new[2][2];
    for (int k = 0; k < 2; k++) {
        const double *A = k == 0 ? K1 : K2;
        const double *Dk = k == 0 ? cal->D1 : cal->D2;
        const int nd = k == 0 ? cal->n_d1 : cal->n_d2;
        Pt2f pts[4];
        for (int i = 0; i < 4; i++) {
            const int j = (i < 2) ? 0 : 1;
            pts[i].x = (float)((i % 2) * (nx - 1));
            pts[i].y = (float)(j * (ny - 1));
        }
        undistort_points(pts, 4, A, Dk, nd, nullptr, nullptr, 0);
        // cvConvertPointsHomogeneous to float32 (x, y, 1), then cvProjectPoints2 with rotation R_k, zero translation,
        // camera matrix diag(fc_new, fc_new, 1) and no distortion, result stored as float32
        const M3 &Rk = k == 0 ? R1 : R2;
        // cvProjectPoints2 converts the rotation matrix to a vector and back (Rodrigues both ways)
        double rv[3];
        rodrigues_mat(Rk, rv);
        const M3 Rp = rodrigues_vec(rv);
        double sx = 0, sy = 0;
        for (int i = 0; i < 4; i++) {
            const double X = pts[i].x, Y = pts[i].y, Z = 1.0;
            const double x = Rp.m[0][0] * X + Rp.m[0][1] * Y + Rp.m[0][2] * Z;
            const double y = Rp.m[1][0] * X + Rp.m[1][1] * Y + Rp.m[1][2] * Z;
            double z = Rp.m[2][0] * X + Rp.m[2][1] * Y + Rp.m[2][2] * Z;
            z = z ? 1. / z : 1;
            const float px = (float)(x * z * fc_new + 0.0);
            const float py = (float)(y * z * fc_new + 0.0);
            sx += px;
            sy += py;
        }
        cc_new[k][0] = (nx - 1) / 2.0 - sx / 4;
        cc_new[k][1] = (ny - 1) / 2.0 - sy / 4;
    }
    // CALIB_ZERO_DISPARITY (the reference always passes it, stereo_vision.cu:447)
    cc_new[0][0] = cc_new[1][0] = (cc_new[0][0] + cc_new[1][0]) * 0.5;
    cc_new[0][1] = cc_new[1][1] = (cc_new[0][1] + cc_new[1][1]) * 0.5;

    double P1[12] = {0}, P2[12] = {0};
    P1[0] = P1[5] = fc_new;
    P1[2] = cc_new[0][0];
    P1[6] = cc_new[0][1];
    P1[10] = 1;
    memcpy(P2, P1, sizeof(P1));
    P2[2] = cc_new[1][0];
    P2[6] = cc_new[1][1];
    P2[idx * 4 + 3] = t[idx] * fc_new;  // baseline * focal length

    alpha = std::min(alpha, 1.);
    Rectd inner1, inner2, outer1, outer2;
    get_rectangles(K1, cal->D1, cal->n_d1, R1, P1, calib_w, calib_h, &inner1, &outer1);
    get_rectangles(K2, cal->D2, cal->n_d2, R2, P2, calib_w, calib_h, &inner2, &outer2);
    {
        const double cx1_0 = cc_new[0][0], cy1_0 = cc_new[0][1], cx2_0 = cc_new[1][0], cy2_0 = cc_new[1][1];
        const double cx1 = new_w * cx1_0 / calib_w, cy1 = new_h * cy1_0 / calib_h;
        const double cx2 = new_w * cx2_0 / calib_w, cy2 = new_h * cy2_0 / calib_h;
        double s = 1.;
        if (alpha >= 0) {
            double s0 = std::max(std::max(std::max((double)cx1 / (cx1_0 - inner1.x), (double)cy1 / (cy1_0 - inner1.y)),
                                          (double)(new_w - 1 - cx1) / (inner1.x + inner1.width - cx1_0)),
                                 (double)(new_h - 1 - cy1) / (inner1.y + inner1.height - cy1_0));
            s0 = std::max(std::max(std::max(std::max((double)cx2 / (cx2_0 - inner2.x), (double)cy2 / (cy2_0 - inner2.y)),
                                            (double)(new_w - 1 - cx2) / (inner2.x + inner2.width - cx2_0)),
                                   (double)(new_h - 1 - cy2) / (inner2.y + inner2.height - cy2_0)),
                          s0);
            double s1 = std::min(std::min(std::min((double)cx1 / (cx1_0 - outer1.x), (double)cy1 / (cy1_0 - outer1.y)),
                                          (double)(new_w - 1 - cx1) / (outer1.x + outer1.width - cx1_0)),
                                 (double)(new_h - 1 - cy1) / (outer1.y + outer1.height - cy1_0));
            s1 = std::min(std::min(std::min(std::min((double)cx2 / (cx2_0 - outer2.x), (double)cy2 / (cy2_0 - outer2.y)),
                                            (double)(new_w - 1 - cx2) / (outer2.x + outer2.width - cx2_0)),
                                   (double)(new_h - 1 - cy2) / (outer2.y + outer2.height - cy2_0)),
                          s1);
            s = s0 * (1 - alpha) + s1 * alpha;
        }
        fc_new *= s;
        cc_new[0][0] = cx1;
        cc_new[0][1] = cy1;
        cc_new[1][0] = cx2;
        cc_new[1][1] = cy2;
        P1[0] = P1[5] = fc_new;
        P1[2] = cx1;
        P1[6] = cy1;
        P2[0] = P2[5] = fc_new;
        P2[2] = cx2;
        P2[6] = cy2;
        P2[idx * 4 + 3] = s * P2[idx * 4 + 3];
    }
    if (R1o) memcpy(R1o, R1.m, sizeof(double) * 9);
    if (R2o) memcpy(R2o, R2.m, sizeof(double) * 9);
    if (P1o) memcpy(P1o, P1, sizeof(P1));
    if (P2o) memcpy(P2o, P2, sizeof(P2));
    if (Qo) {
        const double q[16] = {1, 0, 0, -cc_new[0][0], 0, 1, 0, -cc_new[0][1], 0, 0, 0, fc_new, 0, 0, -1. / t[idx],
                              (idx == 0 ? cc_new[0][0] - cc_new[1][0] : cc_new[0][1] - cc_new[1][1]) / t[idx]};
        memcpy(Qo, q, sizeof(q));
    }
    return SVB_OK;
}

}  // extern "C"
