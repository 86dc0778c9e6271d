// Host stage: divide-and-conquer Delaunay triangulation with alternating cuts on integer coordinates.
//
// The support points lie on a 5-pixel lattice, so they are massively co-circular and the Delaunay
// triangulation is not unique; the dense-matching prior depends on WHICH triangulation is chosen and the
// rasteriser on the ORDER of the triangle list (SURVEY.md finding 7, Appendix A).  This implementation therefore
// follows the decision rules of the reference's triangulator (Triangle 1.6 as vendored in
// src/common_includes/elas/triangle.cpp, switches "zQB", dwyer = 1):
//   - lexicographic (x, y) sort, duplicates dropped keeping the first of each run           (triangle.cpp:5882-5903)
//     (the sort is an unstable randomised quicksort; it is re-stated with its LCG so that the surviving
//      duplicate -- possible in the right image where x = u - d -- is the same one)         (:5183-5229, :3833-3836)
//   - alternating-axis median partition, subsets of <= 3 vertices x-sorted                   (:5243-5325, :5904-5913)
//   - recursion with ghost ("bounding") triangles, 2- and 3-vertex base cases                (:5670-5815)
//   - hull merge: strict ccw > 0 for the lower tangent, strict incircle > 0 for edge
//     deletion and for choosing the right candidate, horizontal-cut handle rotation          (:5362-5651)
//   - output = live non-ghost triangle records in creation order, corners (org, dest, apex)  (:7449-7500)
// Unlike the reference (float coordinates + adaptive-precision floating point predicates) every predicate here
// is evaluated exactly in 64-bit integers: |coordinates| < 2^13 keeps orient2d below 2^28 and incircle below 2^56.
// Triangle records are plain index triples in one flat arena (no pointer pool, no per-call malloc).
#include "host_delaunay.h"

#include <algorithm>
#include <cstring>

namespace svb {

namespace {

struct Handle {
    int t;  // triangle record
    int o;  // orientation 0..2
};

const int PLUS1[3] = {1, 2, 0};
const int MINUS1[3] = {2, 0, 1};

struct Mesh {
    const int32_t *X;  // valid for index -1 (sentinel) .. n-1
    const int32_t *Y;
    int32_t *nbr;  // 3 per record, encoded handle (t << 2 | o)
    int32_t *vtx;  // 3 per record, vertex id or -1 ("NULL")
    int ntri;

    static int enc(Handle h) { return (h.t << 2) | h.o; }
    static Handle dec(int e) { return Handle{e >> 2, e & 3}; }

    int org(Handle h) const { return vtx[3 * h.t + PLUS1[h.o]]; }
    int dest(Handle h) const { return vtx[3 * h.t + MINUS1[h.o]]; }
    int apex(Handle h) const { return vtx[3 * h.t + h.o]; }
    void setorg(Handle h, int v) { vtx[3 * h.t + PLUS1[h.o]] = v; }
    void setdest(Handle h, int v) { vtx[3 * h.t + MINUS1[h.o]] = v; }
    void setapex(Handle h, int v) { vtx[3 * h.t + h.o] = v; }
    Handle sym(Handle h) const { return dec(nbr[3 * h.t + h.o]); }
    static Handle lnext(Handle h) { return Handle{h.t, PLUS1[h.o]}; }
    static Handle lprev(Handle h) { return Handle{h.t, MINUS1[h.o]}; }
    void bond(Handle a, Handle b) {
        nbr[3 * a.t + a.o] = enc(b);
        nbr[3 * b.t + b.o] = enc(a);
    }
    // record 0 is the "outer space" record: its neighbours are itself and its vertices are NULL
    Handle make() {
        const int t = ntri++;
        nbr[3 * t] = nbr[3 * t + 1] = nbr[3 * t + 2] = 0;
        vtx[3 * t] = vtx[3 * t + 1] = vtx[3 * t + 2] = -1;
        return Handle{t, 0};
    }

    // exact orientation: > 0 iff a, b, c are counter-clockwise
    int64_t ccw(int a, int b, int c) const {
        return (int64_t)(X[a] - X[c]) * (Y[b] - Y[c]) - (int64_t)(Y[a] - Y[c]) * (X[b] - X[c]);
    }
    // exact in-circle: > 0 iff d lies inside the circle through a, b, c (a, b, c counter-clockwise)
    int64_t incircle(int a, int b, int c, int d) const {
        const int64_t adx = X[a] - X[d], ady = Y[a] - Y[d];
        const int64_t bdx = X[b] - X[d], bdy = Y[b] - Y[d];
        const int64_t cdx = X[c] - X[d], cdy = Y[c] - Y[d];
        const int64_t alift = adx * adx + ady * ady;
        const int64_t blift = bdx * bdx + bdy * bdy;
        const int64_t clift = cdx * cdx + cdy * cdy;
        return alift * (bdx * cdy - cdx * bdy) + blift * (cdx * ady - adx * cdy) + clift * (adx * bdy - bdx * ady);
    }

    void merge(Handle &farleft, Handle &innerleft, Handle &innerright, Handle &farright, int axis);
    void recurse(const int32_t *sorted, int count, int axis, Handle &farleft, Handle &farright);
};

// Knit two adjacent triangulations together (triangle.cpp:5362-5651).
void Mesh::merge(Handle &farleft, Handle &innerleft, Handle &innerright, Handle &farright, int axis) {
    int innerleftdest = dest(innerleft), innerleftapex = apex(innerleft);
    int innerrightorg = org(innerright), innerrightapex = apex(innerright);
    if (axis == 1) {
        // horizontal cut: move the extreme handles from leftmost/rightmost to bottommost/topmost vertices
        int farleftpt = org(farleft), farleftapex = apex(farleft);
        int farrightpt = dest(farright), farrightapex = apex(farright);
        while (Y[farleftapex] < Y[farleftpt]) {
            farleft = sym(lnext(farleft));
            farleftpt = farleftapex;
            farleftapex = apex(farleft);
        }
        Handle check = sym(innerleft);
        int checkv = apex(check);
        while (Y[checkv] > Y[innerleftdest]) {
            innerleft = lnext(check);
            innerleftapex = innerleftdest;
            innerleftdest = checkv;
            check = sym(innerleft);
            checkv = apex(check);
        }
        while (Y[innerrightapex] < Y[innerrightorg]) {
            innerright = sym(lnext(innerright));
            innerrightorg = innerrightapex;
            innerrightapex = apex(innerright);
        }
        check = sym(farright);
        checkv = apex(check);
        while (Y[checkv] > Y[farrightpt]) {
            farright = lnext(check);
            farrightapex = farrightpt;
            farrightpt = checkv;
            check = sym(farright);
            checkv = apex(check);
        }
        (void)farrightapex;
    }
    // lower common tangent
    bool changed;
    do {
        changed = false;
        if (ccw(innerleftdest, innerleftapex, innerrightorg) > 0) {
            innerleft = sym(lprev(innerleft));
            innerleftdest = innerleftapex;
            innerleftapex = apex(innerleft);
            changed = true;
        }
        if (ccw(innerrightapex, innerrightorg, innerleftdest) > 0) {
            innerright = sym(lnext(innerright));
            innerrightorg = innerrightapex;
            innerrightapex = apex(innerright);
            changed = true;
        }
    } while (changed);

    Handle leftcand = sym(innerleft);
    Handle rightcand = sym(innerright);
    // bottom bounding record
    Handle base = make();
    bond(base, innerleft);
    base = lnext(base);
    bond(base, innerright);
    base = lnext(base);
    setorg(base, innerrightorg);
    setdest(base, innerleftdest);
    if (innerleftdest == org(farleft)) farleft = lnext(base);
    if (innerrightorg == dest(farright)) farright = lprev(base);

    int lowerleft = innerleftdest, lowerright = innerrightorg;
    int upperleft = apex(leftcand), upperright = apex(rightcand);
    while (true) {
        const bool leftfinished = ccw(upperleft, lowerleft, lowerright) <= 0;
        const bool rightfinished = ccw(upperright, lowerleft, lowerright) <= 0;
        if (leftfinished && rightfinished) {
            // top bounding record
            Handle top = make();
            setorg(top, lowerleft);
            setdest(top, lowerright);
            bond(top, base);
            top = lnext(top);
            bond(top, rightcand);
            top = lnext(top);
            bond(top, leftcand);
            if (axis == 1) {
                // restore the extreme handles to the leftmost / rightmost vertices
                int farleftpt = org(farleft), farleftapex = apex(farleft);
                int farrightpt = dest(farright), farrightapex = apex(farright);
                Handle check = sym(farleft);
                int checkv = apex(check);
                while (X[checkv] < X[farleftpt]) {
                    farleft = lprev(check);
                    farleftapex = farleftpt;
                    farleftpt = checkv;
                    check = sym(farleft);
                    checkv = apex(check);
                }
                (void)farleftapex;
                while (X[farrightapex] > X[farrightpt]) {
                    farright = sym(lprev(farright));
                    farrightpt = farrightapex;
                    farrightapex = apex(farright);
                }
            }
            return;
        }
        if (!leftfinished) {
            // would deleting the left candidate edge expose a vertex that violates the Delaunay property?
            Handle next = sym(lprev(leftcand));
            int nextapex = apex(next);
            if (nextapex >= 0) {
                bool bad = incircle(lowerleft, lowerright, upperleft, nextapex) > 0;
                while (bad) {
                    // edge flip: the left triangulation gains one bounding record
                    next = lnext(next);
                    const Handle topcasing = sym(next);
                    next = lnext(next);
                    const Handle sidecasing = sym(next);
                    bond(next, topcasing);
                    bond(leftcand, sidecasing);
                    leftcand = lnext(leftcand);
                    const Handle outercasing = sym(leftcand);
                    next = lprev(next);
                    bond(next, outercasing);
                    setorg(leftcand, lowerleft);
                    setdest(leftcand, -1);
                    setapex(leftcand, nextapex);
                    setorg(next, -1);
                    setdest(next, upperleft);
                    setapex(next, nextapex);
                    upperleft = nextapex;
                    next = sidecasing;
                    nextapex = apex(next);
                    bad = nextapex >= 0 ? incircle(lowerleft, lowerright, upperleft, nextapex) > 0 : false;
                }
            }
        }
        if (!rightfinished) {
            Handle next = sym(lnext(rightcand));
            int nextapex = apex(next);
            if (nextapex >= 0) {
                bool bad = incircle(lowerleft, lowerright, upperright, nextapex) > 0;
                while (bad) {
                    next = lprev(next);
                    const Handle topcasing = sym(next);
                    next = lprev(next);
                    const Handle sidecasing = sym(next);
                    bond(next, topcasing);
                    bond(rightcand, sidecasing);
                    rightcand = lprev(rightcand);
                    const Handle outercasing = sym(rightcand);
                    next = lnext(next);
                    bond(next, outercasing);
                    setorg(rightcand, -1);
                    setdest(rightcand, lowerright);
                    setapex(rightcand, nextapex);
                    setorg(next, upperright);
                    setdest(next, -1);
                    setapex(next, nextapex);
                    upperright = nextapex;
                    next = sidecasing;
                    nextapex = apex(next);
                    bad = nextapex >= 0 ? incircle(lowerleft, lowerright, upperright, nextapex) > 0 : false;
                }
            }
        }
        if (leftfinished || (!rightfinished && incircle(upperleft, lowerleft, lowerright, upperright) > 0)) {
            // new edge lowerleft -- upperright
            bond(base, rightcand);
            base = lprev(rightcand);
            setdest(base, lowerleft);
            lowerright = upperright;
            rightcand = sym(base);
            upperright = apex(rightcand);
        } else {
            // new edge upperleft -- lowerright (also taken on a co-circular tie)
            bond(base, leftcand);
            base = lnext(leftcand);
            setorg(base, lowerright);
            lowerleft = upperleft;
            leftcand = sym(base);
            upperleft = apex(leftcand);
        }
    }
}

// triangle.cpp:5670-5815
void Mesh::recurse(const int32_t *s, int count, int axis, Handle &farleft, Handle &farright) {
    if (count == 2) {
        // an edge: two bounding records glued along all three sides
        farleft = make();
        setorg(farleft, s[0]);
        setdest(farleft, s[1]);
        farright = make();
        setorg(farright, s[1]);
        setdest(farright, s[0]);
        bond(farleft, farright);
        farleft = lprev(farleft);
        farright = lnext(farright);
        bond(farleft, farright);
        farleft = lprev(farleft);
        farright = lnext(farright);
        bond(farleft, farright);
        farleft = lprev(farright);  // origin of farleft = s[0]
        return;
    }
    if (count == 3) {
        Handle mid = make(), t1 = make(), t2 = make(), t3 = make();
        const int64_t area = ccw(s[0], s[1], s[2]);
        if (area == 0) {
            // collinear: two edges, four bounding records
            setorg(mid, s[0]);
            setdest(mid, s[1]);
            setorg(t1, s[1]);
            setdest(t1, s[0]);
            setorg(t2, s[2]);
            setdest(t2, s[1]);
            setorg(t3, s[1]);
            setdest(t3, s[2]);
            bond(mid, t1);
            bond(t2, t3);
            mid = lnext(mid);
            t1 = lprev(t1);
            t2 = lnext(t2);
            t3 = lprev(t3);
            bond(mid, t3);
            bond(t1, t2);
            mid = lnext(mid);
            t1 = lprev(t1);
            t2 = lnext(t2);
            t3 = lprev(t3);
            bond(mid, t1);
            bond(t2, t3);
            farleft = t1;
            farright = t2;
        } else {
            // one real triangle (mid) surrounded by three bounding records
            setorg(mid, s[0]);
            setdest(t1, s[0]);
            setorg(t3, s[0]);
            if (area > 0) {
                setdest(mid, s[1]);
                setorg(t1, s[1]);
                setdest(t2, s[1]);
                setapex(mid, s[2]);
                setorg(t2, s[2]);
                setdest(t3, s[2]);
            } else {
                setdest(mid, s[2]);
                setorg(t1, s[2]);
                setdest(t2, s[2]);
                setapex(mid, s[1]);
                setorg(t2, s[1]);
                setdest(t3, s[1]);
            }
            bond(mid, t1);
            mid = lnext(mid);
            bond(mid, t2);
            mid = lnext(mid);
            bond(mid, t3);
            t1 = lprev(t1);
            t2 = lnext(t2);
            bond(t1, t2);
            t1 = lprev(t1);
            t3 = lprev(t3);
            bond(t1, t3);
            t2 = lnext(t2);
            t3 = lprev(t3);
            bond(t2, t3);
            farleft = t1;
            farright = area > 0 ? t2 : lnext(farleft);
        }
        return;
    }
    const int divider = count >> 1;
    Handle innerleft, innerright;
    recurse(s, divider, 1 - axis, farleft, innerleft);
    recurse(s + divider, count - divider, 1 - axis, innerright, farright);
    merge(farleft, innerleft, innerright, farright, axis);
}

struct Lcg {  // triangle.cpp:3833-3836, seeded with 1 by triangleinit (:3818)
    unsigned long seed = 1;
    unsigned long next(unsigned int choices) {
        seed = (seed * 1366ul + 150889ul) % 714025ul;
        return seed / (714025ul / choices + 1);
    }
};

struct LexXY {
    const int32_t *X, *Y;
    bool less(int a, int px, int py) const { return X[a] < px || (X[a] == px && Y[a] < py); }
    bool greater(int a, int px, int py) const { return X[a] > px || (X[a] == px && Y[a] > py); }
};

// Randomised Hoare quicksort, re-stated so that equal-coordinate vertices end up in the reference's order.
void lex_quicksort(int32_t *a, int n, const LexXY &c, Lcg &rng) {
    if (n == 2) {
        if (c.greater(a[0], c.X[a[1]], c.Y[a[1]])) std::swap(a[0], a[1]);
        return;
    }
    const int pivot = (int)rng.next((unsigned)n);
    const int px = c.X[a[pivot]], py = c.Y[a[pivot]];
    int left = -1, right = n;
    while (left < right) {
        do {
            left++;
        } while (left <= right && c.less(a[left], px, py));
        do {
            right--;
        } while (left <= right && c.greater(a[right], px, py));
        if (left < right) std::swap(a[left], a[right]);
    }
    if (left > 1) lex_quicksort(a, left, c, rng);
    if (right < n - 2) lex_quicksort(a + right + 1, n - right - 1, c, rng);
}

// Alternating-axis partition (triangle.cpp:5307-5325).  Only the SETS on either side of each median matter
// (keys are distinct after duplicate removal), so a deterministic selection replaces the randomised one.
void alternate_axes(int32_t *a, int n, int axis, const int32_t *X, const int32_t *Y) {
    const int divider = n >> 1;
    if (n <= 3) axis = 0;
    const int32_t *K0 = axis ? Y : X, *K1 = axis ? X : Y;
    std::nth_element(a, a + divider, a + n, [K0, K1](int p, int q) { return K0[p] < K0[q] || (K0[p] == K0[q] && K1[p] < K1[q]); });
    if (n - divider >= 2) {
        if (divider >= 2) alternate_axes(a, divider, 1 - axis, X, Y);
        alternate_axes(a + divider, n - divider, 1 - axis, X, Y);
    }
}

}  // namespace

int delaunay_xy(const int32_t *x, const int32_t *y, int n, int32_t *tri_out, int cap, DelaunayScratch &scratch) {
    if (n < 3) return 0;
    // arena: [X sentinel + n][Y sentinel + n][sorted n][nbr 3*R][vtx 3*R],  R <= 1 + 2 records per vertex + merges
    const size_t max_records = 1 + (size_t)4 * n + 16;
    const size_t need = 2 * (size_t)(n + 1) + n + 6 * max_records;
    if (scratch.storage.size() < need) scratch.storage.resize(need);
    int32_t *base = scratch.storage.data();
    int32_t *X = base + 1, *Y = X + n + 1, *sorted = Y + n, *nbr = sorted + n, *vtx = nbr + 3 * max_records;
    X[-1] = 0;
    Y[-1] = 0;
    std::memcpy(X, x, sizeof(int32_t) * n);
    std::memcpy(Y, y, sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) sorted[i] = i;

    LexXY cmp{X, Y};
    Lcg rng;
    lex_quicksort(sorted, n, cmp, rng);
    int m = 0;  // drop duplicates, keeping the first of each run (triangle.cpp:5890-5903)
    for (int j = 1; j < n; j++)
        if (!(X[sorted[m]] == X[sorted[j]] && Y[sorted[m]] == Y[sorted[j]])) sorted[++m] = sorted[j];
    m++;
    if (m < 3) return 0;
    {
        const int divider = m >> 1;
        if (m - divider >= 2) {
            if (divider >= 2) alternate_axes(sorted, divider, 1, X, Y);
            alternate_axes(sorted + divider, m - divider, 1, X, Y);
        }
    }

    Mesh mesh;
    mesh.X = X;
    mesh.Y = Y;
    mesh.nbr = nbr;
    mesh.vtx = vtx;
    mesh.ntri = 0;
    mesh.make();  // record 0: outer space
    Handle hullleft, hullright;
    mesh.recurse(sorted, m, 0, hullleft, hullright);

    int count = 0;
    for (int t = 1; t < mesh.ntri; t++) {
        const int a = vtx[3 * t + 1], b = vtx[3 * t + 2], c = vtx[3 * t];  // org, dest, apex at orientation 0
        if (a < 0 || b < 0 || c < 0) continue;                            // bounding record (removeghosts, :5817-5859)
        if (count < cap) {
            tri_out[3 * count] = a;
            tri_out[3 * count + 1] = b;
            tri_out[3 * count + 2] = c;
        }
        count++;
    }
    return count;
}

int delaunay_support(const int32_t *support, int n, int right_image, int32_t *tri_out, int cap, DelaunayScratch &scratch) {
    if (n < 3) return 0;
    // coordinates are staged at the tail of the arena, past everything delaunay_xy lays out for n points
    const size_t max_records = 1 + (size_t)4 * n + 16;
    const size_t need = 2 * (size_t)(n + 1) + n + 6 * max_records;
    if (scratch.storage.size() < need + 2 * (size_t)n) scratch.storage.resize(need + 2 * (size_t)n);
    int32_t *xs = scratch.storage.data() + need, *ys = xs + n;
    for (int i = 0; i < n; i++) {
        const int u = support[3 * i], v = support[3 * i + 1], d = support[3 * i + 2];
        xs[i] = right_image ? u - d : u;  // elas.cpp:451-461
        ys[i] = v;
    }
    return delaunay_xy(xs, ys, n, tri_out, cap, scratch);
}

}  // namespace svb
