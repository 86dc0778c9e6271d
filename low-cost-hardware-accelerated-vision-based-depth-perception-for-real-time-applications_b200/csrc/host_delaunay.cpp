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
// Triangle records are 32-byte rows {3 neighbour handles, 3 vertices} in one flat arena (no pointer pool, no per-call
// malloc); a handle is one int (8 * record + orientation).
//
// What is restated and what is merely equivalent: the vertex sort and the median partitions of the reference are
// randomised, but only two things they produce reach the output -- WHICH of several equal-coordinate vertices
// survives, and the SETS on either side of every median (subsets of <= 3 come out x-sorted).  So the common
// duplicate-free case runs an LSD radix sort and a k-d split over presorted rank lists (two branch-free linear passes
// per level); only an input that does contain duplicates goes through the restated randomised quicksort.
#include "host_delaunay.h"

#include <algorithm>
#include <cstring>

namespace svb {

namespace {

// Record r occupies R[8r .. 8r+7] = {nbr0, nbr1, nbr2, -, vtx0, vtx1, vtx2, -}; a handle is (8r + orientation), i.e.
// the index of its own neighbour slot, and its apex sits four ints further.
struct Pt {
    int32_t x, y;
};

struct Mesh {
    const Pt *__restrict P;  // coordinates by vertex id (= position in the lexicographic order); index -1 is the "NULL" sentinel
    int32_t *__restrict R;
    int ntri;

    static int lnext(int e) { return (e & 3) == 2 ? e - 2 : e + 1; }
    static int lprev(int e) { return (e & 3) == 0 ? e + 2 : e - 1; }
    int &nbr(int e) { return R[e]; }
    int &vtx(int e) { return R[e + 4]; }
    int apex(int e) { return vtx(e); }
    int org(int e) { return vtx(lnext(e)); }
    int dest(int e) { return vtx(lprev(e)); }
    void setapex(int e, int v) { vtx(e) = v; }
    void setorg(int e, int v) { vtx(lnext(e)) = v; }
    void setdest(int e, int v) { vtx(lprev(e)) = v; }
    int sym(int e) { return nbr(e); }
    void bond(int a, int b) {
        nbr(a) = b;
        nbr(b) = a;
    }
    // record 0 is the "outer space" record: its neighbours are itself and its vertices are NULL
    int make() {
        const int t = ntri++;
        int32_t *r = R + 8 * t;
        r[0] = r[1] = r[2] = r[3] = 0;
        r[4] = r[5] = r[6] = r[7] = -1;
        return t << 3;
    }

    // exact orientation: > 0 iff a, b, c are counter-clockwise.  |coordinate differences| < 2^14: 32-bit exact.
    static int32_t ccw(Pt a, Pt b, Pt c) { return (a.x - c.x) * (b.y - c.y) - (a.y - c.y) * (b.x - c.x); }
    int32_t ccw(int a, int b, int c) const { return ccw(P[a], P[b], P[c]); }
    // exact in-circle: > 0 iff d lies inside the circle through a, b, c (a, b, c counter-clockwise).
    // lifts and 2x2 minors stay below 2^29 (32-bit), their products below 2^58 (64-bit).
    static int64_t incircle(Pt a, Pt b, Pt c, Pt d) {
        const int32_t adx = a.x - d.x, ady = a.y - d.y;
        const int32_t bdx = b.x - d.x, bdy = b.y - d.y;
        const int32_t cdx = c.x - d.x, cdy = c.y - d.y;
        const int32_t alift = adx * adx + ady * ady;
        const int32_t blift = bdx * bdx + bdy * bdy;
        const int32_t clift = cdx * cdx + cdy * cdy;
        return (int64_t)alift * (bdx * cdy - cdx * bdy) + (int64_t)blift * (cdx * ady - adx * cdy) +
               (int64_t)clift * (adx * bdy - bdx * ady);
    }
    int64_t incircle(int a, int b, int c, int d) const { return incircle(P[a], P[b], P[c], P[d]); }
    // The same determinant for one circle (a, b, c) and many query points: translated to a, expanded along the query's
    // row; the three cofactors are computed once.  test(d) == incircle(a, b, c, d) exactly (|cofactors| < 2^44, terms < 2^59).
    struct Circle {
        Pt a;
        int64_t bx, by, cx, cy, k0, k1, k2;
        // orientation part only: k2 = 2 x signed area of (a, b, c) = ccw(c, a, b)
        Circle(Pt a_, Pt b, Pt c) : a(a_), bx(b.x - a_.x), by(b.y - a_.y), cx(c.x - a_.x), cy(c.y - a_.y), k0(0), k1(0) { k2 = bx * cy - by * cx; }
        void finish() {  // the two cofactors that need the lifts
            const int64_t bl = bx * bx + by * by, cl = cx * cx + cy * cy;
            k0 = by * cl - bl * cy;
            k1 = bx * cl - bl * cx;
        }
        int64_t test(Pt d) const {
            const int64_t dx = d.x - a.x, dy = d.y - a.y;
            return dy * k1 - dx * k0 - (dx * dx + dy * dy) * k2;  // = -det[b'; c'; d'] = det[a-d; b-d; c-d]
        }
    };

    void merge(int &farleft, int &innerleft, int &innerright, int &farright, int axis);
    void recurse(const int32_t *sorted, int count, int axis, int &farleft, int &farright);
};

// Knit two adjacent triangulations together (triangle.cpp:5362-5651).
void Mesh::merge(int &farleft, int &innerleft, int &innerright, int &farright, int axis) {
    int innerleftdest = dest(innerleft), innerleftapex = apex(innerleft);
    int innerrightorg = org(innerright), innerrightapex = apex(innerright);
    if (axis == 1) {
        // horizontal cut: move the extreme handles from leftmost/rightmost to bottommost/topmost vertices
        int farleftpt = org(farleft), farleftapex = apex(farleft);
        int farrightpt = dest(farright);
        while (P[farleftapex].y < P[farleftpt].y) {
            farleft = sym(lnext(farleft));
            farleftpt = farleftapex;
            farleftapex = apex(farleft);
        }
        int check = sym(innerleft);
        int checkv = apex(check);
        while (P[checkv].y > P[innerleftdest].y) {
            innerleft = lnext(check);
            innerleftapex = innerleftdest;
            innerleftdest = checkv;
            check = sym(innerleft);
            checkv = apex(check);
        }
        while (P[innerrightapex].y < P[innerrightorg].y) {
            innerright = sym(lnext(innerright));
            innerrightorg = innerrightapex;
            innerrightapex = apex(innerright);
        }
        check = sym(farright);
        checkv = apex(check);
        while (P[checkv].y > P[farrightpt].y) {
            farright = lnext(check);
            farrightpt = checkv;
            check = sym(farright);
            checkv = apex(check);
        }
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

    int leftcand = sym(innerleft);
    int rightcand = sym(innerright);
    // bottom bounding record
    int base = make();
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
    Pt pll = P[lowerleft], plr = P[lowerright], pul = P[upperleft], pur = P[upperright];  // coordinates ride along
    while (true) {
        // circles through the base edge and either candidate; their orientation term is the "finished" test
        // (ccw(upper, lowerleft, lowerright) = 2 x area of (lowerleft, lowerright, upper), elas' triangle.cpp:5480-5483)
        Circle cleft(pll, plr, pul), cright(pll, plr, pur);
        const bool leftfinished = cleft.k2 <= 0;
        const bool rightfinished = cright.k2 <= 0;
        if (leftfinished && rightfinished) {
            // top bounding record
            int top = make();
            setorg(top, lowerleft);
            setdest(top, lowerright);
            bond(top, base);
            top = lnext(top);
            bond(top, rightcand);
            top = lnext(top);
            bond(top, leftcand);
            if (axis == 1) {
                // restore the extreme handles to the leftmost / rightmost vertices
                int farleftpt = org(farleft);
                int farrightpt = dest(farright), farrightapex = apex(farright);
                int check = sym(farleft);
                int checkv = apex(check);
                while (P[checkv].x < P[farleftpt].x) {
                    farleft = lprev(check);
                    farleftpt = checkv;
                    check = sym(farleft);
                    checkv = apex(check);
                }
                while (P[farrightapex].x > P[farrightpt].x) {
                    farright = sym(lprev(farright));
                    farrightpt = farrightapex;
                    farrightapex = apex(farright);
                }
            }
            return;
        }
        if (!leftfinished) {
            cleft.finish();  // used by the left deletion test and by the final choice
            // would deleting the left candidate edge expose a vertex that violates the Delaunay property?
            int next = sym(lprev(leftcand));
            int nextapex = apex(next);
            while (nextapex >= 0 && cleft.test(P[nextapex]) > 0) {
                // edge flip: the left triangulation gains one bounding record
                next = lnext(next);
                const int topcasing = sym(next);
                next = lnext(next);
                const int sidecasing = sym(next);
                bond(next, topcasing);
                bond(leftcand, sidecasing);
                leftcand = lnext(leftcand);
                const int outercasing = sym(leftcand);
                next = lprev(next);
                bond(next, outercasing);
                setorg(leftcand, lowerleft);
                setdest(leftcand, -1);
                setapex(leftcand, nextapex);
                setorg(next, -1);
                setdest(next, upperleft);
                setapex(next, nextapex);
                upperleft = nextapex;
                pul = P[nextapex];
                cleft = Circle(pll, plr, pul);
                cleft.finish();
                next = sidecasing;
                nextapex = apex(next);
            }
        }
        if (!rightfinished) {
            int next = sym(lnext(rightcand));
            int nextapex = apex(next);
            cright.finish();
            while (nextapex >= 0 && cright.test(P[nextapex]) > 0) {
                next = lprev(next);
                const int topcasing = sym(next);
                next = lprev(next);
                const int sidecasing = sym(next);
                bond(next, topcasing);
                bond(rightcand, sidecasing);
                rightcand = lprev(rightcand);
                const int outercasing = sym(rightcand);
                next = lnext(next);
                bond(next, outercasing);
                setorg(rightcand, -1);
                setdest(rightcand, lowerright);
                setapex(rightcand, nextapex);
                setorg(next, upperright);
                setdest(next, -1);
                setapex(next, nextapex);
                upperright = nextapex;
                pur = P[nextapex];
                cright = Circle(pll, plr, pur);
                cright.finish();
                next = sidecasing;
                nextapex = apex(next);
            }
        }
        // incircle(pul, pll, plr, pur): the same circle as `cleft` (a cyclic shift of the rows leaves the determinant alone)
        if (leftfinished || (!rightfinished && cleft.test(pur) > 0)) {
            // new edge lowerleft -- upperright
            bond(base, rightcand);
            base = lprev(rightcand);
            setdest(base, lowerleft);
            lowerright = upperright;
            plr = pur;
            rightcand = sym(base);
            upperright = apex(rightcand);
            pur = P[upperright];
        } else {
            // new edge upperleft -- lowerright (also taken on a co-circular tie)
            bond(base, leftcand);
            base = lnext(leftcand);
            setorg(base, lowerright);
            lowerleft = upperleft;
            pll = pul;
            leftcand = sym(base);
            upperleft = apex(leftcand);
            pul = P[upperleft];
        }
    }
}

// triangle.cpp:5670-5815
void Mesh::recurse(const int32_t *s, int count, int axis, int &farleft, int &farright) {
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
        int mid = make(), t1 = make(), t2 = make(), t3 = make();
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
    int innerleft, innerright;
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

// Lexicographic order of duplicate-free input: LSD radix sort, 3 passes of 9 bits over key = (x + 8192) << 13 | y.
// item = key << 32 | index.  Returns false (nothing usable in `a`) when a coordinate is out of range.
const int RADIX_BITS = 9, RADIX = 1 << RADIX_BITS;
bool radix_lex_sort(const int32_t *x, const int32_t *y, int n, uint64_t *a, uint64_t *b) {
    uint32_t hist[3][RADIX];
    std::memset(hist, 0, sizeof(hist));
    uint32_t bad = 0;
    for (int i = 0; i < n; i++) {
        const uint32_t xb = (uint32_t)(x[i] + 8192), yb = (uint32_t)y[i];
        bad |= (xb >> 14) | (yb >> 13);
        const uint32_t key = xb << 13 | yb;
        a[i] = (uint64_t)key << 32 | (uint32_t)i;
        hist[0][key & (RADIX - 1)]++;
        hist[1][(key >> RADIX_BITS) & (RADIX - 1)]++;
        hist[2][(key >> (2 * RADIX_BITS)) & (RADIX - 1)]++;
    }
    if (bad) return false;
    for (int p = 0; p < 3; p++) {
        uint32_t sum = 0;
        for (int k = 0; k < RADIX; k++) {
            const uint32_t c = hist[p][k];
            hist[p][k] = sum;
            sum += c;
        }
    }
    for (int p = 0; p < 3; p++) {
        const int shift = 32 + p * RADIX_BITS;
        uint32_t *h = hist[p];
        for (int i = 0; i < n; i++) b[h[(a[i] >> shift) & (RADIX - 1)]++] = a[i];
        std::swap(a, b);
    }
    return true;  // three passes: the result is in the array that was passed as `b`
}

// Alternating-axis median partition (triangle.cpp:5243-5325) over rank pairs e = xrank << HALF | yrank (32-bit
// elements while the ranks fit 16 bits, 64-bit ones otherwise).  Lx holds the subset in x order, Ly the same subset in
// y order; splitting along one axis is a cut of that axis' list and a stable (order-preserving) split of the other
// one.  Subsets of <= 3 are emitted in x order.
template <typename E, int HALF>
void kd_partition(E *Lx, E *Ly, E *tmp, int n, int axis, int32_t *out) {
    const E LOW = ((E)1 << HALF) - 1;
    if (n <= 3) {
        for (int i = 0; i < n; i++) out[i] = (int32_t)(Lx[i] >> HALF);
        return;
    }
    const int divider = n >> 1;
    int lo = 0, hi = 0;
    if (axis == 0) {
        const E pivot = Lx[divider] >> HALF;
        for (int i = 0; i < n; i++) {
            const E e = Ly[i];
            const int low = (e >> HALF) < pivot;
            Ly[lo] = e;  // lo <= i: never overtakes the read position
            tmp[hi] = e;
            lo += low;
            hi += 1 - low;
        }
        for (int i = 0; i < hi; i++) Ly[divider + i] = tmp[i];
    } else {
        const E pivot = Ly[divider] & LOW;
        for (int i = 0; i < n; i++) {
            const E e = Lx[i];
            const int low = (e & LOW) < pivot;
            Lx[lo] = e;
            tmp[hi] = e;
            lo += low;
            hi += 1 - low;
        }
        for (int i = 0; i < hi; i++) Lx[divider + i] = tmp[i];
    }
    kd_partition<E, HALF>(Lx, Ly, tmp, divider, 1 - axis, out);
    kd_partition<E, HALF>(Lx + divider, Ly + divider, tmp, n - divider, 1 - axis, out + divider);
}

template <typename E, int HALF>
void kd_order(const int32_t *by_y, int m, void *arena, int32_t *out) {
    E *Lx = (E *)arena, *Ly = Lx + m, *tmp = Ly + m;
    for (int r = 0; r < m; r++) Lx[by_y[r]] = (E)(uint32_t)by_y[r] << HALF | (E)(uint32_t)r;
    for (int r = 0; r < m; r++) Ly[r] = Lx[by_y[r]];
    kd_partition<E, HALF>(Lx, Ly, tmp, m, 0, out);
}

}  // namespace

int delaunay_xy(const int32_t *x, const int32_t *y, int n, int32_t *tri_out, int cap, DelaunayScratch &scratch) {
    if (n < 3) return 0;
    // arena (int32 units): [5 n uint64: sort ping-pong / Lx, Ly, tmp][{x,y} sentinel + n][vid n][sorted n]
    //                      [records 8 R],  R <= 1 + 2 records per vertex + merges
    const size_t max_records = 1 + (size_t)4 * n + 16;
    const size_t need = 10 * (size_t)n + 2 * (size_t)(n + 1) + 2 * (size_t)n + 8 * max_records + 2;
    if (scratch.storage.size() < need) scratch.storage.resize(need);
    int32_t *base = scratch.storage.data();
    base += ((uintptr_t)base & 7) ? 1 : 0;  // 8-byte alignment for the uint64 arrays
    uint64_t *A = (uint64_t *)base, *B = A + n;  // radix buffers, then the rank lists of step 3
    Pt *P = (Pt *)(A + 5 * (size_t)n) + 1;
    int32_t *vid = (int32_t *)(P + n), *sorted = vid + n, *R = sorted + n;
    P[-1] = Pt{0, 0};

    // 1. lexicographic order, duplicates dropped (keeping the first of each run, triangle.cpp:5890-5903)
    int m = 0;
    bool fast = radix_lex_sort(x, y, n, A, B);
    if (fast) {
        const uint64_t *S = B;  // see radix_lex_sort
        for (int i = 1; i < n; i++)
            if ((S[i] >> 32) == (S[i - 1] >> 32)) {
                fast = false;  // which duplicate survives depends on the reference's sort: take the restated path
                break;
            }
        if (fast) {
            for (int i = 0; i < n; i++) vid[i] = (int32_t)(uint32_t)S[i];
            m = n;
        }
    }
    if (!fast) {
        for (int i = 0; i < n; i++) vid[i] = i;
        LexXY cmp{x, y};
        Lcg rng;
        lex_quicksort(vid, n, cmp, rng);
        for (int j = 1; j < n; j++)
            if (!(x[vid[m]] == x[vid[j]] && y[vid[m]] == y[vid[j]])) vid[++m] = vid[j];
        m++;
    }
    if (m < 3) return 0;
    int ymin = y[vid[0]], ymax = ymin;
    for (int i = 0; i < m; i++) {
        P[i] = Pt{x[vid[i]], y[vid[i]]};
        ymin = std::min(ymin, P[i].y);
        ymax = std::max(ymax, P[i].y);
    }

    // 2. rank of every vertex in (y, x) order: a stable sort by y of the x-ordered list
    int32_t *by_y = (int32_t *)(A + 3 * (size_t)n);  // the last two of the five uint64 arrays
    if ((int64_t)ymax - ymin + 2 <= (int64_t)(8 * max_records)) {  // the counting array borrows the record arena
        const int span = ymax - ymin + 1;
        int32_t *cnt = R;  // free until step 4
        std::memset(cnt, 0, sizeof(int32_t) * (span + 1));
        for (int i = 0; i < m; i++) cnt[P[i].y - ymin + 1]++;
        for (int k = 0; k < span; k++) cnt[k + 1] += cnt[k];
        for (int i = 0; i < m; i++) {
            const int pos = cnt[P[i].y - ymin]++;
            by_y[pos] = i;
        }
    } else {
        for (int i = 0; i < m; i++) by_y[i] = i;
        std::stable_sort(by_y, by_y + m, [P](int p, int q) { return P[p].y < P[q].y; });
    }

    // 3. alternating cuts; `sorted` receives the vertex ids in recursion order
    if (m <= 65535)
        kd_order<uint32_t, 16>(by_y, m, A, sorted);
    else
        kd_order<uint64_t, 32>(by_y, m, A, sorted);

    // 4. divide and conquer
    Mesh mesh;
    mesh.P = P;
    mesh.R = R;
    mesh.ntri = 0;
    mesh.make();  // record 0: outer space
    int hullleft, hullright;
    mesh.recurse(sorted, m, 0, hullleft, hullright);

    int count = 0;
    for (int t = 1; t < mesh.ntri; t++) {
        const int32_t *r = R + 8 * t + 4;
        const int a = r[1], b = r[2], c = r[0];  // org, dest, apex at orientation 0
        if ((a | b | c) < 0) continue;           // bounding record (removeghosts, :5817-5859)
        if (count < cap) {
            tri_out[3 * count] = vid[a];
            tri_out[3 * count + 1] = vid[b];
            tri_out[3 * count + 2] = vid[c];
        }
        count++;
    }
    return count;
}

int delaunay_support_ordered(const int32_t *support, int n, int right_image, const int32_t *order, int32_t *tri_out, int cap,
                             DelaunayScratch &scratch) {
    if (n < 3) return 0;
    // arena (int32 units): [{x,y} sentinel + n][identity sequence n][records 8 R]
    const size_t max_records = 1 + (size_t)4 * n + 16;
    const size_t need = 2 * (size_t)(n + 1) + (size_t)n + 8 * max_records + 2;
    if (scratch.storage.size() < need) scratch.storage.resize(need);
    int32_t *base = scratch.storage.data();
    base += ((uintptr_t)base & 7) ? 1 : 0;
    Pt *P = (Pt *)base + 1;
    int32_t *seq = (int32_t *)(P + n), *R = seq + n;
    P[-1] = Pt{0, 0};
    for (int i = 0; i < n; i++) {
        const int id = order[i];
        if ((unsigned)id >= (unsigned)n) return -1;
        const int32_t *sp = support + 3 * id;
        P[i] = Pt{right_image ? sp[0] - sp[2] : sp[0], sp[1]};  // elas.cpp:451-461
        seq[i] = i;
    }
    Mesh mesh;
    mesh.P = P;
    mesh.R = R;
    mesh.ntri = 0;
    mesh.make();  // record 0: outer space
    int hullleft, hullright;
    mesh.recurse(seq, n, 0, hullleft, hullright);
    int count = 0;
    for (int t = 1; t < mesh.ntri; t++) {
        const int32_t *r = R + 8 * t + 4;
        const int a = r[1], b = r[2], c = r[0];
        if ((a | b | c) < 0) continue;
        if (count < cap) {
            tri_out[3 * count] = order[a];
            tri_out[3 * count + 1] = order[b];
            tri_out[3 * count + 2] = order[c];
        }
        count++;
    }
    return count;
}

int delaunay_support(const int32_t *support, int n, int right_image, int32_t *tri_out, int cap, DelaunayScratch &scratch) {
    if (n < 3) return 0;
    // coordinates are staged at the tail of the arena, past everything delaunay_xy lays out for n points
    const size_t max_records = 1 + (size_t)4 * n + 16;
    const size_t need = 10 * (size_t)n + 2 * (size_t)(n + 1) + 2 * (size_t)n + 8 * max_records + 2;
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
