// The divide-and-conquer Delaunay mesh shared by the host stage (host_delaunay.cpp) and the device kernel that runs the lower merge
// levels (k_delaunay.cu): ONE source for the decision rules, so both produce the same records in the same places.
//
// Decision rules of the reference's triangulator (Triangle 1.6 as vendored in src/common_includes/elas/triangle.cpp, switches "zQB",
// dwyer = 1) that this code follows -- see host_delaunay.cpp for the whole list:
//   - recursion with ghost ("bounding") triangles, 2- and 3-vertex base cases                (triangle.cpp:5670-5815)
//   - hull merge: strict ccw > 0 for the lower tangent, strict incircle > 0 for edge
//     deletion and for choosing the right candidate, horizontal-cut handle rotation          (:5362-5651)
// Every predicate is evaluated exactly in integers: x in [-8192, 16383], y in [0, 8191] keep orient2d below 2^29 and incircle below 2^60.
//
// Records.  Record r occupies R[8r .. 8r+7] = {nbr0, nbr1, nbr2, *, vtx0, vtx1, vtx2, *}; a handle is (8r + orientation), i.e. the
// index of its own neighbour slot, and its apex sits four elements further.  RT = int32_t on the host, uint16_t in the device's shared
// memory (handles < 65536, i.e. <= 4096 vertices; vertex ids are read back sign-extended so that -1 stays the NULL vertex).
// Records are never freed (edge flips rewrite records in place), so the record index of everything is known in advance: a subtree of
// c >= 2 vertices creates exactly 2c - 2 records (2 per 2-vertex leaf, 4 per 3-vertex leaf, 2 per merge), in depth-first order.  A
// node whose records start at b therefore owns [b, b + 2c - 2); its children start at b and b + 2*(c/2) - 2, its own merge creates
// b + 2c - 4 (bottom bounding record) and b + 2c - 3 (top).  That is what lets independent subtrees be built by different device
// threads straight into their final places.  The two spare slots of a node's LAST record (b + 2c - 3) carry the node's result
// handles (farleft in slot 3, farright in slot 7) from the level that built the node to the level that merges it.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif

namespace svb {

struct Pt {
    int32_t x, y;
};

// One node of the recursion tree: the vertices [first, first + count) (consecutive ids in recursion order), the axis of ITS merge
// (0: vertical cut, 1: horizontal cut) and the first of its 2 count - 2 records.
struct DelaunayNode {
    int first, count, axis, b;
    bool exists;
};

// Node k (0 .. 2^depth - 1, left to right) at `depth` of the tree over n vertices; exists = false if an ancestor is already a leaf.
SVB_HD DelaunayNode delaunay_node_at(int n, int depth, int k) {
    DelaunayNode nd = {0, n, 0, 1, true};
    for (int level = depth - 1; level >= 0; level--) {
        if (nd.count <= 3) {
            nd.exists = false;
            return nd;
        }
        const int div = nd.count >> 1;
        if ((k >> level) & 1) {
            nd.first += div;
            nd.b += 2 * div - 2;
            nd.count -= div;
        } else {
            nd.count = div;
        }
        nd.axis ^= 1;
    }
    return nd;
}

// depth of the deepest level that holds a node (all of its nodes are leaves)
SVB_HD int delaunay_max_depth(int n) {
    int c = n, depth = 0;
    while (c > 3) {
        c = (c + 1) >> 1;  // the larger child
        depth++;
    }
    return depth;
}

template <typename RT>
struct MeshT {
    const Pt *__restrict P;  // coordinates by vertex id (= position in the lexicographic order); index -1 is the "NULL" sentinel
    RT *__restrict R;
    int ntri;

    SVB_HD static int lnext(int e) { return (e & 3) == 2 ? e - 2 : e + 1; }
    SVB_HD static int lprev(int e) { return (e & 3) == 0 ? e + 2 : e - 1; }
    SVB_HD int nbr(int e) const { return (int)R[e]; }                                                    // uint16_t: zero-extended
    SVB_HD int vtx(int e) const { return sizeof(RT) == 2 ? (int)(int16_t)R[e + 4] : (int)R[e + 4]; }  // uint16_t: sign-extended (-1 = NULL)
    SVB_HD void set_vtx(int e, int v) { R[e + 4] = (RT)v; }
    SVB_HD int apex(int e) const { return vtx(e); }
    SVB_HD int org(int e) const { return vtx(lnext(e)); }
    SVB_HD int dest(int e) const { return vtx(lprev(e)); }
    SVB_HD void setapex(int e, int v) { set_vtx(e, v); }
    SVB_HD void setorg(int e, int v) { set_vtx(lnext(e), v); }
    SVB_HD void setdest(int e, int v) { set_vtx(lprev(e), v); }
    SVB_HD int sym(int e) const { return nbr(e); }
    SVB_HD void bond(int a, int b) {
        R[a] = (RT)b;
        R[b] = (RT)a;
    }
    // record 0 is the "outer space" record: its neighbours are itself and its vertices are NULL
    SVB_HD int make() {
        const int t = ntri++;
        RT *r = R + 8 * t;
        r[0] = r[1] = r[2] = r[3] = (RT)0;
        r[4] = r[5] = r[6] = r[7] = (RT)-1;
        return t << 3;
    }

    // exact orientation: > 0 iff a, b, c are counter-clockwise.  |x differences| < 2^15, |y differences| < 2^13: 32-bit exact.
    SVB_HD static int32_t ccw(Pt a, Pt b, Pt c) { return (a.x - c.x) * (b.y - c.y) - (a.y - c.y) * (b.x - c.x); }
    SVB_HD int32_t ccw(int a, int b, int c) const { return ccw(P[a], P[b], P[c]); }
    // exact in-circle: > 0 iff d lies inside the circle through a, b, c (a, b, c counter-clockwise).
    // lifts stay below 2^31 and 2x2 minors below 2^29 (32-bit), their products below 2^60 (64-bit).
    SVB_HD static int64_t incircle(Pt a, Pt b, Pt c, Pt d) {
        const int32_t adx = a.x - d.x, ady = a.y - d.y;
        const int32_t bdx = b.x - d.x, bdy = b.y - d.y;
        const int32_t cdx = c.x - d.x, cdy = c.y - d.y;
        const int32_t alift = adx * adx + ady * ady;
        const int32_t blift = bdx * bdx + bdy * bdy;
        const int32_t clift = cdx * cdx + cdy * cdy;
        return (int64_t)alift * (bdx * cdy - cdx * bdy) + (int64_t)blift * (cdx * ady - adx * cdy) +
               (int64_t)clift * (adx * bdy - bdx * ady);
    }
    SVB_HD int64_t incircle(int a, int b, int c, int d) const { return incircle(P[a], P[b], P[c], P[d]); }
    // The same determinant for one circle (a, b, c) and many query points: translated to a, expanded along the query's
    // row; the three cofactors are computed once.  test(d) == incircle(a, b, c, d) exactly (|cofactors| < 2^45, terms < 2^58).
    struct Circle {
        Pt a;
        int64_t bx, by, cx, cy, k0, k1, k2;
        // orientation part only: k2 = 2 x signed area of (a, b, c) = ccw(c, a, b)
        SVB_HD Circle(Pt a_, Pt b, Pt c) : a(a_), bx(b.x - a_.x), by(b.y - a_.y), cx(c.x - a_.x), cy(c.y - a_.y), k0(0), k1(0) { k2 = bx * cy - by * cx; }
        SVB_HD void finish() {  // the two cofactors that need the lifts
            const int64_t bl = bx * bx + by * by, cl = cx * cx + cy * cy;
            k0 = by * cl - bl * cy;
            k1 = bx * cl - bl * cx;
        }
        SVB_HD int64_t test(Pt d) const {
            const int64_t dx = d.x - a.x, dy = d.y - a.y;
            return dy * k1 - dx * k0 - (dx * dx + dy * dy) * k2;  // = -det[b'; c'; d'] = det[a-d; b-d; c-d]
        }
    };

    // node bookkeeping (see the header comment): result handles of a finished node ride in the spare slots of its last record
    SVB_HD static int node_records(int count) { return 2 * count - 2; }
    SVB_HD void store_node_result(int b, int count, int farleft, int farright) {
        RT *r = R + 8 * (b + node_records(count) - 1);
        r[3] = (RT)farleft;
        r[7] = (RT)farright;
    }
    SVB_HD void load_node_result(int b, int count, int &farleft, int &farright) const {
        const RT *r = R + 8 * (b + node_records(count) - 1);
        farleft = (int)r[3];
        farright = (int)r[7];
        if (sizeof(RT) == 2) {
            farleft &= 0xFFFF;
            farright &= 0xFFFF;
        }
    }

    // Level-synchronous form of the recursion: builds node `nd`, whose children (if any) are finished, and stores its result handles.
    // All nodes of one level are independent of each other -- one device thread each (k_delaunay.cu).
    SVB_HD void build_node(const DelaunayNode &nd) {
        int farleft, farright;
        if (nd.count <= 3) {
            ntri = nd.b;
            leaf(nd.first, nd.count, farleft, farright);
        } else {
            const int div = nd.count >> 1;
            int innerleft, innerright;
            load_node_result(nd.b, div, farleft, innerleft);
            load_node_result(nd.b + node_records(div), nd.count - div, innerright, farright);
            ntri = nd.b + node_records(nd.count) - 2;
            merge(farleft, innerleft, innerright, farright, nd.axis);
        }
        store_node_result(nd.b, nd.count, farleft, farright);
    }

    SVB_HD void merge(int &farleft, int &innerleft, int &innerright, int &farright, int axis);
    // base cases: the vertices first .. first + count - 1 (count = 2 or 3) are consecutive ids in x order
    SVB_HD void leaf(int first, int count, int &farleft, int &farright);
    // host recursion (depth first).  With device_depth >= 0, nodes at that depth (or leaves above it) were built by the device: their
    // records are in place and their result handles are read back instead of recursing.
    void recurse(int first, int count, int axis, int depth, int b, int device_depth, int &farleft, int &farright);
};


// Knit two adjacent triangulations together (triangle.cpp:5362-5651).
template <typename RT>
SVB_HD void MeshT<RT>::merge(int &farleft, int &innerleft, int &innerright, int &farright, int axis) {
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

// triangle.cpp:5670-5815, the 2- and 3-vertex base cases
template <typename RT>
SVB_HD void MeshT<RT>::leaf(int first, int count, int &farleft, int &farright) {
    const int s0 = first, s1 = first + 1, s2 = first + 2;

    if (count == 2) {
        // an edge: two bounding records glued along all three sides
        farleft = make();
        setorg(farleft, s0);
        setdest(farleft, s1);
        farright = make();
        setorg(farright, s1);
        setdest(farright, s0);
        bond(farleft, farright);
        farleft = lprev(farleft);
        farright = lnext(farright);
        bond(farleft, farright);
        farleft = lprev(farleft);
        farright = lnext(farright);
        bond(farleft, farright);
        farleft = lprev(farright);  // origin of farleft = s0
        return;
    }
    if (count == 3) {
        int mid = make(), t1 = make(), t2 = make(), t3 = make();
        const int64_t area = ccw(s0, s1, s2);
        if (area == 0) {
            // collinear: two edges, four bounding records
            setorg(mid, s0);
            setdest(mid, s1);
            setorg(t1, s1);
            setdest(t1, s0);
            setorg(t2, s2);
            setdest(t2, s1);
            setorg(t3, s1);
            setdest(t3, s2);
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
            setorg(mid, s0);
            setdest(t1, s0);
            setorg(t3, s0);
            if (area > 0) {
                setdest(mid, s1);
                setorg(t1, s1);
                setdest(t2, s1);
                setapex(mid, s2);
                setorg(t2, s2);
                setdest(t3, s2);
            } else {
                setdest(mid, s2);
                setorg(t1, s2);
                setdest(t2, s2);
                setapex(mid, s1);
                setorg(t2, s1);
                setdest(t3, s1);
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
}

// triangle.cpp:5670-5815
template <typename RT>
void MeshT<RT>::recurse(int first, int count, int axis, int depth, int b, int device_depth, int &farleft, int &farright) {
    if (count <= 3) {
        if (device_depth >= 0 && depth >= device_depth) {
            load_node_result(b, count, farleft, farright);
            return;
        }
        ntri = b;
        leaf(first, count, farleft, farright);
        return;
    }
    if (depth == device_depth) {
        load_node_result(b, count, farleft, farright);
        return;
    }
    const int divider = count >> 1;
    int innerleft, innerright;
    recurse(first, divider, 1 - axis, depth + 1, b, device_depth, farleft, innerleft);
    recurse(first + divider, count - divider, 1 - axis, depth + 1, b + node_records(divider), device_depth, innerright, farright);
    ntri = b + node_records(count) - 2;
    merge(farleft, innerleft, innerright, farright, axis);
}

}  // namespace svb
