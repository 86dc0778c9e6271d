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
// is evaluated exactly in integers: x in [-8192, 16383] and y in [0, 8191] (what frames of up to 8192 x 8192 pixels and disparities up
// to 4095 produce; the stage entry points reject anything else) keep orient2d below 2^29 (32 bits) and incircle below 2^61 (64 bits).
// Triangle records are 32-byte rows {3 neighbour handles, 3 vertices} in one flat arena (no pointer pool, no per-call
// malloc); a handle is one int (8 * record + orientation).
//
// What is restated and what is merely equivalent: the vertex sort and the median partitions of the reference are
// randomised, but only two things they produce reach the output -- WHICH of several equal-coordinate vertices
// survives, and the SETS on either side of every median (subsets of <= 3 come out x-sorted).  So the common
// duplicate-free case runs an LSD radix sort and a k-d split over presorted rank lists (two branch-free linear passes
// per level); only an input that does contain duplicates goes through the restated randomised quicksort.
#include "host_delaunay.h"

#include "delaunay_mesh.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace svb {

namespace {

using Mesh = MeshT<int32_t>;

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

    // 4. divide and conquer on the vertices renumbered in recursion order (a node's vertices are then consecutive ids)
    Pt *P2 = (Pt *)A + 1;  // the sort / rank arrays are free now: 5 n uint64 >= (n + 1) points + n ids
    int32_t *vid2 = (int32_t *)(P2 + m);
    for (int i = 0; i < m; i++) {
        P2[i] = P[sorted[i]];
        vid2[i] = vid[sorted[i]];
    }
    P2[-1] = Pt{0, 0};
    Mesh mesh;
    mesh.P = P2;
    mesh.R = R;
    mesh.ntri = 0;
    mesh.make();  // record 0: outer space
    int hullleft, hullright;
    int par_depth = -1;
    if (scratch.par_threads > 1 && m >= 2048) {
        // the 2^par_depth subtrees of one depth on threads of their own, each straight into its own record range
        par_depth = 1;
        while (par_depth < 2 && (2 << par_depth) <= scratch.par_threads) par_depth++;  // four subtrees: more threads cost more than they gain (measured)
        std::vector<std::thread> threads;
        for (int k = 0; k < (1 << par_depth); k++) {
            const DelaunayNode nd = delaunay_node_at(m, par_depth, k);
            if (!nd.exists) continue;
            threads.emplace_back([nd, par_depth, P2, R] {
                Mesh sub;
                sub.P = P2;
                sub.R = R;
                sub.ntri = 0;
                int fl, fr;
                sub.recurse(nd.first, nd.count, nd.axis, par_depth, nd.b, -1, fl, fr);
                sub.store_node_result(nd.b, nd.count, fl, fr);
            });
        }
        for (auto &t : threads) t.join();
    }
    mesh.recurse(0, m, 0, 0, 1, par_depth, hullleft, hullright);

    int count = 0;
    for (int t = 1; t < 2 * m - 1; t++) {
        const int32_t *r = R + 8 * t + 4;
        const int a = r[1], b = r[2], c = r[0];  // org, dest, apex at orientation 0
        if ((a | b | c) < 0) continue;           // bounding record (removeghosts, :5817-5859)
        if (count < cap) {
            tri_out[3 * count] = vid2[a];
            tri_out[3 * count + 1] = vid2[b];
            tri_out[3 * count + 2] = vid2[c];
        }
        count++;
    }
    return count;
}

int delaunay_support_ordered(const int32_t *support, int n, int right_image, const int32_t *order, int m, int32_t *tri_out, int cap,
                             DelaunayScratch &scratch) {
    if (m < 3 || m > n) return m < 3 && m >= 0 ? 0 : -1;
    // arena (int32 units): [{x,y} sentinel + m][records 8 R]
    const size_t max_records = 1 + (size_t)4 * m + 16;
    const size_t need = 2 * (size_t)(m + 1) + 8 * max_records + 2;
    if (scratch.storage.size() < need) scratch.storage.resize(need);
    int32_t *base = scratch.storage.data();
    base += ((uintptr_t)base & 7) ? 1 : 0;
    Pt *P = (Pt *)base + 1;
    int32_t *R = (int32_t *)(P + m);
    P[-1] = Pt{0, 0};
    for (int i = 0; i < m; i++) {
        const int id = order[i];
        if ((unsigned)id >= (unsigned)n) return -1;
        const int32_t *sp = support + 3 * id;
        P[i] = Pt{right_image ? sp[0] - sp[2] : sp[0], sp[1]};  // elas.cpp:451-461
    }
    Mesh mesh;
    mesh.P = P;
    mesh.R = R;
    mesh.ntri = 0;
    mesh.make();  // record 0: outer space
    int hullleft, hullright;
    mesh.recurse(0, m, 0, 0, 1, -1, hullleft, hullright);
    int count = 0;
    for (int t = 1; t < 2 * m - 1; t++) {
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

int delaunay_support_levels(const int32_t *support, int n, int right_image, const int32_t *order, int host_levels, int32_t *tri_out, int cap,
                            DelaunayScratch &scratch) {
    if (n < 3) return 0;
    if (n > 4096) return -1;
    const size_t records = 2 * (size_t)n;
    // arena (int32 units): [{x,y} sentinel + n][int32 records 8 R][uint16 records 8 R]
    const size_t need = 2 * (size_t)(n + 1) + 8 * records + 4 * records + 4;
    if (scratch.storage.size() < need) scratch.storage.resize(need);
    int32_t *base = scratch.storage.data();
    base += ((uintptr_t)base & 7) ? 1 : 0;
    Pt *P = (Pt *)base + 1;
    int32_t *R32 = (int32_t *)(P + n);
    uint16_t *R16 = (uint16_t *)(R32 + 8 * records);
    P[-1] = Pt{0, 0};
    for (int i = 0; i < n; i++) {
        const int id = order[i];
        if ((unsigned)id >= (unsigned)n) return -1;
        const int32_t *sp = support + 3 * id;
        P[i] = Pt{right_image ? sp[0] - sp[2] : sp[0], sp[1]};
    }
    // "device": levels max_depth .. host_levels, every node of a level independently of the others
    MeshT<uint16_t> dm;
    dm.P = P;
    dm.R = R16;
    dm.ntri = 0;
    dm.make();
    const int max_depth = delaunay_max_depth(n);
    if (host_levels < 0) host_levels = 0;
    for (int depth = max_depth; depth >= host_levels; depth--)
        for (int k = (1 << depth) - 1; k >= 0; k--) {  // any order within a level must do: run it backwards
            const DelaunayNode nd = delaunay_node_at(n, depth, k);
            if (nd.exists) dm.build_node(nd);
        }
    // records widen to the host's 32-bit layout (neighbour handles zero-extended, vertex ids sign-extended)
    for (size_t t = 0; t < 2 * (size_t)n - 1; t++)
        for (int q = 0; q < 8; q++) R32[8 * t + q] = (q & 3) == 3 ? (int32_t)R16[8 * t + q] : (q < 4 ? (int32_t)R16[8 * t + q] : (int32_t)(int16_t)R16[8 * t + q]);
    Mesh mesh;
    mesh.P = P;
    mesh.R = R32;
    mesh.ntri = 0;
    if (host_levels > 0) {
        int hullleft, hullright;
        mesh.recurse(0, n, 0, 0, 1, host_levels, hullleft, hullright);
    }
    int count = 0;
    for (int t = 1; t < 2 * n - 1; t++) {
        const int32_t *r = R32 + 8 * t + 4;
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
