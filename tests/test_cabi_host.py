"""CPU tests of the boundary and the host stage: the C-ABI library loads and exports every symbol that include/*.h
declares, refuses to compute without a CUDA device, mirrors Elas::parameters, and its host Delaunay stage
reproduces the reference's triangle lists (set AND order) on the golden support lists and on hard random cases."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        if os.path.basename(h) == "elas.h":
            continue  # C++ classes: checked through their mangled names below
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(([^;{}()]*)\)\s*;", src):
            name = m.group(1)
            if name in ("defined", "sizeof", "__attribute__"):
                continue
            names.append(name)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(svb):
    lib = svb.load()
    names = declared_symbols()
    assert len(names) >= 30, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing


def test_library_exports_the_cpp_entry_points(svb):
    """include/elas.h: Elas::parameters(setting), Elas(parameters), Elas::process(...) (Itanium-mangled)."""
    lib = svb.load()
    for sym in ("_ZN4Elas10parametersC1ENS_7settingE", "_ZN4ElasC1ENS_10parametersE", "_ZN4ElasD1Ev", "_ZN4Elas7processEPhS0_PfS1_PKi"):
        assert hasattr(lib, sym), sym
    for sym in ("generatePointCloud", "clean", "getColor"):
        assert hasattr(lib, sym), sym


def test_no_cpu_fallback_without_device(svb):
    """Without a CUDA device the library must fail loudly instead of computing on the CPU."""
    if svb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = svb.default_params(svb.ROBOTICS)
    with pytest.raises(svb.SvbError) as e:
        svb.Context(p, 640, 240)
    assert "no CUDA device" in str(e.value) or "CPU" in str(e.value)


@pytest.mark.parametrize("setting", [0, 1])
def test_default_params_mirror_reference(svb, ref, setting):
    """svb_default_params == Elas::parameters(setting) as compiled from the reference header."""
    a = svb.default_params(setting)
    b = ref.params(setting)
    for name, _ in a._fields_:
        assert getattr(a, name) == getattr(b, name), name
    assert C.sizeof(a) == C.sizeof(b) == 23 * 4


def test_pipeline_preset(svb, ref):
    a = svb.default_params(svb.PIPELINE)
    b = ref.pipeline_params()
    for name, _ in a._fields_:
        assert getattr(a, name) == getattr(b, name), name


def test_unsupported_arguments_are_rejected(svb):
    lib = svb.load()
    assert lib.svb_default_params(0, None) != 0
    assert lib.svb_stage_delaunay(None, 0, 0, None, 0, None) != 0
    assert lib.svb_synth_pair(0, 4, 4, 0, None, None) != 0


@pytest.mark.parametrize("key", ["robotics_0", "pipeline_0", "robotics_7", "pipeline_7"])
def test_host_delaunay_matches_golden_order(svb, golden, key):
    s = golden[key + "_support"]
    for side in (0, 1):
        got = svb.delaunay(s, side)
        want = golden[key + "_tri%d" % (side + 1)]
        assert got.shape == want.shape
        assert np.array_equal(got, want), "%s side %d: triangle list differs from Triangle 1.6's" % (key, side)


def lattice_support(rng, n, W=1242, H=375, step=5, dmax=60):
    """Column-major lattice support points like computeSupportMatches emits (massively co-circular)."""
    cw, ch = (W + step - 1) // step, (H + step - 1) // step
    cells = [(u, v) for u in range(1, cw) for v in range(1, ch)]
    pick = sorted(rng.choice(len(cells), size=min(n, len(cells)), replace=False))
    base = rng.integers(0, dmax)
    pts = []
    for i in pick:
        u, v = cells[i]
        d = int(np.clip(base + rng.integers(-3, 4) + (v * step) // 40, 0, min(dmax, u * step)))
        pts.append((u * step, v * step, d))
    return np.array(pts, np.int32)


@pytest.mark.parametrize("seed,n", [(1, 3), (2, 4), (3, 7), (4, 50), (5, 400), (6, 2500), (7, 6000)])
def test_host_delaunay_matches_reference_on_random_lattices(svb, ref, seed, n):
    rng = np.random.default_rng(seed)
    s = lattice_support(rng, n)
    for side in (0, 1):
        want = ref.delaunay(s, side)
        got = svb.delaunay(s, side)
        assert np.array_equal(got, want), "seed %d n %d side %d" % (seed, n, side)


def kd_order(x, y):
    """numpy statement of the vertex order k_order.cu computes: lexicographic (x, y) sort, then alternating-axis median
    cuts (x first), subsets of <= 3 left x-sorted (triangle.cpp:5243-5325, 5882-5913)."""
    idx = np.lexsort((y, x))

    def cut(ids, axis):
        if len(ids) <= 3:
            return list(ids[np.lexsort((y[ids], x[ids]))])
        k = np.lexsort((y[ids], x[ids])) if axis == 0 else np.lexsort((x[ids], y[ids]))
        ids = ids[k]
        d = len(ids) >> 1
        return cut(ids[:d], 1 - axis) + cut(ids[d:], 1 - axis)

    return np.array(cut(idx, 0), np.int32)


@pytest.mark.parametrize("seed,n", [(1, 3), (2, 4), (3, 7), (4, 50), (5, 400), (6, 2500)])
def test_host_recursion_on_a_given_vertex_order(svb, ref, seed, n):
    """delaunay_support_ordered (the host half of the pipeline's stage) fed with the order the device kernel is specified
    to deliver; left image only (lattice points are distinct there, the right image may hold duplicates)."""
    s = lattice_support(np.random.default_rng(seed), n)
    order = kd_order(s[:, 0].astype(np.int64), s[:, 1].astype(np.int64))
    assert np.array_equal(svb.delaunay_ordered(s, 0, order), ref.delaunay(s, 0))
    with pytest.raises(Exception):
        svb.delaunay_ordered(s, 0, np.zeros(len(s), np.int32) + len(s))  # not a permutation


@pytest.mark.parametrize("seed,n", [(11, 3), (12, 4), (13, 5), (14, 6), (15, 7), (16, 8), (17, 9), (18, 13), (19, 50), (20, 401), (21, 1954), (22, 4096)])
def test_level_synchronous_delaunay_matches_reference(svb, ref, seed, n):
    """The divide-and-conquer as the DEVICE runs it (k_delaunay.cu): bottom-up, level by level, every node of a level built
    independently straight into its final records (a subtree of c vertices owns 2c - 2 records, known in advance), in 16-bit
    records; host_levels = how many top levels are left to the host recursion.  Restated on the host from the same source
    (delaunay_mesh.h) -- must give the reference's triangle list, order and corner rotation included, for every split."""
    s = lattice_support(np.random.default_rng(seed), n)
    order = kd_order(s[:, 0].astype(np.int64), s[:, 1].astype(np.int64))
    want = ref.delaunay(s, 0)
    for host_levels in (0, 1, 2, 3, 5, 20):
        got = svb.delaunay_levels(s, 0, order, host_levels)
        assert np.array_equal(got, want), "n %d host_levels %d" % (n, host_levels)
    # the right image (x = u - d), when it holds no duplicate coordinates
    xr = (s[:, 0] - s[:, 2]).astype(np.int64)
    if len(set(zip(xr.tolist(), s[:, 1].tolist()))) == len(s):
        order_r = kd_order(xr, s[:, 1].astype(np.int64))
        assert np.array_equal(svb.delaunay_levels(s, 1, order_r, 0), ref.delaunay(s, 1))


def test_host_delaunay_degenerate_inputs(svb, ref):
    # all collinear (one lattice column): Triangle emits no triangle
    col = np.array([(50, 5 * i, 3) for i in range(1, 30)], np.int32)
    assert len(svb.delaunay(col, 0)) == len(ref.delaunay(col, 0)) == 0
    # one lattice row
    row = np.array([(5 * i, 100, 2) for i in range(1, 40)], np.int32)
    assert len(svb.delaunay(row, 0)) == len(ref.delaunay(row, 0)) == 0
    # fewer than 3 points
    assert len(svb.delaunay(col[:2], 0)) == 0
    # duplicates in the right image (x = u - d collides): same survivors as the reference
    dup = np.array([(100, 50, 10), (95, 50, 5), (200, 80, 20), (60, 120, 1), (300, 20, 9), (110, 50, 20)], np.int32)
    assert np.array_equal(svb.delaunay(dup, 1), ref.delaunay(dup, 1))
    # with the corner points add_corners appends (negative-free, includes u beyond W)
    rng = np.random.default_rng(11)
    s = lattice_support(rng, 300)
    corners = np.array([(0, 0, 7), (0, 374, 9), (1241, 0, 4), (1241, 374, 30), (1245, 0, 4), (1271, 374, 30)], np.int32)
    s2 = np.concatenate([s, corners])
    for side in (0, 1):
        assert np.array_equal(svb.delaunay(s2, side), ref.delaunay(s2, side))


def test_host_delaunay_coordinate_range(svb, ref):
    """The stage entry points hold a list to the range the integer predicates are exact for (x -8192 .. 16383, y 0 .. 8191: what
    frames of up to 8192 x 8192 pixels with disparities up to 4095 produce); inside it -- far outside the radix sort's fast path,
    negative x in the right image, corner points beyond the frame -- the lists still equal the reference's."""
    rng = np.random.default_rng(23)
    # an 8192 x 8192 frame's extremes: right-image x down to -4095, corner points up to 8191 + 4095
    u = rng.integers(0, 8192, 400)
    v = rng.integers(0, 8192, 400)
    d = rng.integers(0, 4096, 400)
    s = np.unique(np.stack([u, v, d], 1).astype(np.int32), axis=0)
    s = s[np.unique(s[:, :2], axis=0, return_index=True)[1]]  # one point per (u, v), as the lattice gives
    s = np.concatenate([s, np.array([(8191 + 4095, 0, 4095), (8191 + 4000, 8191, 4000)], np.int32)])
    for side in (0, 1):
        assert np.array_equal(svb.delaunay(s, side), ref.delaunay(s, side))
    for bad in ((16384, 10, 0), (10, 8192, 0), (10, -1, 0), (0, 10, 8193), (-8193, 10, 0)):
        pts = np.concatenate([s[:10], np.array([bad], np.int32)])
        side = 1 if bad[2] else 0
        with pytest.raises(svb.SvbError) as e:
            svb.delaunay(pts, side)
        assert "coordinate range" in str(e.value)
        with pytest.raises(svb.SvbError):
            svb.delaunay_ordered(pts, side, np.arange(len(pts), dtype=np.int32))


def test_synth_pair_is_deterministic_and_has_known_disparity(svb):
    L1, R1 = svb.synth_pair(3, 320, 120)
    L2, R2 = svb.synth_pair(3, 320, 120)
    assert np.array_equal(L1, L2) and np.array_equal(R1, R2)
    L3, _ = svb.synth_pair(4, 320, 120)
    assert not np.array_equal(L1, L3)
    # top band has disparity 8: right(x) ~ left(x + 8) up to +-2 noise
    diff = np.abs(R1[:40, :300].astype(int) - L1[:40, 8:308].astype(int))
    assert diff.max() <= 3
    assert L1.std() > 10


@pytest.mark.parametrize("threads", ["2", "4", "16"])
def test_host_delaunay_subtrees_in_parallel(svb, ref, monkeypatch, threads):
    """A large list (one 4K frame has 14 700 support points) builds the subtrees of one recursion depth on threads of their own --
    their record ranges are disjoint and known in advance -- and merges the levels above afterwards: same list as the reference,
    order included, whatever the thread count."""
    monkeypatch.setenv("SVB_DELAUNAY_PAR", threads)
    for seed, n in ((31, 2048), (32, 6000), (33, 15000)):
        s = lattice_support(np.random.default_rng(seed), n, W=3840, H=2160, dmax=200)
        for side in (0, 1):
            assert np.array_equal(svb.delaunay(s, side), ref.delaunay(s, side)), (seed, n, side, threads)
