"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the reference's serial ELAS
(oracle/_ref) on the same inputs, and against the committed golden fixtures.

Bars (BASELINE.json north_star): bit-exact descriptors, support matches and pre-filter integer disparities; the
filtered float disparity within 1e-3 px on >= 99.9 % of valid pixels with an identical invalid mask (the
kernels actually reproduce it bit for bit, which is what is asserted); point cloud within 1e-4 relative."""
import os

import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu

FLOAT_TOL_PX = 1e-3  # north_star tolerance for post-filter disparities
FLOAT_FRAC = 0.999
POINT_RTOL = 1e-4  # north_star tolerance for the point cloud


def presets(svb, ref):
    return {
        "robotics": (svb.default_params(svb.ROBOTICS), ref.params(0)),
        "pipeline": (svb.default_params(svb.PIPELINE), ref.pipeline_params()),
        "middlebury": (svb.default_params(svb.MIDDLEBURY), ref.params(1)),
    }


def assert_all_equal(res, allow=()):
    bad = {k: v for k, v in res.items() if isinstance(v, dict) and not v["equal"] and k not in allow}
    assert not bad, "stages differing from the oracle: %s" % bad


def assert_float_bar(res):
    assert res["D1_mask_equal"]
    assert res.get("D1_frac_within_1e-3", 1.0) >= FLOAT_FRAC


@pytest.mark.parametrize("pname", ["robotics", "pipeline", "middlebury"])
def test_kitti0_staged_parity_with_injected_triangles(svb, ref, kitti_gray, pname):
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    p, p_ref = presets(svb, ref)[pname]
    ctx = svb.Context(p, L.shape[1], L.shape[0])
    try:
        res, t, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=True)
        assert_all_equal(res)
        assert_float_bar(res)
        iso = parity.isolated_parity(ctx, ref, p_ref, t)
        assert_all_equal(iso)
    finally:
        ctx.close()


@pytest.mark.parametrize("frame", [0, 7])
@pytest.mark.parametrize("pname", ["robotics", "pipeline"])
def test_kitti_end_to_end_against_golden(svb, kitti_gray, golden, pname, frame):
    """Own host Delaunay, no oracle in the loop: compare with the committed outputs of the reference."""
    L, R = kitti_gray["L%d" % frame], kitti_gray["R%d" % frame]
    p = svb.default_params(svb.ROBOTICS if pname == "robotics" else svb.PIPELINE)
    ctx = svb.Context(p, L.shape[1], L.shape[0])
    try:
        ctx.set_tap_mode(True)
        D1, D2 = ctx.process(L, R)
        key = "%s_%d" % (pname, frame)
        assert np.array_equal(ctx.tap("support"), golden[key + "_support"])
        assert np.array_equal(ctx.tap("tri1"), golden[key + "_tri1"])
        assert np.array_equal(ctx.tap("tri2"), golden[key + "_tri2"])
        assert np.array_equal(ctx.tap("D1raw").astype(np.int16), golden[key + "_D1raw"])
        assert np.array_equal(ctx.tap("D2raw").astype(np.int16), golden[key + "_D2raw"])
        want = golden[key + "_D1"]
        assert np.array_equal(D1 >= 0, want >= 0)
        both = want >= 0
        err = np.abs(D1[both] - want[both])
        assert (err <= FLOAT_TOL_PX).mean() >= FLOAT_FRAC
        assert np.array_equal(D1, want)  # in fact bit-exact
    finally:
        ctx.close()


@pytest.mark.parametrize("W,H,slanted,pname", [(1242, 375, 0, "pipeline"), (1242, 375, 1, "robotics"), (640, 240, 0, "robotics"),
                                                (333, 127, 1, "pipeline"), (1920, 1080, 0, "pipeline")])
def test_synthetic_end_to_end_parity(svb, ref, W, H, slanted, pname):
    """Synthetic frames of the bench workload (and the 1080p config) end to end with the product's own Delaunay;
    ragged sizes (W, H not multiples of the tile / lattice / grid sizes) included."""
    L, R = svb.synth_pair(5, W, H, slanted)
    p, p_ref = presets(svb, ref)[pname]
    ctx = svb.Context(p, W, H)
    try:
        res, t, (D1, _) = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
        assert_all_equal(res)
        assert_float_bar(res)
        if not slanted and W >= 640:
            # sanity against the known scene: >= 95 % of valid pixels within 1 px of the true band disparity
            truth = np.empty((H, W), np.float32)
            truth[: H // 3] = 8
            truth[H // 3: 2 * H // 3] = 24
            truth[2 * H // 3:] = 48
            v = D1 >= 0
            assert v.mean() > 0.5
            assert (np.abs(D1[v] - truth[v]) <= 1).mean() >= 0.95
    finally:
        ctx.close()


def test_4k_disp512_single_gpu_parity(svb, ref):
    """BASELINE.json configs[3] shape: 3840x2160, disparity range 512 (disp_max 511), whole pipeline with L/R check and
    all post-filters, on one GPU.  Exercises the large-frame code paths (lattice filters in global memory, 16-word
    cell masks, the global-memory gap column walk)."""
    W, H = 3840, 2160
    L, R = svb.synth_pair(9, W, H, 0)
    p = svb.default_params(svb.MIDDLEBURY, disp_max=511)
    p_ref = ref.params(1, disp_max=511)
    ctx = svb.Context(p, W, H)
    try:
        D1, D2 = ctx.process(L, R)
        R1, R2, _ = ref.process(p_ref, L, R)
        assert np.array_equal(D1, R1)
        assert np.array_equal(D2, R2)
        assert (D1 >= 0).mean() > 0.5
    finally:
        ctx.close()


@pytest.mark.parametrize("W,H,setting", [(1242, 375, "pipeline"), (1242, 375, "robotics"), (640, 241, "middlebury"), (1241, 376, "pipeline")])
def test_subsampling_parity(svb, ref, kitti_gray, W, H, setting):
    """subsampling = 1 (elas.h:81-83): descriptors on even rows only, lattice step 6, disparities for even (u, v) only in
    (W/2) x (H/2) maps, d/2 warps in the L/R check, 28-pixel speckles, gap width/2+1, 4-tap adaptive mean -- stage by
    stage against the oracle, odd and even image sizes."""
    if (W, H) == (1242, 375):
        L, R = kitti_gray["L0"], kitti_gray["R0"]
    else:
        L, R = svb.synth_pair(21, W, H, 1)
    over = {"subsampling": 1}
    if setting == "pipeline":
        p, p_ref = svb.default_params(svb.PIPELINE, **over), ref.pipeline_params(subsampling=1)
    elif setting == "robotics":
        p, p_ref = svb.default_params(svb.ROBOTICS, **over), ref.params(0, **over)
    else:
        p, p_ref = svb.default_params(svb.MIDDLEBURY, **over), ref.params(1, **over)
    ctx = svb.Context(p, W, H)
    try:
        assert (ctx.Dw, ctx.Dh) == (W // 2, H // 2)
        res, t, (D1, D2) = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
        assert_all_equal(res)
        assert_float_bar(res)
        assert D1.shape == (H // 2, W // 2) and (D1 >= 0).mean() > 0.3
        # Elas::process of the oracle in one go gives the same half-size maps
        R1, R2, _ = ref.process(p_ref, L, R)
        assert np.array_equal(D1, parity.half(R1, ctx)) and np.array_equal(D2, parity.half(R2, ctx))
    finally:
        ctx.close()


def test_few_support_points_leaves_outputs_untouched(svb, ref):
    """elas.cpp:64-69: a textureless pair yields < 3 support points; D1/D2 stay as the caller passed them."""
    W, H = 320, 120
    L = np.full((H, W), 77, np.uint8)
    p = svb.default_params(svb.ROBOTICS)
    ctx = svb.Context(p, W, H)
    try:
        with pytest.raises(svb.SvbError) as e:
            ctx.process(L, L)
        assert e.value.code == svb.ERR_FEW_SUPPORT
        D1, D2, _ = ref.process(ref.params(0), L, L)
        assert not D1.any() and not D2.any()  # the oracle leaves the zero-initialised buffers alone too
    finally:
        ctx.close()


def test_uniform_noise_pair(svb, ref):
    """Uncorrelated noise: most candidates fail the ratio / L-R tests; whatever survives must match exactly."""
    rng = np.random.default_rng(9)
    W, H = 400, 160
    L = rng.integers(0, 256, (H, W), dtype=np.uint8)
    R = rng.integers(0, 256, (H, W), dtype=np.uint8)
    p, p_ref = presets(svb, ref)["pipeline"]
    ctx = svb.Context(p, W, H)
    try:
        t = ref.staged(p_ref, L, R)
        if len(t["support"]) < 3:
            with pytest.raises(svb.SvbError):
                ctx.process(L, R)
        else:
            res, _, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
            assert_all_equal(res)
    finally:
        ctx.close()


def test_process_is_deterministic_and_restartable(svb, kitti_gray):
    L, R = kitti_gray["L7"], kitti_gray["R7"]
    p = svb.default_params(svb.PIPELINE)
    ctx = svb.Context(p, L.shape[1], L.shape[0])
    try:
        a = ctx.process(L, R)
        b = ctx.process(R, L)  # different content in between
        c = ctx.process(L, R)
        assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
        assert not np.array_equal(a[0], b[0])
        st = ctx.stats()
        assert st["kernel_launches"] >= 15 and st["frames"] == 1
    finally:
        ctx.close()


def test_strided_input(svb, kitti_gray, golden):
    """dims[2] (bytes per line) larger than the width, as Elas::process allows (elas.cpp:33-50)."""
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    H, W = L.shape
    stride = W + 37
    Lp = np.full((H, stride), 255, np.uint8)
    Rp = np.full((H, stride), 255, np.uint8)
    Lp[:, :W] = L
    Rp[:, :W] = R
    p = svb.default_params(svb.ROBOTICS)
    ctx = svb.Context(p, W, H)
    try:
        import ctypes as C

        D1 = np.zeros((H, W), np.float32)
        D2 = np.zeros((H, W), np.float32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = ctx.lib.svb_process(ctx.h, vp(Lp), vp(Rp), stride, vp(D1), vp(D2))
        assert rc == 0
        assert np.array_equal(D1, golden["robotics_0_D1"])
    finally:
        ctx.close()


def test_adaptive_mean_true_abs_switch(svb, ref, kitti_gray, golden):
    """SURVEY.md finding 3: mode 1 is the parallel reference's true-abs weight; it must differ from the serial
    bit-mask form only through the weights (numpy restatement of src/parallel_includes/elas/elas.cpp:1552-1612)."""
    D = golden["robotics_0_D1raw"].astype(np.float32)
    D[D < 0] = -10
    H, W = D.shape
    p = svb.default_params(svb.ROBOTICS)
    ctx = svb.Context(p, W, H)
    try:
        serial = ctx.adaptive_mean(D)
        assert np.array_equal(serial, ref.adaptive_mean(ref.params(0), D))
        ctx.set_mean_mode(1)
        true_abs = ctx.adaptive_mean(D)
        ctx.set_mean_mode(0)
        # numpy restatement with |x| weights (float64 accumulation: compare with a tolerance)
        Dc = np.where(D < 0, -10.0, D).astype(np.float64)
        tmp = np.where(D < 0, -10.0, 0.0)
        for c in range(4, W - 3):
            win = Dc[3:H - 3, c - 4:c + 4]
            w = np.maximum(0, 4 - np.abs(win - win[:, 4:5]))
            ws = w.sum(1)
            val = (w * win).sum(1) / np.where(ws > 0, ws, 1)
            ok = (ws > 0) & (val >= 0)
            tmp[3:H - 3, c] = np.where(ok, val, tmp[3:H - 3, c])
        out = D.astype(np.float64).copy()
        for c in range(4, H - 3):
            win = tmp[c - 4:c + 4, 3:W - 3]
            w = np.maximum(0, 4 - np.abs(win - win[4:5, :]))
            ws = w.sum(0)
            val = (w * win).sum(0) / np.where(ws > 0, ws, 1)
            ok = (ws > 0) & (val >= 0)
            out[c, 3:W - 3] = np.where(ok, val, out[c, 3:W - 3])
        assert np.allclose(true_abs, out, atol=1e-3)
        assert not np.array_equal(true_abs, serial)
    finally:
        ctx.close()


def test_reproject_parity(svb, golden, golden_meta):
    D = golden["pipeline_0_D1"]
    H, W = D.shape
    ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H)
    try:
        for XR, XT in ((np.eye(3), np.zeros(3)), (np.array(golden_meta["XR"]), np.array(golden_meta["XT"]))):
            dm, pts = ctx.reproject(D, np.array(golden_meta["Q"]), XR, XT)
            dm_o, pts_o = parity.reproject_oracle(D, golden_meta["Q"], XR, XT)
            assert np.array_equal(dm, dm_o)
            fin = np.isfinite(pts_o).all(1)
            assert np.array_equal(np.isfinite(pts).all(1), fin)
            assert np.array_equal(np.isnan(pts), np.isnan(pts_o))
            rel = np.abs(pts[fin] - pts_o[fin]) / np.maximum(np.abs(pts_o[fin]), 1e-300)
            assert rel.max() <= POINT_RTOL
            assert np.array_equal(pts, pts_o, equal_nan=True)  # same operation order, correctly rounded quotients: bit-identical
        # general Q (w depends on x and y too) and rows scaled into the exponent ranges where the shared-reciprocal
        # division falls back to IEEE division (tiny numerators, huge / tiny quotients)
        Qg = np.array(golden_meta["Q"], np.float64)
        Qg[3, 0], Qg[3, 1], Qg[3, 3] = 1e-4, -3e-4, 0.37
        for scale in ((1, 1, 1, 1), (1e-42, 1e-300, 1e200, 1), (1, 1e250, 1, 1e-60), (1e-310, 1, 1, 1e300)):
            Qs = Qg * np.array(scale, np.float64)[:, None]
            dm, pts = ctx.reproject(D, Qs, np.array(golden_meta["XR"]), np.array(golden_meta["XT"]))
            with np.errstate(all="ignore"):
                dm_o, pts_o = parity.reproject_oracle(D, Qs, golden_meta["XR"], golden_meta["XT"])
            assert np.array_equal(pts, pts_o, equal_nan=True), scale
        # a map with invalid pixels: d8 = 0 -> w = 0 -> inf / nan exactly like the reference kernel
        D2 = golden["robotics_0_D1"]
        dm, pts = ctx.reproject(D2, np.array(golden_meta["Q"]))
        dm_o, pts_o = parity.reproject_oracle(D2, golden_meta["Q"], np.eye(3), np.zeros(3))
        assert np.array_equal(dm, dm_o)
        assert np.array_equal(np.isfinite(pts), np.isfinite(pts_o))
        assert np.array_equal(pts, pts_o, equal_nan=True)
    finally:
        ctx.close()


def test_reproject_against_the_reference_kernel(svb, golden, golden_meta, kitti_gray):
    """Row 19 pinned to the reference ITSELF: `projectParallel` (stereo_vision.cu:188-212), cut out of the reference's driver and
    compiled by oracle/build_ref.sh (oracle/_ref/libproject_ref.so, reference Makefile flags = FMA contraction on), run on this
    GPU.  Checked, BIT FOR BIT (NaNs at the same places): (1) the product's u8 + reprojection kernel, stand-alone and as the last
    phase of the fused tail kernel, which spell out the reference build's fused multiply-adds; (2) the CPU restatement
    oracle/project_port.c behind tests/parity.py::reproject_oracle, which the other tests use.  The formula without contraction
    (parity.reproject_numpy) is reported next to them: it differs by rounding only (<= 1e-4 relative, north_star's bar)."""
    from oracle.ref import RefProject

    rp = RefProject()
    assert "nvcc" in rp.flags and "-O2" in rp.flags
    Q = np.array(golden_meta["Q"])
    Qg = Q.copy()
    Qg[3, 0], Qg[3, 1], Qg[3, 3] = 1e-4, -3e-4, 0.37  # a general Q: w depends on x and y too
    Qn = Q.copy()
    Qn[3, 2], Qn[0, 1], Qn[1, 0], Qn[2, 2] = -Q[3, 2], 0.013, -0.021, 0.5  # negative w, every entry of the upper rows in play
    XR, XT = np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
    cases = [(golden["pipeline_0_D1"], Q, np.eye(3), np.zeros(3)),
             (golden["pipeline_7_D1"], Q, XR, XT),
             (golden["robotics_0_D1"], Q, XR, XT),  # invalid pixels: d8 = 0 -> w = 0
             (golden["robotics_7_D1"], Qg, XR, XT),
             (golden["pipeline_7_D1"], Qn, XR.T.copy(), -XT)]
    H, W = cases[0][0].shape
    ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H)

    def same_bits(got, want, name):
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), name
        neq = (got.view(np.uint64) != want.view(np.uint64)) & ~nan
        assert not neq.any(), (name, int(neq.sum()), got[neq][:4], want[neq][:4])

    try:
        worst = 0.0
        for D, q, xr, xt in cases:
            dm, pts = ctx.reproject(D, q, xr, xt)
            dm_o, pts_o = parity.reproject_oracle(D, q, xr, xt)
            assert np.array_equal(dm, dm_o)
            want = rp.project(dm, q, xr, xt)  # the reference kernel on the product's u8 map (= cv convertTo of the float map)
            same_bits(pts, want, "product (k_reproject)")
            same_bits(pts_o, want, "oracle/project_port.c")
            same_bits(svb.reproject_u8(dm, q, xr, xt), want, "product (k_reproject_u8)")
            plain = parity.reproject_numpy(dm.astype(np.float64), q, xr, xt)
            assert np.array_equal(np.isnan(plain), np.isnan(want)) and np.array_equal(np.isinf(plain), np.isinf(want))
            fin = np.isfinite(want)
            rel = np.abs(plain[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300)
            assert rel.max() <= POINT_RTOL
            worst = max(worst, float(rel.max()))
        print("uncontracted formula against the reference kernel, worst relative difference: %.3g" % worst)
    finally:
        ctx.close()
    # the fused tail kernel's projection (the batch pipeline's path) on real frames, against the reference kernel run on the u8
    # map of the same disparity
    L = np.stack([kitti_gray["L0"], kitti_gray["L7"]])
    R = np.stack([kitti_gray["R0"], kitti_gray["R7"]])
    H, W = L.shape[1:]
    ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=2)
    try:
        ctx.set_calibration(Qg, XR, XT)
        ctx.batch_upload(L, R)
        ctx.batch_run(2, svb.OUT_DISPARITY | svb.OUT_POINTS)
        for i in range(2):
            dm, _ = parity.reproject_oracle(ctx.batch_disparity(i), Qg, XR, XT)
            same_bits(ctx.batch_points(i), rp.project(dm, Qg, XR, XT), "product (k_post_fused), frame %d" % i)
    finally:
        ctx.close()


@pytest.mark.parametrize("setting,support_m,dense_m", [("ROBOTICS", 4.84, 7.09), ("MIDDLEBURY", 5.89, 11.43)])
def test_hypothesis_counters_match_instrumented_reference(svb, kitti_gray, setting, support_m, dense_m):
    """svb_set_eval_counting: the counters behind bench.py's pixel-disparity evals/s.  Known answers = the counts of an
    instrumented copy of the reference on kitti frame 0 (three significant digits)."""
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    ctx = svb.Context(svb.default_params(getattr(svb, setting)), L.shape[1], L.shape[0])
    try:
        D_plain = ctx.process(L, R)
        ctx.set_eval_counting(True)
        D_counted = ctx.process(L, R)
        s, d = ctx.eval_counts()
        assert abs(s / 1e6 - support_m) <= 0.0051, s
        assert abs(d / 1e6 - dense_m) <= 0.0051, d
        assert ctx.eval_counts() == (0, 0)  # reading resets
        for a, b in zip(D_plain, D_counted):  # the counting variants compute the same maps
            assert np.array_equal(a, b)
        ctx.set_eval_counting(False)
    finally:
        ctx.close()


@pytest.mark.parametrize("dups", ["device", "host"])
def test_device_vertex_order_feeds_the_same_triangulation(svb, ref, monkeypatch, dups):
    """The Delaunay stage as the pipeline runs it -- vertex order (k_order.cu, including the replay of the reference's randomised sort
    for lists with duplicate coordinates) and divide-and-conquer (k_delaunay.cu) on the device -- against the reference's triangulator
    on random lattices, both sides; a list the device hands back to the host (more than 4096 points) still gives the reference's output."""
    import test_cabi_host as H

    monkeypatch.setenv("SVB_DELAUNAY_DUPS", dups)  # who replays the reference's sort for lists with duplicate coordinates
    ctx = svb.Context(svb.default_params(svb.MIDDLEBURY), 1242, 375)
    monkeypatch.delenv("SVB_DELAUNAY_DUPS")
    try:
        for seed, n, expect_device in [(1, 3, True), (2, 4, True), (3, 7, True), (4, 50, True), (5, 400, True), (6, 2500, True),
                                       (8, 4096, True), (9, 5, True), (10, 6, True), (11, 9, True), (12, 13, True),
                                       (13, 1001, True), (14, 1954, True), (7, 6000, False)]:
            s = H.lattice_support(np.random.default_rng(seed), n)
            for side in (0, 1):
                want = ref.delaunay(s, side)
                got, used = ctx.delaunay_pipeline(s, side)
                assert np.array_equal(got, want), "seed %d n %d side %d" % (seed, n, side)
                xs = s[:, 0] - s[:, 2] if side else s[:, 0]
                has_dups = len(set(zip(xs.tolist(), s[:, 1].tolist()))) < len(s)
                if has_dups and dups == "host":
                    assert used == 0, (seed, n, side, used)  # flagged by the device, made by the host stage
                else:
                    assert (used > 0) == (expect_device and len(s) <= 4096), (seed, n, side, used)
                    if used:
                        # the whole stage ran on the device (k_order.cu + k_delaunay.cu) -- also for the right image, where x = u - d
                        # collides for many of these lists and the survivor is whatever the reference's randomised sort leaves first
                        assert used == 2, (seed, n, side, used)
            if n >= 400:
                xr = s[:, 0] - s[:, 2]
                assert len(set(zip(xr.tolist(), s[:, 1].tolist()))) < len(s), "the generator is expected to produce duplicates in the right image"
        # a small list with duplicates in the right image
        dup = np.array([(100, 50, 10), (95, 50, 5), (200, 80, 20), (60, 120, 1), (300, 20, 9), (110, 50, 20)], np.int32)
        got, used = ctx.delaunay_pipeline(dup, 1)
        assert used == (2 if dups == "device" else 0) and np.array_equal(got, ref.delaunay(dup, 1))
        got, used = ctx.delaunay_pipeline(dup, 0)
        assert used == 2 and np.array_equal(got, ref.delaunay(dup, 0))
        # a single distinct vertex is left after the duplicate removal: no triangle (the reference's triangulator itself crashes on this
        # input, so there is nothing to compare with)
        one = np.array([(100, 50, 10), (95, 50, 5), (105, 50, 15)], np.int32)
        got, used = ctx.delaunay_pipeline(one, 1)
        assert len(got) == 0
        # collinear input: no triangles either way
        col = np.array([(50, 5 * i, 3) for i in range(1, 30)], np.int32)
        got, used = ctx.delaunay_pipeline(col, 0)
        assert used and len(got) == 0
    finally:
        ctx.close()


def _sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _digests():
    import json
    import os

    from conftest import GOLDEN

    return json.load(open(os.path.join(GOLDEN, "dataset_digests.json")))


@pytest.mark.parametrize("name", ["cones", "aloe", "raindeer", "urban1", "urban2", "urban3", "urban4"])
def test_reference_profile_pairs(svb, ref, name):
    """ALL SEVEN of the reference's own runProfiling inputs (datasets/profile: 900x750, 1282x1110, 1342x1110 -- the largest
    inputs the reference ships -- and 4 x 1344x391) with runProfiling's parameters (default preset, postprocess_only_left =
    false, stereo_vision.cu:727-730): both maps bit for bit against the live oracle AND against the committed digests."""
    import os

    from conftest import GOLDEN

    z = np.load(os.path.join(GOLDEN, "profile_gray.npz"))
    L, R = z[name + "_L"], z[name + "_R"]
    p = svb.default_params(svb.ROBOTICS, postprocess_only_left=0)
    ctx = svb.Context(p, L.shape[1], L.shape[0])
    try:
        D1, D2 = ctx.process(L, R)
        nsup = ctx.stats()["support_points"]
    finally:
        ctx.close()
    W1, W2, _ = ref.process(ref.params(0, postprocess_only_left=0), L, R)
    assert np.array_equal(D1, W1) and np.array_equal(D2, W2)
    want = _digests()["profile_runprofiling"][name]
    assert (nsup, _sha(D1), _sha(D2)) == (want["support"], want["D1"], want["D2"])
    assert (D1 >= 0).mean() > 0.5


def test_kitti_mini_all_21_pairs(svb, ref, kitti_gray):
    """Every stereo pair of datasets/kitti_mini (BASELINE configs[0]) through the frame-batch pipeline with the driver's preset:
    disparity bit for bit against the live oracle and against the committed digests, support point counts included; the point
    cloud of every frame against the CPU restatement of projectParallel."""
    n = 21
    L = np.stack([kitti_gray["L%d" % i] for i in range(n)])
    R = np.stack([kitti_gray["R%d" % i] for i in range(n)])
    H, W = L.shape[1:]
    dig = _digests()["kitti_pipeline"]
    ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=8)
    try:
        Q = np.array([[1.0, 0, 0, -738.7995529174805], [0, 1.0, 0, -254.7572193145752], [0, 0, 0, 1027.8551581758902], [0, 0, 1.8616160699568378, 0]])
        ctx.set_calibration(Q)
        ctx.batch_upload(L, R)
        ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
        nsup = ctx.batch_frame_support(n)
        for i in range(n):
            D1 = ctx.batch_disparity(i)
            W1, _, _ = ref.process(ref.pipeline_params(), L[i], R[i])
            assert np.array_equal(D1, W1), "kitti pair %d" % i
            assert (int(nsup[i]), _sha(D1), int((D1 >= 0).sum())) == (dig[str(i)]["support"], dig[str(i)]["D1"], dig[str(i)]["valid1"]), i
            if i % 5 == 0:
                _, pts_o = parity.reproject_oracle(D1, Q, np.eye(3), np.zeros(3))
                pts = ctx.batch_points(i)
                fin = np.isfinite(pts_o).all(1)
                assert np.array_equal(np.isfinite(pts).all(1), fin)
                assert (np.abs(pts[fin] - pts_o[fin]) / np.maximum(np.abs(pts_o[fin]), 1e-300)).max() <= POINT_RTOL
    finally:
        ctx.close()


@pytest.mark.parametrize("gamma,beta,sigma,sradius", [(1.0, 0.01, 1.0, 2.0), (0.5, 0.004, 1.5, 3.0), (15.0, 0.05, 0.8, 2.0)])
def test_non_preset_prior_parameters(svb, ref, kitti_gray, gamma, beta, sigma, sradius):
    """Elas::parameters is a public drop-in field: gamma / beta / sigma outside the two presets give prior tables far below the
    presets' -14 (P[0] = log(gamma / (gamma + 1)) / beta: -69, -274, ...), which the packed matching key has to carry
    (Dims::cost_bias).  Raw integer disparities and the final maps against the oracle."""
    L, R = kitti_gray["L3"], kitti_gray["R3"]
    over = dict(gamma=gamma, beta=beta, sigma=sigma, sradius=sradius)
    p = svb.default_params(svb.ROBOTICS, **over)
    p_ref = ref.params(0, **over)
    ctx = svb.Context(p, L.shape[1], L.shape[0])
    try:
        res, t, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
        assert_all_equal(res)
        assert_float_bar(res)
    finally:
        ctx.close()


def _random_params(rng):
    """A random point of Elas::parameters' space (elas.h:58-83) inside what the reference itself handles."""
    sigma = float(rng.choice([0.7, 1.0, 1.5]))
    sradius = float(rng.choice([2.0, 3.0]))
    return dict(
        disp_min=int(rng.choice([0, 0, 2])), disp_max=int(rng.choice([63, 127, 255])),
        support_threshold=float(rng.choice([0.8, 0.85, 0.95])), support_texture=int(rng.choice([5, 10, 20])),
        candidate_stepsize=int(rng.choice([3, 5, 7])), incon_window_size=int(rng.choice([3, 5])), incon_threshold=int(rng.choice([3, 5, 7])),
        incon_min_support=int(rng.choice([3, 5])), add_corners=int(rng.integers(0, 2)), grid_size=int(rng.choice([12, 20, 31])),
        beta=float(rng.choice([0.02, 0.03])), gamma=float(rng.choice([3.0, 5.0, 8.0])), sigma=sigma, sradius=sradius,
        match_texture=int(rng.choice([0, 1, 4])), lr_threshold=int(rng.choice([1, 2, 3])), speckle_sim_threshold=float(rng.choice([1.0, 2.0])),
        speckle_size=int(rng.choice([50, 200, 400])), ipol_gap_width=int(rng.choice([3, 7, 5000])), filter_median=int(rng.integers(0, 2)),
        filter_adaptive_mean=int(rng.integers(0, 2)), postprocess_only_left=int(rng.integers(0, 2)), subsampling=0)


@pytest.mark.parametrize("seed", list(range(24)))
def test_random_parameter_sets_against_the_oracle(svb, ref, kitti_gray, seed):
    """Elas::parameters is a public drop-in struct: a dozen random parameter sets (lattice step, grid size, windows, thresholds,
    prior shape and plane radius, filters on / off, corners, gap width, both-map post-processing) on a real frame and on a ragged
    synthetic one, with and without subsampling, final maps bit for bit against the oracle; the raw integer maps too (tap mode) for
    every third set."""
    rng = np.random.default_rng(1000 + seed)
    over = _random_params(rng)
    if seed % 4 == 3:
        over["subsampling"] = 1  # half-resolution maps (elas.h:81-83)
    if seed % 2 == 0:
        L, R = kitti_gray["L%d" % (seed % 21)], kitti_gray["R%d" % (seed % 21)]
        L, R = np.ascontiguousarray(L[40:300, 100:900]), np.ascontiguousarray(R[40:300, 100:900])  # 800 x 260 crop
    else:
        L, R = svb.synth_pair(50 + seed, 517, 203, seed & 2)
    H, W = L.shape
    p = svb.default_params(svb.ROBOTICS, **over)
    p_ref = ref.params(0, **over)
    ctx = svb.Context(p, W, H)
    try:
        want1, want2, _ = ref.process(p_ref, L, R)
        if seed % 3 == 0:
            res, _, (D1, D2) = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
            assert_all_equal(res)
        else:
            D1, D2 = ctx.process(L, R)
        assert np.array_equal(D1, parity.half(want1, ctx)), over
        assert np.array_equal(D2, parity.half(want2, ctx)), over
    finally:
        ctx.close()


@pytest.mark.parametrize("sweeps", ["8", "1"])
def test_multi_cta_lattice_filters(svb, ref, kitti_gray, monkeypatch, sweeps):
    """The large-lattice form of the order-dependent lattice filters (one launch per fixpoint sweep, strip kernels for the two
    redundant-point passes, column-wise compaction: what 4K frames use) forced onto KITTI-size and ragged frames: candidate lattice,
    support list and everything downstream against the oracle.  sweeps = 1 cuts the multi-CTA sweeps short so that the one-CTA
    finisher has to complete the fixpoint iteration."""
    monkeypatch.setenv("SVB_SF_MULTI", "1")
    monkeypatch.setenv("SVB_SF_SWEEPS", sweeps)
    import subprocess
    import sys
    import textwrap

    # the switches are read once per process: run the check in a fresh interpreter
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import numpy as np
        from conftest import load_binding
        import parity
        from oracle.ref import RefElas
        svb = load_binding().binding
        ref = RefElas()
        z = np.load(%r)
        cases = [(z["L3"], z["R3"], svb.default_params(svb.PIPELINE), ref.pipeline_params()),
                 (z["L11"], z["R11"], svb.default_params(svb.ROBOTICS), ref.params(0))]
        Ls, Rs = svb.synth_pair(31, 517, 203, 1)
        cases.append((Ls, Rs, svb.default_params(svb.MIDDLEBURY), ref.params(1)))
        for L, R, p, p_ref in cases:
            ctx = svb.Context(p, L.shape[1], L.shape[0])
            res, t, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
            bad = {k: v for k, v in res.items() if isinstance(v, dict) and not v["equal"]}
            assert not bad, bad
            ctx.close()
        print("multi-CTA lattice filters: OK")
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
            os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kitti_gray.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("W,H", [(1242, 375), (97, 66), (33, 31), (64, 32), (640, 1217)])
def test_gap_interpolation_on_random_validity_maps(svb, ref, W, H):
    """gapInterpolation alone (elas.cpp:1126-1295) on maps whose holes are drawn at random: isolated pixels, long runs, whole invalid
    rows and columns, holes touching every border, widths and heights that are and are not multiples of the 32-position words the
    kernels work on; gap widths below, around and above the word size, corners on and off."""
    rng = np.random.default_rng(W * 1000 + H)
    for case, (gap, corners, density) in enumerate([(3, 0, 0.3), (3, 1, 0.7), (7, 1, 0.5), (31, 1, 0.9), (40, 0, 0.95), (5000, 1, 0.97),
                                                   (5000, 0, 0.5), (5000, 1, 0.999)]):
        D = rng.integers(0, 64, (H, W)).astype(np.float32) + rng.integers(0, 4, (H, W)).astype(np.float32) / 4
        holes = rng.random((H, W)) < density
        for _ in range(6):  # whole lines, and lines that are invalid up to / from a random position
            holes[rng.integers(0, H), :] = True
            holes[:, rng.integers(0, W)] = True
            holes[rng.integers(0, H), : rng.integers(1, W)] = True
            holes[rng.integers(0, H), rng.integers(0, W - 1):] = True
            holes[: rng.integers(1, H), rng.integers(0, W)] = True
            holes[rng.integers(0, H - 1):, rng.integers(0, W)] = True
        D[holes] = np.where(rng.random(int(holes.sum())) < 0.5, -1.0, -10.0).astype(np.float32)
        if case == 7:
            D[:] = -10.0
            D[H // 2, W // 3] = 5.0  # one valid pixel in the whole map
        q = svb.default_params(svb.ROBOTICS, ipol_gap_width=gap, add_corners=corners)
        q_ref = ref.params(0, ipol_gap_width=gap, add_corners=corners)
        c = svb.Context(q, W, H)
        try:
            got = c.gap_interpolation(D)
        finally:
            c.close()
        want = ref.gap_interpolation(q_ref, D)
        assert np.array_equal(got, want), (W, H, gap, corners, density, int((got != want).sum()))


def test_patch_form_of_support_matching(svb, ref, kitti_gray, monkeypatch):
    """Frames wider than 3845 pixels (more lattice candidates per row than one CTA holds) use the patch form of support matching
    (k_support_match); SVB_MATCH_ROWS=0 forces it onto KITTI-size, ragged and subsampled frames: candidate lattice and everything
    downstream against the oracle."""
    monkeypatch.setenv("SVB_MATCH_ROWS", "0")
    import subprocess
    import sys
    import textwrap

    # the switch is read once per process: run the check in a fresh interpreter
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import numpy as np
        from conftest import load_binding
        import parity
        from oracle.ref import RefElas
        svb = load_binding().binding
        ref = RefElas()
        z = np.load(%r)
        cases = [(z["L5"], z["R5"], svb.default_params(svb.PIPELINE), ref.pipeline_params()),
                 (z["L5"], z["R5"], svb.default_params(svb.PIPELINE, subsampling=1), ref.pipeline_params(subsampling=1))]
        Ls, Rs = svb.synth_pair(32, 517, 203, 1)
        cases.append((Ls, Rs, svb.default_params(svb.ROBOTICS), ref.params(0)))
        for L, R, p, p_ref in cases:
            ctx = svb.Context(p, L.shape[1], L.shape[0])
            res, t, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
            bad = {k: v for k, v in res.items() if isinstance(v, dict) and not v["equal"]}
            assert not bad, bad
            ctx.close()
        print("patch form: OK")
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)),
            os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kitti_gray.npz"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
