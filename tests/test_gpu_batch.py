"""GPU tests of the frame-batch pipeline (BASELINE.json configs[1]): chunked, multi-lane execution with the host
Delaunay stage in between must give, frame for frame, what the single-frame Elas::process drop-in gives, both
from device-resident inputs and end to end from host buffers."""
import numpy as np
import pytest

import parity

pytestmark = pytest.mark.gpu


def make_batch(svb, n, W, H):
    L = np.zeros((n, H, W), np.uint8)
    R = np.zeros((n, H, W), np.uint8)
    for i in range(n):
        svb.synth_pair(100 + i, W, H, i & 1, L[i], R[i])
    return L, R


@pytest.mark.parametrize("chunk,n", [(4, 11), (8, 8), (1, 3)])
def test_batch_equals_single_frame_process(svb, golden_meta, chunk, n):
    W, H = 1242, 375
    L, R = make_batch(svb, n, W, H)
    p = svb.default_params(svb.PIPELINE)
    one = svb.Context(p, W, H, chunk=1)
    ctx = svb.Context(p, W, H, chunk=chunk)
    try:
        Q, XR, XT = np.array(golden_meta["Q"]), np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
        ctx.set_calibration(Q, XR, XT)
        ctx.batch_upload(L, R)
        ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
        st = ctx.stats()
        assert st["frames"] == n and st["frames_failed"] == 0 and st["kernel_launches"] > 0
        for i in range(n):
            D1, _ = one.process(L[i], R[i])
            assert np.array_equal(ctx.batch_disparity(i), D1), "frame %d" % i
            _, pts_o = parity.reproject_oracle(D1, Q, XR, XT)
            pts = ctx.batch_points(i)
            fin = np.isfinite(pts_o).all(1)
            assert np.array_equal(np.isfinite(pts).all(1), fin)
            rel = np.abs(pts[fin] - pts_o[fin]) / np.maximum(np.abs(pts_o[fin]), 1e-300)
            assert rel.max() <= 1e-4
        # run again: same result (arenas are reused, nothing may leak from the previous batch)
        ctx.batch_run(n, svb.OUT_DISPARITY)
        D1, _ = one.process(L[n - 1], R[n - 1])
        assert np.array_equal(ctx.batch_disparity(n - 1), D1)
    finally:
        ctx.close()
        one.close()


def test_batch_against_oracle(svb, ref):
    """Batch path vs the reference itself on a few frames (both presets)."""
    W, H = 1242, 375
    n = 3
    L, R = make_batch(svb, n, W, H)
    for setting, p_ref in ((svb.PIPELINE, ref.pipeline_params()), (svb.ROBOTICS, ref.params(0))):
        ctx = svb.Context(svb.default_params(setting), W, H, chunk=2)
        try:
            ctx.batch_upload(L, R)
            ctx.batch_run(n, svb.OUT_DISPARITY)
            for i in range(n):
                D1_ref, _, _ = ref.process(p_ref, L[i], R[i])
                assert np.array_equal(ctx.batch_disparity(i), D1_ref)
        finally:
            ctx.close()


def test_batch_from_host_buffers_matches_resident_path(svb, golden_meta):
    W, H = 1242, 375
    n = 6
    L, R = make_batch(svb, n, W, H)
    p = svb.default_params(svb.PIPELINE)
    ctx = svb.Context(p, W, H, chunk=4)
    try:
        ctx.set_calibration(np.array(golden_meta["Q"]))
        ctx.batch_upload(L, R)
        ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
        want_D = [ctx.batch_disparity(i) for i in range(n)]
        want_P = [ctx.batch_points(i) for i in range(n)]
        hl = svb.PinnedArray((n, H, W), np.uint8)
        hr = svb.PinnedArray((n, H, W), np.uint8)
        hD = svb.PinnedArray((n, H, W), np.float32)
        hP = svb.PinnedArray((n, H * W, 3), np.float64)
        hl.array[:] = L
        hr.array[:] = R
        ctx.batch_run_host(hl.array, hr.array, svb.OUT_DISPARITY | svb.OUT_POINTS, hD.array, hP.array)
        for i in range(n):
            assert np.array_equal(hD.array[i], want_D[i])
            assert np.array_equal(hP.array[i], want_P[i], equal_nan=True)
        for a in (hl, hr, hD, hP):
            a.free()
    finally:
        ctx.close()


def test_batch_with_a_textureless_frame(svb):
    """A frame with < 3 support points must not disturb its neighbours in the chunk; it is counted as failed."""
    W, H = 640, 240
    n = 4
    L, R = make_batch(svb, n, W, H)
    L[2] = 90
    R[2] = 90
    p = svb.default_params(svb.ROBOTICS)
    ctx = svb.Context(p, W, H, chunk=4)
    one = svb.Context(p, W, H, chunk=1)
    try:
        ctx.batch_upload(L, R)
        ctx.batch_run(n, svb.OUT_DISPARITY)
        assert ctx.stats()["frames_failed"] == 1
        for i in (0, 1, 3):
            D1, _ = one.process(L[i], R[i])
            assert np.array_equal(ctx.batch_disparity(i), D1)
        assert (ctx.batch_disparity(2) == 0).all()  # the driver's zero-initialised map, untouched (elas_b200.h: svb_batch_frame_support)
    finally:
        ctx.close()
        one.close()


def test_batch_is_identical_across_stream_modes_lanes_and_vertex_order(svb, golden_meta, monkeypatch):
    """The same frames through the multi-lane pipeline, the single-stream pipeline, six lanes, and with the device
    vertex order switched off (complete host Delaunay path): bit-identical maps and point clouds."""
    W, H, n = 640, 240, 44
    pairs = [svb.synth_pair(300 + i, W, H, i & 1) for i in range(n)]
    Ls = np.stack([p[0] for p in pairs])
    Rs = np.stack([p[1] for p in pairs])
    Q = np.array(golden_meta["Q"])

    def run(single_stream, env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=8)
        for k in env:
            monkeypatch.delenv(k)
        try:
            ctx.set_calibration(Q)
            ctx.set_single_stream(single_stream)
            ctx.batch_upload(Ls, Rs)
            ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
            st = ctx.stats()
            if env.get("SVB_GPU_ORDER") == "0" or env.get("SVB_DELAUNAY_DEVICE") == "0":
                assert st["delaunay_lists_device"] == 0 and st["delaunay_lists_host"] == 2 * n  # every list made by the host stage
            else:
                assert st["delaunay_lists_device"] + st["delaunay_lists_host"] == 2 * n and st["delaunay_lists_device"] >= n  # k_delaunay.cu ran
            return [ctx.batch_disparity(i) for i in range(n)], [ctx.batch_points(i) for i in (0, 7, n - 1)]
        finally:
            ctx.close()

    want = run(False, {})
    for single, env in ((False, {}), (True, {}), (False, {"SVB_LANES": "3"}), (False, {"SVB_GPU_ORDER": "0"}), (False, {"SVB_DELAUNAY_DEVICE": "0"})):
        got = run(single, env)
        assert all(np.array_equal(a, b) for a, b in zip(got[0], want[0])), (single, env)
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got[1], want[1])), (single, env)


def test_batch_frame_without_support_points(svb):
    """A frame that yields fewer than 3 support points (a blank pair under the ROBOTICS preset: no texture, no corners) inside a
    batch: Elas::process would return without touching D (elas.cpp:64-69) and the driver's maps start as zeros, so the batch
    delivers disparity 0 everywhere for THAT frame, reports it per frame, and the neighbours are untouched."""
    W, H = 640, 240
    n = 5
    L, R = make_batch(svb, n, W, H)
    L[2] = 90
    R[2] = 90
    p = svb.default_params(svb.ROBOTICS)
    one = svb.Context(p, W, H, chunk=1)
    ctx = svb.Context(p, W, H, chunk=2)
    try:
        ctx.batch_upload(L, R)
        ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
        nsup = ctx.batch_frame_support(n)
        assert nsup[2] < 3 and (np.delete(nsup, 2) >= 3).all()
        assert ctx.stats()["frames_failed"] == 1
        assert (ctx.batch_disparity(2) == 0).all()
        pts = ctx.batch_points(2)
        assert (pts[:, 2] == 0).all()  # default calibration Q = I: (x, y, d8 = 0)
        for i in (0, 1, 3, 4):
            D1, _ = one.process(L[i], R[i])
            assert np.array_equal(ctx.batch_disparity(i), D1), i
    finally:
        ctx.close()
        one.close()


@pytest.mark.parametrize("setting", ["PIPELINE", "ROBOTICS", "MIDDLEBURY"])
def test_fused_post_chain_equals_stage_by_stage(svb, golden_meta, monkeypatch, setting):
    """k_post_fused.cu (adaptive mean + median + final map + u8 + reprojection in one kernel) against the stage-by-stage kernels
    (SVB_FUSED_POST=0) on the same frames: final maps and point clouds bit for bit, ragged sizes included; all three filter
    combinations (mean + median, mean only, median only)."""
    for (W, H, n) in ((1242, 375, 5), (333, 127, 3), (640, 241, 2)):
        L, R = make_batch(svb, n, W, H)
        p = svb.default_params(getattr(svb, setting), postprocess_only_left=1)
        Q, XR, XT = np.array(golden_meta["Q"]), np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
        outs = []
        for fused in ("1", "0"):
            monkeypatch.setenv("SVB_FUSED_POST", fused)
            ctx = svb.Context(p, W, H, chunk=2)
            try:
                ctx.set_calibration(Q, XR, XT)
                ctx.batch_upload(L, R)
                ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
                outs.append(([ctx.batch_disparity(i) for i in range(n)], [ctx.batch_points(i) for i in range(n)], ctx.stats()["kernel_launches"]))
            finally:
                ctx.close()
        monkeypatch.delenv("SVB_FUSED_POST")
        assert outs[0][2] < outs[1][2]  # fewer launches: the fused kernel really ran
        for i in range(n):
            assert np.array_equal(outs[0][0][i], outs[1][0][i]), (W, H, i)
            assert np.array_equal(outs[0][1][i], outs[1][1][i], equal_nan=True), (W, H, i)


def test_float_disparity_point_cloud(svb, golden_meta, monkeypatch):
    """SVB_OUT_POINTS_FLOATDISP: the float disparity enters Q without the 4x u8 quantisation that clips at 63.75 px.
    Fused and stage-by-stage kernels against the CPU restatement; the u8 drop-in path is unchanged next to it."""
    W, H, n = 640, 240, 3
    L, R = make_batch(svb, n, W, H)
    Q, XR, XT = np.array(golden_meta["Q"]), np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
    p = svb.default_params(svb.PIPELINE)
    for fused in ("1", "0"):
        monkeypatch.setenv("SVB_FUSED_POST", fused)
        ctx = svb.Context(p, W, H, chunk=2)
        try:
            ctx.set_calibration(Q, XR, XT)
            ctx.batch_upload(L, R)
            ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS_FLOATDISP)
            for i in range(n):
                D1 = ctx.batch_disparity(i)
                want = parity.reproject_float_oracle(D1, Q, XR, XT)
                got = ctx.batch_points(i)
                assert np.array_equal(got, want, equal_nan=True), (fused, i)
            with pytest.raises(svb.SvbError):
                ctx.batch_run(n, svb.OUT_POINTS | svb.OUT_POINTS_FLOATDISP)
            ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
            _, want8 = parity.reproject_oracle(ctx.batch_disparity(0), Q, XR, XT)
            assert np.array_equal(ctx.batch_points(0), want8, equal_nan=True)
        finally:
            ctx.close()
    monkeypatch.delenv("SVB_FUSED_POST")


def test_bgra_batch_input_and_device_pointers(svb, golden_meta):
    """(1) a batch of BGRA frames, converted on the device, gives what the gray path gives on cv::cvtColor's output
    (svb_stage_bgra_to_gray is pinned to cv2 elsewhere); (2) svb_batch_device_ptrs hands out the resident results: read back through
    the CUDA runtime (cudaMemcpy on the raw pointers), they are the maps and clouds the download calls return."""
    import ctypes as C

    W, H, n = 640, 240, 5
    L, R = make_batch(svb, n, W, H)
    rng = np.random.default_rng(5)
    # BGRA frames whose gray conversion is NOT trivially one of the channels
    Lb = np.stack([np.stack([L[i], np.roll(L[i], 1, 1), rng.integers(0, 255, (H, W), dtype=np.uint8), np.full((H, W), 255, np.uint8)], -1) for i in range(n)])
    Rb = np.stack([np.stack([R[i], np.roll(R[i], 1, 1), rng.integers(0, 255, (H, W), dtype=np.uint8), np.full((H, W), 255, np.uint8)], -1) for i in range(n)])
    Q = np.array(golden_meta["Q"])
    ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=2)
    try:
        ctx.set_calibration(Q)
        Lg = np.stack([ctx.bgra_to_gray(Lb[i]) for i in range(n)])
        Rg = np.stack([ctx.bgra_to_gray(Rb[i]) for i in range(n)])
        assert not np.array_equal(Lg, L)
        ctx.batch_upload(Lg, Rg)
        ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
        want = [ctx.batch_disparity(i) for i in range(n)]
        ctx.batch_upload_bgra(Lb, Rb)
        ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
        for i in range(n):
            assert np.array_equal(ctx.batch_disparity(i), want[i]), i
        d1, pts, frames, dev = ctx.batch_device_ptrs()
        assert d1 and pts and frames == n and dev >= 0
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        got_D = np.zeros((n, H, W), np.float32)
        got_P = np.zeros((n, H * W, 3), np.float64)
        assert rt.cudaMemcpy(got_D.ctypes.data, d1, got_D.nbytes, 2) == 0  # cudaMemcpyDeviceToHost
        assert rt.cudaMemcpy(got_P.ctypes.data, pts, got_P.nbytes, 2) == 0
        for i in range(n):
            assert np.array_equal(got_D[i], want[i])
            assert np.array_equal(got_P[i], ctx.batch_points(i), equal_nan=True)
        ctx.batch_run(n, svb.OUT_DISPARITY)
        d1, pts, _, _ = ctx.batch_device_ptrs()
        assert d1 and not pts
    finally:
        ctx.close()


def test_owner_map_generations_wrap(svb):
    """The owner maps are not cleared per chunk: their entries carry a 7-bit generation number (svb_internal.h).  One lane is driven
    through more than 127 chunks, so that the number wraps and the maps are cleared in between; frames that repeat an input must
    repeat its result exactly, before and after the wrap, and differently textured frames in between must not leak into them."""
    W, H = 333, 127
    kinds = 3
    n = 140
    L = np.zeros((n, H, W), np.uint8)
    R = np.zeros((n, H, W), np.uint8)
    for i in range(n):
        svb.synth_pair(500 + (i % kinds), W, H, (i % kinds) & 1, L[i], R[i])
    p = svb.default_params(svb.PIPELINE)
    ctx = svb.Context(p, W, H, chunk=1)  # one lane, one frame per chunk: 140 generations
    try:
        ctx.batch_upload(L, R)
        ctx.batch_run(n, svb.OUT_DISPARITY)
        want = [ctx.process(L[k], R[k])[0] for k in range(kinds)]
        for i in range(n):
            assert np.array_equal(ctx.batch_disparity(i), want[i % kinds]), "frame %d" % i
        assert (want[0] >= 0).mean() > 0.5
    finally:
        ctx.close()
