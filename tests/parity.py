"""Shared parity machinery: run the CUDA path (through the C-ABI) and the reference oracle on the same inputs and
report, stage by stage, how they differ.  Used by the -m gpu tests and by tests/parity_report.py."""
import numpy as np


def cmp_exact(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return {"equal": False, "shape": (a.shape, b.shape), "mismatch": -1}
    neq = a != b
    if a.dtype.kind == "f":
        neq &= ~(np.isnan(a) & np.isnan(b))
    n = int(neq.sum())
    return {"equal": n == 0, "mismatch": n, "size": int(a.size)}


def half(a, ctx):
    """The oracle is handed full-size buffers; with subsampling it fills the first (W/2)*(H/2) floats (elas.h:157-160)."""
    a = np.asarray(a)
    if a.shape == (ctx.Dh, ctx.Dw):
        return a
    return a.reshape(-1)[: ctx.Dh * ctx.Dw].reshape(ctx.Dh, ctx.Dw)


def staged_parity(ctx, ref, p, L, R, inject=True):
    """Full-pipeline run in tap mode against the oracle's staged run.  With inject=True the oracle's triangle lists
    are injected so that every GPU stage is judged on identical inputs (stage isolation); with inject=False the
    product's own host Delaunay is used (end to end)."""
    t = ref.staged(p, L, R)
    out = {}
    ctx.set_tap_mode(True)
    if inject and "tri1" in t:
        ctx.inject_triangles(0, t["tri1"])
        ctx.inject_triangles(1, t["tri2"])
    else:
        ctx.inject_triangles(0, None)
        ctx.inject_triangles(1, None)
    D1, D2 = ctx.process(L, R)
    names = ["desc1", "desc2", "dcan_raw", "dcan", "support", "tri1", "tri2", "planes1", "planes2", "grid1", "grid2", "D1raw", "D2raw", "D1lr",
             "D2lr", "D1seg", "D1gap"]
    if p.filter_adaptive_mean:
        names.append("D1mean")
    if p.filter_median:
        names.append("D1med")
    for nm in names:
        got = ctx.tap(nm)
        want = t[nm]
        if nm.startswith("D"):
            want = half(want, ctx)
        if nm in ("planes1", "planes2"):
            out[nm] = cmp_exact(got.view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
        else:
            out[nm] = cmp_exact(got, want)
    D1_ref, D2_ref = half(t["D1"], ctx), half(t["D2"], ctx)
    out["D1"] = cmp_exact(D1, D1_ref)
    out["D2"] = cmp_exact(D2, D2_ref)
    # float tolerance statistics for the filtered map (north_star: <= 1e-3 px on >= 99.9 % of valid pixels, same invalid mask)
    v_ref = D1_ref >= 0
    v_got = D1 >= 0
    out["D1_mask_equal"] = bool(np.array_equal(v_ref, v_got))
    both = v_ref & v_got
    if both.any():
        err = np.abs(D1[both] - D1_ref[both])
        out["D1_frac_within_1e-3"] = float((err <= 1e-3).mean())
        out["D1_max_err"] = float(err.max())
    ctx.inject_triangles(0, None)
    ctx.inject_triangles(1, None)
    return out, t, (D1, D2)


def isolated_parity(ctx, ref, p, t):
    """Feed each GPU stage the ORACLE's input for that stage and compare with the oracle's output."""
    out = {}
    out["descriptor"] = None
    raw, fin, pts = ctx.support(t["desc1"], t["desc2"])
    out["support.dcan_raw"] = cmp_exact(raw, t["dcan_raw"])
    out["support.dcan"] = cmp_exact(fin, t["dcan"])
    out["support.list"] = cmp_exact(pts, t["support"])
    for side, nm in ((0, "1"), (1, "2")):
        out["planes" + nm] = cmp_exact(ctx.planes(t["support"], t["tri" + nm]).view(np.uint32), t["planes" + nm].view(np.uint32))
        out["grid" + nm] = cmp_exact(ctx.grid(t["support"], side), t["grid" + nm])
        out["disparity" + nm] = cmp_exact(ctx.disparity(t["support"], t["tri" + nm], t["desc1"], t["desc2"], side), t["D%sraw" % nm])
    a, b = ctx.lr_check(t["D1raw"], t["D2raw"])
    out["lr.D1"] = cmp_exact(a, t["D1lr"])
    out["lr.D2"] = cmp_exact(b, t["D2lr"])
    out["segments"] = cmp_exact(ctx.remove_small_segments(t["D1lr"]), t["D1seg"])
    out["gap"] = cmp_exact(ctx.gap_interpolation(t["D1seg"]), t["D1gap"])
    if "D1mean" in t:
        out["mean"] = cmp_exact(ctx.adaptive_mean(t["D1gap"]), t["D1mean"])
    if "D1med" in t:
        src = t["D1mean"] if "D1mean" in t else t["D1gap"]
        out["median"] = cmp_exact(ctx.median(src), t["D1med"])
    del out["descriptor"]
    return out


def reproject_oracle(D, Q, XR, XT):
    """CPU restatement of generateDisparityMap's tail + projectParallel (stereo_vision.cu:324,188-212), float64: the u8 conversion in
    numpy, the projection by oracle/project_port.c, which fuses the multiply-adds the reference's own nvcc build fuses (pinned bit for
    bit to the reference kernel on the GPU box, test_reproject_against_the_reference_kernel)."""
    from oracle.ref import port_project

    H, W = D.shape
    d8 = np.clip(np.rint(D.astype(np.float32) * np.float32(4.0)), 0, 255).astype(np.uint8)
    return d8, port_project(d8.astype(np.float64), H, W, Q, XR, XT)


def reproject_float_oracle(D, Q, XR, XT):
    """SVB_OUT_POINTS_FLOATDISP: projectParallel's arithmetic (stereo_vision.cu:188-212, as above) on the filtered float disparity itself
    (invalid pixels as 0) in place of the u8 value -- new relative to the reference, so the port is the oracle."""
    from oracle.ref import port_project

    H, W = D.shape
    d = np.maximum(D.astype(np.float32), np.float32(0)).astype(np.float64)
    return port_project(d, H, W, Q, XR, XT)


def reproject_numpy(d, Q, XR, XT):
    """projectParallel's formula as written in the source (stereo_vision.cu:188-212), every product and sum rounded on its own: what
    the kernel would compute WITHOUT nvcc's contraction.  Only a cross-check of the port (they agree to a few ulps of the larger
    terms); d: (H, W) values that enter Q."""
    H, W = d.shape
    x = np.tile(np.arange(W, dtype=np.float64), H)
    y = np.repeat(np.arange(H, dtype=np.float64), W)
    d = np.asarray(d, np.float64).reshape(-1)
    Q = np.asarray(Q, np.float64)
    pos = [Q[j, 0] * x + Q[j, 1] * y + Q[j, 2] * d + Q[j, 3] for j in range(4)]
    with np.errstate(divide="ignore", invalid="ignore"):
        X, Y, Z = pos[0] / pos[3], pos[1] / pos[3], pos[2] / pos[3]
        XR = np.asarray(XR, np.float64).reshape(3, 3)
        XT = np.asarray(XT, np.float64).reshape(3)
        return np.stack([XR[j, 0] * X + XR[j, 1] * Y + XR[j, 2] * Z + XT[j] for j in range(3)], 1)
