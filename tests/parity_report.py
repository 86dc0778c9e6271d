"""Diagnostic (not a test): prints the stage-by-stage parity table on the GPU box.
   python tests/parity_report.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import GOLDEN, load_binding  # noqa: E402
from oracle.ref import RefElas  # noqa: E402
import parity  # noqa: E402


def show(title, res):
    print("== " + title)
    for k, v in res.items():
        if isinstance(v, dict):
            flag = "OK " if v.get("equal") else "BAD"
            print("   %-18s %s mismatch=%s of %s %s" % (k, flag, v.get("mismatch"), v.get("size"), v.get("shape", "")))
        else:
            print("   %-18s %s" % (k, v))
    sys.stdout.flush()


def main():
    svb = load_binding().binding
    ref = RefElas()
    z = np.load(os.path.join(GOLDEN, "kitti_gray.npz"))
    L, R = z["L0"], z["R0"]
    H, W = L.shape
    print("devices", svb.device_count())
    for pname, setting, p_ref in (("robotics", svb.ROBOTICS, ref.params(0)), ("pipeline", svb.PIPELINE, ref.pipeline_params())):
        p = svb.default_params(setting)
        ctx = svb.Context(p, W, H, chunk=1)
        t0 = time.time()
        res, t, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=True)
        show("kitti0 %s staged (oracle triangles injected)  %.1fs" % (pname, time.time() - t0), res)
        res2, _, _ = parity.staged_parity(ctx, ref, p_ref, L, R, inject=False)
        show("kitti0 %s end-to-end (own Delaunay)" % pname, {k: res2[k] for k in ("tri1", "tri2", "D1raw", "D2raw", "D1", "D2", "D1_mask_equal")})
        show("kitti0 %s isolated stages" % pname, parity.isolated_parity(ctx, ref, p_ref, t))
        ctx.close()
    if "--quick" not in sys.argv:
        Ls, Rs = svb.synth_pair(1)
        p = svb.default_params(svb.PIPELINE)
        ctx = svb.Context(p, W, H, chunk=1)
        res, t, _ = parity.staged_parity(ctx, ref, ref.pipeline_params(), Ls, Rs, inject=False)
        show("synthetic frame 1 pipeline end-to-end", res)
        with open(os.path.join(GOLDEN, "golden_meta.json")) as f:
            meta = json.load(f)
        D1 = t["D1"]
        dm, pts = ctx.reproject(D1, np.array(meta["Q"]), np.array(meta["XR"]), np.array(meta["XT"]))
        dm_o, pts_o = parity.reproject_oracle(D1, meta["Q"], meta["XR"], meta["XT"])
        fin = np.isfinite(pts_o).all(1)
        rel = np.abs(pts[fin] - pts_o[fin]) / np.maximum(np.abs(pts_o[fin]), 1e-12)
        print("== reproject: dmap equal", np.array_equal(dm, dm_o), "finite mask equal", np.array_equal(np.isfinite(pts).all(1), fin), "max rel err",
              rel.max() if rel.size else None)
        ctx.close()


if __name__ == "__main__":
    main()
