#!/usr/bin/env python
"""Small end-to-end cases that touch every code path of the library once (a quick manual check on a GPU box).
Sizes are tiny on purpose:
single-frame process (both presets), subsampling, batch pipeline with reprojection, BGRA point cloud, row-band split."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

svb = load_package().binding
W, H = 320, 121
L, R = svb.synth_pair(3, W, H, 1)
for setting, over in ((svb.ROBOTICS, {}), (svb.PIPELINE, {}), (svb.MIDDLEBURY, {"subsampling": 1}), (svb.MIDDLEBURY, {"disp_max": 63})):
    ctx = svb.Context(svb.default_params(setting, **over), W, H)
    D1, D2 = ctx.process(L, R)
    print("process", setting, over, "valid", float((D1 >= 0).mean()))
    ctx.close()
n = 5
Ls = np.stack([svb.synth_pair(10 + i, W, H, i & 1)[0] for i in range(n)])
Rs = np.stack([svb.synth_pair(10 + i, W, H, i & 1)[1] for i in range(n)])
ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H, chunk=2)
ctx.set_calibration(np.array([[1, 0, 0, -160.0], [0, 1, 0, -60.0], [0, 0, 0, 300.0], [0, 0, 2.0, 0]]))
ctx.batch_upload(Ls, Rs)
ctx.batch_run(n, svb.OUT_DISPARITY | svb.OUT_POINTS)
print("batch frames", ctx.stats()["frames"], "valid", float((ctx.batch_disparity(n - 1) >= 0).mean()))
bgra = np.ascontiguousarray(np.stack([L, L, L, np.full_like(L, 255)], -1))
bgra_r = np.ascontiguousarray(np.stack([R, R, R, np.full_like(R, 255)], -1))
pts, dmap, D1, _ = ctx.point_cloud_bgra(bgra, bgra_r)
print("point cloud finite", float(np.isfinite(pts).all(1).mean()))
ctx.close()
g = svb.BandGroup(svb.default_params(svb.ROBOTICS), W, H, [0, 0, 0])
D1, _ = g.process(L, R)
print("bands valid", float((D1 >= 0).mean()))
g.close()
print("sanitize_case done")
