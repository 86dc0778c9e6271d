#!/usr/bin/env python
"""Row-band split timing at 3840x2160 / disparity range 512 (BASELINE.json configs[3]): one frame over 1, 2, 4, 8 GPUs.
   python tools/band_bench.py [--reps 5]      (one process drives all visible GPUs; prints one JSON line per band count)"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--disp_max", type=int, default=511)
    args = ap.parse_args()
    svb = load_package().binding
    nd = svb.device_count()
    W, H = args.width, args.height
    L0, R0 = svb.synth_pair(9, W, H, 0)
    pin = [svb.PinnedArray((H, W), np.uint8), svb.PinnedArray((H, W), np.uint8), svb.PinnedArray((H, W), np.float32),
           svb.PinnedArray((H, W), np.float32)]
    L, R, P1, P2 = (a.array for a in pin)  # pinned host buffers: transfers at full PCIe rate
    L[:] = L0
    R[:] = R0
    p = svb.default_params(svb.MIDDLEBURY, disp_max=args.disp_max)
    base = None
    want = None
    for n in (1, 2, 4, 8):
        if n > nd:
            break
        g = svb.BandGroup(p, W, H, list(range(n)))
        D1, D2 = g.process(L, R, P1, P2)  # warm-up
        if want is None:
            want = (D1.copy(), D2.copy())
        same = bool(np.array_equal(D1, want[0]) and np.array_equal(D2, want[1]))
        ms = []
        for _ in range(args.reps):
            g.process(L, R, P1, P2)
            ms.append(g.stats()["gpu_ms"])
        st = g.stats()
        g.close()
        t = float(np.median(ms))
        base = base or t
        print(json.dumps({"workload": "%dx%d disp %d, one frame, row bands" % (W, H, args.disp_max + 1), "gpus": n, "ms_per_frame": t,
                          "speedup_vs_1": base / t, "identical_to_1_gpu": same, "p2p_MB": st["p2p_bytes"] / 1e6, "p2p_copies": st["p2p_copies"],
                          "peer_links": st["peer_links"], "host_delaunay_ms": st["delaunay_ms"], "support_points": st["support_points"]}))


if __name__ == "__main__":
    main()
