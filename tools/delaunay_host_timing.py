#!/usr/bin/env python
"""Host Delaunay stage on one large list (a 4K frame's 15 000 lattice points): milliseconds per list with the subtrees of one recursion
depth on threads of their own (SVB_DELAUNAY_PAR = 1, 2, 4)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_binding  # noqa: E402
import test_cabi_host as H  # noqa: E402

svb = load_binding().binding
s = H.lattice_support(np.random.default_rng(5), 15000, W=3840, H=2160, dmax=200)
for par in ("1", "2", "4", "8"):
    os.environ["SVB_DELAUNAY_PAR"] = par
    svb.delaunay(s, 0)
    ts = []
    for _ in range(20):
        t = time.perf_counter()
        svb.delaunay(s, 0)
        ts.append(time.perf_counter() - t)
    print("%2s threads: min %.2f ms, median %.2f ms per list" % (par, min(ts) * 1e3, sorted(ts)[10] * 1e3))
