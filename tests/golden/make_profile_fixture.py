#!/usr/bin/env python
"""Writes tests/golden/profile_gray.npz: two of the reference's own `runProfiling` inputs (datasets/profile/*.pgm,
stereo_vision.cu:766-780), so that the GPU parity tests also run on real images of other sizes (900x750, 1344x391).
Run in the build container, where /root/reference exists."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("REFERENCE_ROOT", "/root/reference") + "/datasets/profile"


def pgm(path):
    b = open(path, "rb").read()
    toks, i = [], 0
    while len(toks) < 4:  # magic, width, height, maxval; '#' comment lines (GIMP) are skipped
        while b[i:i + 1].isspace():
            i += 1
        if b[i:i + 1] == b"#":
            while b[i:i + 1] != b"\n":
                i += 1
            continue
        j = i
        while not b[j:j + 1].isspace():
            j += 1
        toks.append(b[i:j])
        i = j
    i += 1
    w, h = int(toks[1]), int(toks[2])
    return np.frombuffer(b[i:i + w * h], np.uint8).reshape(h, w).copy()


out = {}
for name in ("cones", "urban1"):
    out[name + "_L"] = pgm("%s/%s_left.pgm" % (SRC, name))
    out[name + "_R"] = pgm("%s/%s_right.pgm" % (SRC, name))
np.savez_compressed(os.path.join(HERE, "profile_gray.npz"), **out)
print({k: v.shape for k, v in out.items()})
