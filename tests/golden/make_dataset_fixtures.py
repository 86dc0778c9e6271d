#!/usr/bin/env python
"""Writes the reference's OWN input data as test fixtures, and pins the oracle on all of it.  Run in the build container
(where /root/reference exists); the GPU box only sees the committed files.

  tests/golden/kitti_gray.npz      all 21 stereo pairs of datasets/kitti_mini (RGB PNG -> gray exactly like the reference
                                   driver: sv.py:185-188 BGR->BGRA, stereo_vision.cu:346-347 BGRA->GRAY, via python cv2),
                                   keys L<i> / R<i>, i = 0 .. 20
  tests/golden/profile_gray.npz    all 7 pairs of datasets/profile (the inputs of runProfiling, stereo_vision.cu:699-764):
                                   cones 900x750, aloe 1282x1110, raindeer 1342x1110, urban1-4 1344x391 (P5 PGM; the urban
                                   files carry a '# CREATOR: GIMP' comment line), keys <name>_L / <name>_R
  tests/golden/dataset_digests.json  per pair and preset, from oracle/_ref/libelas_ref.so (= the reference's serial ELAS, strict
                                   IEEE): support points, triangles left/right, valid pixels left/right, sha256 of D1 and D2.
                                   tests/test_oracle_golden.py re-runs the oracle against these digests (CPU); the -m gpu tests
                                   compare the CUDA path with the live oracle on the same pairs.
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref import RefElas  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
PROFILE = ["cones", "aloe", "raindeer", "urban1", "urban2", "urban3", "urban4"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def kitti_gray(i):
    out = []
    for cam in ("image_02", "image_03"):
        im = cv2.imread(os.path.join(REF, "datasets/kitti_mini/%s/data/%010d.png" % (cam, i)))
        im = cv2.cvtColor(im, cv2.COLOR_BGR2BGRA)
        out.append(cv2.cvtColor(im, cv2.COLOR_BGRA2GRAY))
    return out


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


def digest(r, p, L, R):
    t = r.staged(p, L, R)
    return {"support": int(len(t["support"])), "tri1": int(len(t["tri1"])), "tri2": int(len(t["tri2"])),
            "valid1": int((t["D1"] >= 0).sum()), "valid2": int((t["D2"] >= 0).sum()), "D1": sha(t["D1"]), "D2": sha(t["D2"])}


def main():
    r = RefElas()
    kitti, prof, dig = {}, {}, {"oracle_flags": r.flags, "kitti_pipeline": {}, "profile_runprofiling": {}}
    for i in range(21):
        L, R = kitti_gray(i)
        kitti["L%d" % i], kitti["R%d" % i] = L, R
        # the driver's preset (stereo_vision.cu:315-319)
        dig["kitti_pipeline"][str(i)] = digest(r, r.pipeline_params(), L, R)
        print("kitti", i, dig["kitti_pipeline"][str(i)]["support"], dig["kitti_pipeline"][str(i)]["valid1"])
    for name in PROFILE:
        L = pgm("%s/datasets/profile/%s_left.pgm" % (REF, name))
        R = pgm("%s/datasets/profile/%s_right.pgm" % (REF, name))
        prof[name + "_L"], prof[name + "_R"] = L, R
        # runProfiling's parameters: default (ROBOTICS) preset, postprocess_only_left = false (stereo_vision.cu:727-730)
        dig["profile_runprofiling"][name] = dict(digest(r, r.params(0, postprocess_only_left=0), L, R), shape=list(L.shape))
        print(name, L.shape, dig["profile_runprofiling"][name]["support"], dig["profile_runprofiling"][name]["valid1"])
    np.savez_compressed(os.path.join(HERE, "kitti_gray.npz"), **kitti)
    np.savez_compressed(os.path.join(HERE, "profile_gray.npz"), **prof)
    with open(os.path.join(HERE, "dataset_digests.json"), "w") as f:
        json.dump(dig, f, indent=1)


if __name__ == "__main__":
    main()
