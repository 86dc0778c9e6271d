"""Generates the committed golden fixtures from the REFERENCE ITSELF (run where /root/reference exists).

Inputs : datasets/kitti_mini frames (RGB PNG) -> gray exactly like the reference driver does it
         (sv.py:185-188 BGR->BGRA, stereo_vision.cu:346-347 BGRA->GRAY), via python cv2.
Outputs: (tests/golden/kitti_gray.npz, the gray stereo pairs, is written by make_dataset_fixtures.py: all 21 pairs)
         tests/golden/kitti_golden.npz   outputs of oracle/_ref/libelas_ref.so (= the reference's serial ELAS,
                                         strict-IEEE build) for those pairs: support points, both triangle lists,
                                         the final left disparity (float16-exactness checked, else float32) and a
                                         sha256 of every stage tap, for the ROBOTICS preset and the pipeline preset
         tests/golden/q_kitti.json       Q from cv2.stereoRectify with the arguments of stereo_vision.cu:447 and
                                         the gray-conversion probe used by the boundary tests
The reference's own tests hold no known-answer vectors for this path (SURVEY.md finding 6), so these fixtures,
produced by running the reference's code here, are what pins the oracle.
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
from oracle.ref import RefElas, ROBOTICS  # noqa: E402

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
FRAMES = [0, 7]
TAPS = ["desc1", "desc2", "dcan_raw", "dcan_incon", "dcan_redv", "dcan", "support", "tri1", "tri2", "planes1", "planes2", "grid1", "grid2",
        "D1raw", "D2raw", "D1lr", "D2lr", "D1seg", "D1gap", "D1mean", "D1med", "D1", "D2"]


def load_gray(i):
    out = []
    for cam in ("image_02", "image_03"):
        im = cv2.imread(os.path.join(REF, "datasets/kitti_mini/%s/data/%010d.png" % (cam, i)))
        im = cv2.cvtColor(im, cv2.COLOR_BGR2BGRA)
        out.append(cv2.cvtColor(im, cv2.COLOR_BGRA2GRAY))
    return out


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def read_calib(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    g = lambda k: fs.getNode(k).mat()
    T = np.array([fs.getNode("T").at(i).real() for i in range(3)], np.float64)
    return g("K1"), g("K2"), g("D1"), g("D2"), g("R"), T, g("XR"), g("XT")


def main():
    r = RefElas()
    gray = {}
    gold = {}
    hashes = {}
    for i in FRAMES:
        L, R = load_gray(i)
        gray["L%d" % i] = L
        gray["R%d" % i] = R
        for pname, p in (("robotics", r.params(ROBOTICS)), ("pipeline", r.pipeline_params())):
            t = r.staged(p, L, R)
            D1p, D2p, _ = r.process(p, L, R)
            assert np.array_equal(D1p, t["D1"]) and np.array_equal(D2p, t["D2"]), "staged taps disagree with Elas::process"
            key = "%s_%d" % (pname, i)
            hashes[key] = {k: sha(t[k]) for k in TAPS if k in t}
            gold[key + "_support"] = t["support"]
            gold[key + "_tri1"] = t["tri1"]
            gold[key + "_tri2"] = t["tri2"]
            gold[key + "_D1"] = t["D1"]
            gold[key + "_D1raw"] = t["D1raw"].astype(np.int16)  # integers, -1 or -10
            gold[key + "_D2raw"] = t["D2raw"].astype(np.int16)
            print(key, "support", len(t["support"]), "tri", len(t["tri1"]), len(t["tri2"]), "valid", int((t["D1"] >= 0).sum()))
    np.savez_compressed(os.path.join(HERE, "kitti_golden.npz"), **gold)

    # calibration -> Q exactly as findRectificationMap does it (stereo_vision.cu:447), scale_factor 1
    K1, K2, D1, D2, Rm, T, XR, XT = read_calib(os.path.join(REF, "data/calibration/kitti_2011_09_26.yml"))
    size = (1242, 375)
    R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K1, D1, K2, D2, size, Rm, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=0, newImageSize=size)
    rng = np.random.default_rng(123)
    bgra = rng.integers(0, 256, (64, 64, 4), dtype=np.uint8)
    meta = {
        "oracle_flags": r.flags,
        "cv2_version": cv2.__version__,
        "frames": FRAMES,
        "hashes": hashes,
        "Q": Q.tolist(),
        "R1": R1.tolist(),
        "P1": P1.tolist(),
        "P2": P2.tolist(),
        "XR": XR.tolist(),
        "XT": XT.reshape(-1).tolist(),
        "gray_probe_seed": 123,
        "gray_probe_sha": sha(cv2.cvtColor(bgra, cv2.COLOR_BGRA2GRAY)),
    }
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("Q =", Q)


if __name__ == "__main__":
    main()
