"""Generates tests/golden/calib_golden.json: calibration inputs (as numbers) and the outputs of python cv2's
stereoRectify for them, i.e. what the reference driver computes in findRectificationMap()
(src/parallel_includes/main/stereo_vision.cu:368-447: K1/K2 rows 0-1 divided by scale_factor, CALIB_ZERO_DISPARITY,
alpha = 0, newImageSize = calib size).  Run where /root/reference and cv2 exist; the json travels.
cv::stereoRectify is third-party arithmetic the reference does not pin; cv2 4.13 in the build container is the
stand-in oracle."""
import glob
import json
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")


def read_calib(path):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    g = lambda k: fs.getNode(k).mat()
    T = np.array([fs.getNode("T").at(i).real() for i in range(3)], np.float64)
    return g("K1"), g("K2"), g("D1"), g("D2"), g("R"), T, g("XR"), g("XT")


def main():
    cases = []
    files = sorted(glob.glob(os.path.join(REF, "data/calibration/*"))) + [os.path.join(REF, "data/kitti_2011_09_26.yml")]
    for f in files:
        K1, K2, D1, D2, R, T, XR, XT = read_calib(f)
        for size, scale, alpha in (((1242, 375), 1.0, 0.0), ((621, 187), 2.0, 0.0), ((1920, 1080), 1.0, 0.0), ((640, 480), 1.0, 0.0),
                                   ((1242, 375), 1.0, 1.0), ((1242, 375), 1.0, 0.5), ((1242, 375), 1.0, -1.0)):
            K1s, K2s = K1.copy(), K2.copy()
            K1s[:2] /= scale
            K2s[:2] /= scale
            R1, R2, P1, P2, Q, _, _ = cv2.stereoRectify(K1s, D1, K2s, D2, size, R, T, flags=cv2.CALIB_ZERO_DISPARITY, alpha=alpha,
                                                        newImageSize=size)
            cases.append({
                "name": os.path.relpath(f, REF), "size": list(size), "scale_factor": scale, "alpha": alpha,
                "K1": K1.ravel().tolist(), "K2": K2.ravel().tolist(), "D1": D1.ravel().tolist(), "D2": D2.ravel().tolist(),
                "R": R.ravel().tolist(), "T": T.tolist(), "XR": XR.ravel().tolist(), "XT": XT.ravel().tolist(),
                "R1": R1.ravel().tolist(), "R2": R2.ravel().tolist(), "P1": P1.ravel().tolist(), "P2": P2.ravel().tolist(),
                "Q": Q.ravel().tolist(),
            })
    # gray conversion probe: BGRA -> GRAY of cv2 on random pixels (stereo_vision.cu:346-347)
    rng = np.random.default_rng(2024)
    bgra = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    gray = cv2.cvtColor(bgra, cv2.COLOR_BGRA2GRAY)
    out = {"cv2_version": cv2.__version__, "cases": cases, "gray_probe": {"seed": 2024, "shape": [37, 53, 4], "gray": gray.ravel().tolist()}}
    with open(os.path.join(HERE, "calib_golden.json"), "w") as fh:
        json.dump(out, fh)
    print(len(cases), "cases written")


if __name__ == "__main__":
    main()
