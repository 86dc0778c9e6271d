"""GPU tests of the sequence driver binary (build/bin/stereo_vision_parallel == lib/stereo_vision_parallel):
KITTI-style directory in, the reference's stdout line formats out (stereo_vision.cu:691,695), the u8 disparity maps
checked against the committed reference outputs; -P 1 profiling mode against the oracle."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

import parity
from conftest import GOLDEN, PKG_DIR
from test_calibration import write_yaml

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

EXE = os.path.join(PKG_DIR, "lib", "stereo_vision_parallel")


def make_sequence(root, kitti_gray):
    for cam in ("image_02", "image_03"):
        os.makedirs(os.path.join(root, cam, "data"))
    for i, f in enumerate((0, 7)):
        for cam, key in (("image_02", "L%d" % f), ("image_03", "R%d" % f)):
            g = kitti_gray[key]
            cv2.imwrite(os.path.join(root, cam, "data", "%010d.png" % i), np.stack([g, g, g], -1))  # 3-channel like KITTI


def test_image_loop(tmp_path, kitti_gray, golden):
    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        case = [c for c in json.load(f)["cases"] if c["name"] == "data/calibration/kitti_2011_09_26.yml" and c["size"] == [1242, 375]
                and c["alpha"] == 0.0][0]
    seq = tmp_path / "seq"
    make_sequence(str(seq), kitti_gray)
    os.makedirs(tmp_path / "data" / "calibration")
    write_yaml(tmp_path / "data" / "calibration" / "kitti_2011_09_26.yml", case)  # the cwd-relative default (stereo_vision.cu:66)
    dump = tmp_path / "dump"
    os.makedirs(dump)
    # test.sh spells the options "-v=1 -s=0 -p=0 -f=1.0" (test.sh:24-30)
    r = subprocess.run([EXE, "-k", str(seq), "-v=1", "-s=0", "-p=0", "-f=1.0", "--dump_dir=" + str(dump)], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    frames = re.findall(r"^\(FPS=([0-9.]+)\) \((\d+), (\d+)\) \(t_t=([0-9.]+), dmap_t=([0-9.]+), pc_t=([0-9.]+)\)$", r.stdout, re.M)
    assert len(frames) == 2 and frames[0][1:3] == ("375", "1242")
    avg = re.search(r"^AVG_FPS=([0-9.]+)$", r.stdout, re.M)
    assert avg and abs(float(avg.group(1)) - np.mean([float(f[0]) for f in frames])) < 1e-3 * float(avg.group(1)) + 1e-3
    assert "Max files = 2" in r.stdout and "Program exitted successfully!" in r.stdout
    Q = np.array(case["Q"]).reshape(4, 4)
    for i, f in enumerate((0, 7)):
        dm = cv2.imread(str(dump / ("%010d_disp.pgm" % i)), cv2.IMREAD_GRAYSCALE)
        want, _ = parity.reproject_oracle(golden["pipeline_%d_D1" % f], Q, np.eye(3), np.zeros(3))
        assert np.array_equal(dm, want)


def test_batch_extension(tmp_path, kitti_gray):
    seq = tmp_path / "seq"
    make_sequence(str(seq), kitti_gray)
    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        case = json.load(f)["cases"][0]
    write_yaml(tmp_path / "k.yml", case)
    r = subprocess.run([EXE, "--kitti_path=" + str(seq), "-p", "0", "-B", "2", "-c", str(tmp_path / "k.yml")], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert re.search(r"^AVG_FPS=[0-9.]+$", r.stdout, re.M) and "(BATCH frames=2)" in r.stdout


def test_scale_factor(tmp_path, kitti_gray, ref):
    """-f 2.0 (test.sh sweeps it): frames are resized like cv::resize does, K1/K2 are divided by the factor and the
    rectification targets the smaller size; the u8 map equals the oracle's on cv2-resized inputs."""
    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        case = [c for c in json.load(f)["cases"] if c["name"] == "data/calibration/kitti_2011_09_26.yml" and c["size"] == [1242, 375]
                and c["alpha"] == 0.0][0]
    seq = tmp_path / "seq"
    make_sequence(str(seq), kitti_gray)
    write_yaml(tmp_path / "k.yml", case)
    dump = tmp_path / "dump"
    os.makedirs(dump)
    r = subprocess.run([EXE, "-k", str(seq), "-p=0", "-f=2.0", "-c", str(tmp_path / "k.yml"), "-o", str(dump)], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "(187, 621)" in r.stdout
    L = cv2.resize(np.stack([kitti_gray["L0"]] * 3, -1), (621, 187))[..., 0]
    R = cv2.resize(np.stack([kitti_gray["R0"]] * 3, -1), (621, 187))[..., 0]
    D1, _, _ = ref.process(ref.pipeline_params(), L, R)
    want = np.clip(np.rint(D1 * np.float32(4.0)), 0, 255).astype(np.uint8)
    got = cv2.imread(str(dump / "0000000000_disp.pgm"), cv2.IMREAD_GRAYSCALE)
    assert np.array_equal(got, want)


def test_extrapolate_point_cloud(tmp_path, kitti_gray, golden_meta):
    """-e 2 (publishPointCloud, stereo_vision.cu:245-265): the u8 map is resized like cv::resize does and the larger map is
    projected with the same Q."""
    import parity

    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        case = [c for c in json.load(f)["cases"] if c["name"] == "data/calibration/kitti_2011_09_26.yml" and c["size"] == [1242, 375]
                and c["alpha"] == 0.0][0]
    seq = tmp_path / "seq"
    make_sequence(str(seq), kitti_gray)
    write_yaml(tmp_path / "k.yml", case)
    dump = tmp_path / "dump"
    os.makedirs(dump)
    r = subprocess.run([EXE, "-k", str(seq), "-p=0", "-e", "2", "-c", str(tmp_path / "k.yml"), "-o", str(dump)], cwd=tmp_path, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    small = cv2.imread(str(dump / "0000000000_disp.pgm"), cv2.IMREAD_GRAYSCALE)
    big = cv2.imread(str(dump / "0000000000_disp_e.pgm"), cv2.IMREAD_GRAYSCALE)
    assert big.shape == (750, 2484) and np.array_equal(big, cv2.resize(small, (2484, 750)))
    pts = np.fromfile(str(dump / "0000000000_points_e.f64"), np.float64).reshape(-1, 3)
    with np.errstate(all="ignore"):
        _, want = parity.reproject_oracle(big.astype(np.float32) / np.float32(4.0), np.array(case["Q"]).reshape(4, 4), case["XR"], case["XT"])
    fin = np.isfinite(want).all(1)
    assert np.array_equal(np.isfinite(pts).all(1), fin) and fin.any()
    # the driver's own stereoRectify agrees with cv2's Q to 1e-13 (tests/test_calibration.py), hence not bit for bit here
    assert np.allclose(pts[fin], want[fin], rtol=1e-9, atol=1e-9)


def test_unsupported_options_fail_loudly(tmp_path):
    r = subprocess.run([EXE, "-k", str(tmp_path), "-p", "0", "-e", "0"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "extrapolate_point_cloud" in r.stderr
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "Usage" in r.stderr


def test_profiling_mode(tmp_path, kitti_gray, ref):
    prof = tmp_path / "datasets" / "profile"
    os.makedirs(prof)
    L, R = kitti_gray["L0"][:200, :600].copy(), kitti_gray["R0"][:200, :600].copy()
    H, W = L.shape
    (prof / "cones_left.pgm").write_bytes(b"P5\n%d %d\n255\n" % (W, H) + L.tobytes())
    (prof / "cones_right.pgm").write_bytes(b"P5\n# CREATOR: GIMP PNM Filter Version 1.1\n%d %d\n255\n" % (W, H) + R.tobytes())
    r = subprocess.run([EXE, "-P", "1"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Processing: datasets/profile/cones_left.pgm, datasets/profile/cones_right.pgm" in r.stdout and "... done!" in r.stdout
    # runProfiling(): ROBOTICS with both maps post-processed, scaled by the common maximum (stereo_vision.cu:727-747)
    D1, D2, _ = ref.process(ref.params(0, postprocess_only_left=0), L, R)
    dmax = max(D1.max(), D2.max())
    for name, D in (("cones_left_disp.pgm", D1), ("cones_right_disp.pgm", D2)):
        got = cv2.imread(str(prof / name), cv2.IMREAD_GRAYSCALE)
        want = np.maximum(255.0 * D.astype(np.float64) / np.float64(dmax), 0.0).astype(np.uint8)
        assert np.array_equal(got, want), name
