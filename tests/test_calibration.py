"""CPU tests of the calibration boundary (no OpenCV in the product): the YAML-subset reader and the restatement of
cv::stereoRectify against python cv2 4.13's outputs for every calibration file of the reference at several sizes,
scale factors and alphas (tests/golden/calib_golden.json, written by tests/golden/make_calib_golden.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


@pytest.fixture(scope="module")
def calib_golden():
    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        return json.load(f)


def fmt_matrix(name, vals, rows, cols, per_line=3):
    vals = ["%.16e" % v for v in vals]
    lines = [", ".join(vals[i:i + per_line]) for i in range(0, len(vals), per_line)]
    return "%s: !!opencv-matrix\n   rows: %d\n   cols: %d\n   dt: d\n   data: [ %s ]\n" % (name, rows, cols, ",\n       ".join(lines))


def write_yaml(path, c, with_x=True):
    """OpenCV FileStorage YAML 1.0 in the layout of the reference's data/calibration/*.yml."""
    s = "%YAML:1.0\n"
    s += fmt_matrix("K1", c["K1"], 3, 3) + fmt_matrix("K2", c["K2"], 3, 3)
    s += fmt_matrix("D1", c["D1"], 1, len(c["D1"]), 5) + fmt_matrix("D2", c["D2"], 1, len(c["D2"]), 2)
    s += fmt_matrix("R", c["R"], 3, 3)
    s += "T: [ %s ]\n" % ", ".join("%.16e" % v for v in c["T"])
    if with_x:
        s += fmt_matrix("XR", c["XR"], 3, 3) + fmt_matrix("XT", c["XT"], 3, 1, 1)
    with open(path, "w") as f:
        f.write(s)


def test_yaml_reader_roundtrip(svb, calib_golden, tmp_path):
    c = calib_golden["cases"][0]
    p = tmp_path / "calib.yml"
    write_yaml(p, c)
    cal = svb.load_calibration(p)
    for key in ("K1", "K2", "R", "T", "XR", "XT"):
        assert np.array_equal(np.array(getattr(cal, key)), np.array(c[key])), key
    assert cal.n_d1 == len(c["D1"]) and cal.n_d2 == len(c["D2"])
    assert np.array_equal(np.array(cal.D1)[: cal.n_d1], np.array(c["D1"]))
    # XR / XT absent -> identity / zero
    write_yaml(p, c, with_x=False)
    cal = svb.load_calibration(p)
    assert np.array_equal(np.array(cal.XR).reshape(3, 3), np.eye(3)) and not np.array(cal.XT).any()


def test_yaml_reader_errors(svb, tmp_path):
    with pytest.raises(svb.SvbError):
        svb.load_calibration(tmp_path / "does_not_exist.yml")
    p = tmp_path / "broken.yml"
    p.write_text("%YAML:1.0\nK1: !!opencv-matrix\n   rows: 3\n   cols: 3\n   dt: d\n   data: [ 1, 2, 3 ]\n")
    with pytest.raises(svb.SvbError) as e:
        svb.load_calibration(p)
    assert "K1" in str(e.value)


def test_stereo_rectify_matches_cv2(svb, calib_golden, tmp_path):
    assert len(calib_golden["cases"]) >= 50
    worst = 0.0
    for i, c in enumerate(calib_golden["cases"]):
        p = tmp_path / ("c%d.yml" % i)
        write_yaml(p, c)
        cal = svb.load_calibration(p)
        size = tuple(c["size"])
        out = svb.stereo_rectify(cal, size, size, c["scale_factor"], c["alpha"])
        for key in ("R1", "R2", "P1", "P2", "Q"):
            want = np.array(c[key]).reshape(out[key].shape)
            err = np.abs(out[key] - want).max() / max(np.abs(want).max(), 1e-300)
            worst = max(worst, err)
            assert err <= 1e-12, (c["name"], size, c["alpha"], key, err)
    assert worst <= 1e-12


def test_kitti_q_is_the_surveys(svb, calib_golden, golden_meta, tmp_path):
    """SURVEY.md 8a row 19 / tests/golden/golden_meta.json: Q for kitti_2011_09_26.yml at 1242x375."""
    c = [x for x in calib_golden["cases"] if x["name"].endswith("calibration/kitti_2011_09_26.yml") and x["size"] == [1242, 375]
         and x["alpha"] == 0.0][0]
    p = tmp_path / "k.yml"
    write_yaml(p, c)
    out = svb.stereo_rectify(svb.load_calibration(p), (1242, 375))
    assert np.allclose(out["Q"], np.array(golden_meta["Q"]), rtol=1e-13, atol=1e-13)
    assert abs(out["Q"][0, 3] + 738.7995529175) < 1e-6 and abs(out["Q"][3, 2] - 1.861616069957) < 1e-9


def test_yaml_reader_survives_corrupt_files(svb, calib_golden, tmp_path):
    """Truncations, byte flips, absurd dimensions: svb_calib_load_yaml returns a calibration or an error, never crashes; a matrix
    that claims more entries than the structure holds is refused (run under ASan / UBSan as well when the reader changes)."""
    rng = np.random.default_rng(3)
    c = calib_golden["cases"][0]
    p = tmp_path / "good.yml"
    write_yaml(p, c)
    good = p.read_bytes()
    q = tmp_path / "fuzz.yml"
    n_err = n_ok = 0
    for cut in range(0, len(good), 11):
        q.write_bytes(good[:cut])
        try:
            svb.load_calibration(q)
            n_ok += 1
        except svb.SvbError:
            n_err += 1
    for _ in range(300):
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        q.write_bytes(bytes(b))
        try:
            svb.load_calibration(q)
            n_ok += 1
        except svb.SvbError:
            n_err += 1
    assert n_err > 20 and n_ok > 0
    many = ", ".join(["1.0"] * 5000)
    for text in ("%YAML:1.0\nD1: !!opencv-matrix\n   rows: 1\n   cols: 5000\n   dt: d\n   data: [ " + many + " ]\n",
                 "%YAML:1.0\nK1: !!opencv-matrix\n   rows: 2147483647\n   cols: 2147483647\n   dt: d\n   data: [ 1 ]\n",
                 "%YAML:1.0\nK1: !!opencv-matrix\n   rows: -3\n   cols: 3\n   dt: d\n   data: [ " + many + " ]\n",
                 "%YAML:1.0\nT: [ " + many + " ]\n", "K1: [", "", "\x00" * 64):
        q.write_bytes(text.encode())
        with pytest.raises(svb.SvbError):
            svb.load_calibration(q)
