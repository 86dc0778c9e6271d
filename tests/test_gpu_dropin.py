"""GPU tests of the drop-in boundary: the library is driven exactly the way the reference's clients drive theirs.

 * stereo_vision/sv.py (ctypes): CDLL, generatePointCloud with 14 positional arguments, restype = ndpointer(float64,
   (width*height, 3)) aliasing the library-owned buffer, BGRA byte strings as inputs (sv.py:164-189).
 * generateDisparityMap() (C++): `Elas::parameters param(Elas::MIDDLEBURY); param.postprocess_only_left = true;
   param.filter_adaptive_mean = true; ElasGPU elas(param); elas.process(...)` (stereo_vision.cu:315-321), compiled here
   with g++ against include/elas.h.
Results are checked against the committed reference outputs (tests/golden) and the CPU restatement of
projectParallel."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from numpy.ctypeslib import ndpointer

import parity
from conftest import GOLDEN, PKG_DIR, ROOT
from test_calibration import write_yaml

pytestmark = pytest.mark.gpu

LIB = os.path.join(PKG_DIR, "lib", "libelas_b200.so")


def kitti_case():
    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        cg = json.load(f)
    # data/kitti_2011_09_26.yml: the copy with a camera -> robot XR and XT = (-4, 0, 1.7) (SURVEY.md component 16)
    return [c for c in cg["cases"] if c["name"] == "data/kitti_2011_09_26.yml" and c["size"] == [1242, 375] and c["alpha"] == 0.0][0], cg


def gray_to_bgra(g):
    return np.ascontiguousarray(np.stack([g, g, g, np.full_like(g, 255)], -1))


SV_CLIENT = r'''
import ctypes, sys, json
import numpy as np
from numpy.ctypeslib import ndpointer
lib_path, yaml_path, npz_path, out_path = sys.argv[1:5]
z = np.load(npz_path)
L, R = z["L0"], z["R0"]
H, W = L.shape
bgra = lambda g: np.ascontiguousarray(np.stack([g, g, g, np.full_like(g, 255)], -1))
# --- what stereo_vision/sv.py does (sv.py:164-189) ---
lib = ctypes.CDLL(lib_path)
lib.generatePointCloud.restype = ndpointer(dtype=ctypes.c_double, shape=(W * H, 3))
lib.generatePointCloud.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_bool, ctypes.c_bool,
                                   ctypes.c_bool, ctypes.c_bool, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
def generate(left, right):
    return lib.generatePointCloud(left.tobytes(), right.tobytes(), yaml_path.encode(), W, H, True, False, False, False, 1, 1, b"cfg", b"weights",
                                  b"classes")
pts = np.array(generate(bgra(L), bgra(R)))
pts_again = np.array(generate(bgra(L), bgra(R)))
pts7 = np.array(generate(bgra(z["L7"]), bgra(z["R7"])))
lib.getColor.restype = ctypes.POINTER(ctypes.c_ubyte)
col = np.ctypeslib.as_array(lib.getColor(), shape=(H, W, 4)).copy()
np.savez(out_path, pts=pts, pts_again=pts_again, pts7=pts7, col=col)
print("client done", flush=True)
lib.clean()          # exits the interpreter with status 0, like the reference (stereo_vision.cu:125)
print("not reached")
'''


def test_sv_py_style_client(tmp_path, golden, kitti_gray):
    case, _ = kitti_case()
    yaml_path = tmp_path / "kitti.yml"
    write_yaml(yaml_path, case)
    out_path = tmp_path / "out.npz"
    script = tmp_path / "client.py"
    script.write_text(SV_CLIENT)
    r = subprocess.run([sys.executable, str(script), LIB, str(yaml_path), os.path.join(GOLDEN, "kitti_gray.npz"), str(out_path)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "client done" in r.stdout and "not reached" not in r.stdout
    assert "Program exitted successfully!" in r.stdout
    # the per-call line the reference prints and its scripts parse (stereo_vision.cu:630)
    lines = [l for l in r.stdout.splitlines() if l.startswith("(FPS=")]
    assert len(lines) == 3 and "(375, 1242)" in lines[0] and "dmap_t=" in lines[0] and "pc_t=" in lines[0]
    z = np.load(out_path)
    Q = np.array(case["Q"]).reshape(4, 4)
    XR, XT = np.array(case["XR"]), np.array(case["XT"])
    for key, gold in (("pts", "pipeline_0_D1"), ("pts7", "pipeline_7_D1")):
        _, want = parity.reproject_oracle(golden[gold], Q, XR, XT)
        got = z[key]
        fin = np.isfinite(want).all(1)
        assert np.array_equal(np.isfinite(got).all(1), fin)
        rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300)
        assert rel.max() <= 1e-4  # north_star: point cloud within 1e-4 relative
        assert rel.max() <= 1e-12
    assert np.array_equal(z["pts"], z["pts_again"], equal_nan=True)
    assert np.array_equal(z["col"], gray_to_bgra(kitti_gray["L7"]))


def test_point_cloud_bgra_stage(svb, golden, kitti_gray, golden_meta):
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    H, W = L.shape
    ctx = svb.Context(svb.default_params(svb.PIPELINE), W, H)
    try:
        Q, XR, XT = np.array(golden_meta["Q"]), np.array(golden_meta["XR"]), np.array(golden_meta["XT"])
        ctx.set_calibration(Q, XR, XT)
        pts, dmap, D1, times = ctx.point_cloud_bgra(gray_to_bgra(L), gray_to_bgra(R))
        assert np.array_equal(D1, golden["pipeline_0_D1"])
        dm_o, pts_o = parity.reproject_oracle(golden["pipeline_0_D1"], Q, XR, XT)
        assert np.array_equal(dmap, dm_o)
        fin = np.isfinite(pts_o).all(1)
        assert np.allclose(pts[fin], pts_o[fin], rtol=1e-12, atol=0)
        assert times[0] > 0 and times[1] > 0
        # a textureless pair: disparity 0 everywhere -> w = 0 -> non-finite points, like the reference
        flat = np.full((H, W, 4), 90, np.uint8)
        pts, dmap, D1, _ = ctx.point_cloud_bgra(flat, flat)
        assert not dmap.any() and not D1.any()
        assert not np.isfinite(pts).all(1).any()
    finally:
        ctx.close()


def test_point_cloud_with_subsampling_reproduces_the_driver(svb, ref, kitti_gray, golden_meta):
    """generateDisparityMap() hands Elas a full-size zeroed float buffer; with subsampling the half-size map lands in its
    first quarter and the WHOLE buffer is converted and projected (stereo_vision.cu:312-324).  Reproduced as is."""
    L, R = kitti_gray["L7"], kitti_gray["R7"]
    H, W = L.shape
    ctx = svb.Context(svb.default_params(svb.PIPELINE, subsampling=1), W, H)
    try:
        Q = np.array(golden_meta["Q"])
        ctx.set_calibration(Q)
        pts, dmap, D1, _ = ctx.point_cloud_bgra(gray_to_bgra(L), gray_to_bgra(R))
        R1, _, _ = ref.process(ref.pipeline_params(subsampling=1), L, R)  # the oracle is given a zeroed W x H buffer as well
        flat = np.zeros(H * W, np.float32)
        n = (H // 2) * (W // 2)
        flat[:n] = R1.reshape(-1)[:n]
        assert np.array_equal(D1.reshape(-1), flat)
        dm_o, pts_o = parity.reproject_oracle(flat.reshape(H, W), Q, np.eye(3), np.zeros(3))
        assert np.array_equal(dmap, dm_o)
        fin = np.isfinite(pts_o).all(1)
        assert fin.any() and np.allclose(pts[fin], pts_o[fin], rtol=1e-12, atol=0)
    finally:
        ctx.close()


def test_resize_matches_cv2(svb):
    """cv::resize INTER_LINEAR (8-bit, fixed point) for the scale factors the reference's test.sh sweeps (0.5 .. 3.0)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for sw, sh in ((1242, 375), (333, 127)):
        img = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8)
        img[: sh // 3] = (np.arange(sw) % 256).astype(np.uint8)[None, :, None]
        for scale in (0.5, 0.75, 1.0, 1.5, 2.0, 2.5, 3.0):
            dsize = (int(sw / scale), int(sh / scale))
            assert np.array_equal(svb.resize_bgra(img, dsize), cv2.resize(img, dsize)), (sw, sh, scale)
            assert np.array_equal(svb.resize_gray(img[..., 1], dsize), cv2.resize(img[..., 1], dsize)), (sw, sh, scale)


def test_bgra_to_gray_matches_cv2(svb):
    with open(os.path.join(GOLDEN, "calib_golden.json")) as f:
        probe = json.load(f)["gray_probe"]
    rng = np.random.default_rng(probe["seed"])
    h, w, _ = probe["shape"]
    bgra = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    ctx = svb.Context(svb.default_params(svb.ROBOTICS), w, h)
    try:
        got = ctx.bgra_to_gray(bgra)
        assert np.array_equal(got.ravel(), np.array(probe["gray"], np.uint8))
        # the closed form SURVEY.md 7.3 records
        b, g, r = (bgra[..., i].astype(np.int64) for i in range(3))
        assert np.array_equal(got, ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8))
    finally:
        ctx.close()


CPP_CLIENT = r'''
// the body of generateDisparityMap() (src/parallel_includes/main/stereo_vision.cu:304-326) against include/elas.h
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "elas.h"
int main(int argc, char **argv) {
    const int W = atoi(argv[1]), H = atoi(argv[2]);
    std::vector<uint8_t> L((size_t)W * H), R((size_t)W * H);
    FILE *f = fopen(argv[3], "rb");
    if (!f || fread(L.data(), 1, L.size(), f) != L.size() || fread(R.data(), 1, R.size(), f) != R.size()) return 2;
    fclose(f);
    const int32_t dims[3] = {W, H, W};
    std::vector<float> D1((size_t)W * H, 0.f), D2((size_t)W * H, 0.f);
    static Elas::parameters param(Elas::MIDDLEBURY);
    param.postprocess_only_left = true;
    param.subsampling = false;
    param.filter_adaptive_mean = true;
    static ElasGPU elas(param);
    elas.process(L.data(), R.data(), D1.data(), D2.data(), dims);
    elas.process(L.data(), R.data(), D1.data(), D2.data(), dims);  // the object is reused frame after frame
    f = fopen(argv[4], "wb");
    fwrite(D1.data(), 4, D1.size(), f);
    fclose(f);
    Elas::parameters rob;  // default = ROBOTICS
    printf("%d %d %g %d\n", rob.ipol_gap_width, (int)rob.add_corners, rob.support_threshold, (int)sizeof(Elas::parameters));
    return 0;
}
'''


def test_cpp_elas_gpu_class(tmp_path, golden, kitti_gray):
    src = tmp_path / "client.cpp"
    src.write_text(CPP_CLIENT)
    exe = tmp_path / "client"
    libdir = os.path.dirname(LIB)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir, "-lelas_b200",
                    "-Wl,-rpath," + libdir], check=True, timeout=300)
    L, R = kitti_gray["L0"], kitti_gray["R0"]
    H, W = L.shape
    inp = tmp_path / "in.bin"
    inp.write_bytes(L.tobytes() + R.tobytes())
    outp = tmp_path / "out.bin"
    r = subprocess.run([str(exe), str(W), str(H), str(inp), str(outp)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    D1 = np.frombuffer(outp.read_bytes(), np.float32).reshape(H, W)
    assert np.array_equal(D1, golden["pipeline_0_D1"])
    assert r.stdout.split()[:3] == ["3", "0", "0.85"]


def test_reproject_u8_any_size(svb, golden_meta):
    """svb_reproject_u8 (no context, any map size) against the CPU restatement of projectParallel (oracle/project_port.c), bit for bit."""
    import parity

    rng = np.random.default_rng(3)
    for W, H in ((2484, 750), (333, 127), (17, 5)):
        dm = rng.integers(0, 256, (H, W), dtype=np.uint8)
        dm[rng.random((H, W)) < 0.2] = 0
        pts = svb.reproject_u8(dm, np.array(golden_meta["Q"]), golden_meta["XR"], golden_meta["XT"])
        with np.errstate(all="ignore"):
            _, want = parity.reproject_oracle(dm.astype(np.float32) / np.float32(4.0), golden_meta["Q"], golden_meta["XR"], golden_meta["XT"])
        assert np.array_equal(pts, want, equal_nan=True), (W, H)
