"""The reference's Python plugin interface for the path -- class `stereo_vision` of stereo_vision/sv.py:154-192 -- as the package
provides it (`elas_b200.sv.stereo_vision`): same constructor signature, same ctypes prototype, same call and teardown behaviour.
CPU part: everything except the numbers (no CUDA device here: the library logs the error and the cloud stays zero -- there is no
CPU path); the numbers are checked on the GPU box (test_stereo_vision_class at the end of this file)."""
import ast
import inspect
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, PKG_DIR, ROOT, load_binding

# sv.py:156-163 (names, order, defaults); so_lib_path's default differs on purpose: the reference searches site-packages at import
REFERENCE_INIT = [("so_lib_path", None), ("width", 1242), ("height", 375), ("defaultCalibFile", True), ("objectTracking", True), ("graphics", False),
                  ("display", False), ("scale", 1), ("pc_extrapolation", 1), ("YOLO_CFG", "src/yolo/yolov4-tiny.cfg"),
                  ("YOLO_WEIGHTS", "src/yolo/yolov4-tiny.weights"), ("YOLO_CLASSES", "src/yolo/classes.txt"),
                  ("CAMERA_CALIBRATION_YAML", "data/calibration/kitti_2011_09_26.yml"), ("subsampling", False)]
# sv.py:180: the ctypes prototype, 14 of generatePointCloud's 16 parameters
REFERENCE_ARGTYPES = ["c_char_p", "c_char_p", "c_char_p", "c_int", "c_int", "c_bool", "c_bool", "c_bool", "c_bool", "c_int", "c_int", "c_char_p",
                      "c_char_p", "c_char_p"]
REFERENCE_SV = "/root/reference/stereo_vision/sv.py"


def mirror():
    load_binding()
    import importlib

    return importlib.import_module("elas_b200.sv")


def test_constructor_signature_is_the_references():
    cls = mirror().stereo_vision
    sig = inspect.signature(cls.__init__)
    got = [(n, p.default) for n, p in sig.parameters.items() if n != "self"]
    assert got == REFERENCE_INIT
    assert list(inspect.signature(cls.generatePointCloud).parameters) == ["self", "left", "right"]
    assert hasattr(cls, "__del__")


@pytest.mark.skipif(not os.path.exists(REFERENCE_SV), reason="the reference tree is only present in the build container")
def test_signature_against_the_reference_source():
    """The table above, read off the reference's own source (the module itself cannot be imported: it raises IndexError at import when
    no stereo_vision*.so is installed, sv.py:149)."""
    tree = ast.parse(open(REFERENCE_SV).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "stereo_vision"][0]
    fns = {f.name: f for f in cls.body if isinstance(f, ast.FunctionDef)}
    assert set(fns) == {"__init__", "generatePointCloud", "__del__"}
    init = fns["__init__"]
    names = [a.arg for a in init.args.args][1:]
    defaults = [ast.literal_eval(d) if not isinstance(d, ast.Name) else None for d in init.args.defaults]  # so_lib_path's default is a module global
    assert list(zip(names, defaults)) == REFERENCE_INIT
    assert [a.arg for a in fns["generatePointCloud"].args.args] == ["self", "left", "right"]
    # the ctypes prototype: 14 argument types (sv.py:180)
    proto = [n for n in ast.walk(init) if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute) and n.targets[0].attr == "argtypes"][0]
    assert [e.attr for e in proto.value.elts] == REFERENCE_ARGTYPES


CLIENT = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import __graft_entry__ as g
g.load_package()
from elas_b200.sv import stereo_vision
W, H = 96, 64
s = stereo_vision(width=W, height=H, objectTracking=False, CAMERA_CALIBRATION_YAML="/nonexistent/calibration.yml")
print("argtypes", ",".join(t.__name__ for t in s.sv.generatePointCloud.argtypes), flush=True)
rt = s.sv.generatePointCloud.restype
print("restype", rt._dtype_, rt._shape_, flush=True)
rng = np.random.default_rng(1)
L = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
pts = s.generatePointCloud(L, L)
assert pts.shape == (W * H, 3) and pts.dtype == np.float64
print("sum", float(np.abs(pts).sum()), flush=True)       # no calibration -> the zeroed buffer, with or without a GPU
pts2 = s.generatePointCloud(L[..., 0], np.dstack([L, L[..., :1]]))  # gray and 4-channel inputs are accepted
assert pts2.ctypes.data == pts.ctypes.data              # the library-owned buffer, aliased (sv.py:167)
for bad in (L[:10], L.astype(np.float32)):
    try:
        s.generatePointCloud(bad, L)
        print("accepted a bad image", flush=True)
    except (ValueError, TypeError):
        pass
print("client done", flush=True)
del s                                                   # __del__ -> clean() -> exit(0), like the reference (sv.py:191-192)
print("not reached", flush=True)
'''


def test_call_and_teardown_behaviour(tmp_path):
    script = tmp_path / "client.py"
    script.write_text(CLIENT)
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout
    assert "sum 0.0" in out and "client done" in out and "accepted a bad image" not in out
    assert "argtypes " + ",".join(REFERENCE_ARGTYPES) in out and "restype float64 (%d, 3)" % (96 * 64) in out  # sv.py:167,180
    assert "Program exitted successfully!" in out and "not reached" not in out
    assert len([l for l in out.splitlines() if l.startswith("(FPS=")]) == 2  # the per-call line of stereo_vision.cu:630
    # SVB_CLEAN_NO_EXIT=1: clean() returns instead of ending the process
    r = subprocess.run([sys.executable, str(script), ROOT], capture_output=True, text=True, timeout=300, env=dict(os.environ, SVB_CLEAN_NO_EXIT="1"))
    assert r.returncode == 0 and "not reached" in r.stdout


# ---- GPU part (this file sorts last: a failure here cannot keep `pytest -x` from running the parity tests before it)
SV_CLASS_CLIENT = r'''
import sys
import numpy as np
root, yaml_path, npz_path, out_path, lib_path = sys.argv[1:6]
sys.path.insert(0, root)
import __graft_entry__ as g
g.load_package()
from elas_b200.sv import stereo_vision             # the reference's `from stereo_vision.sv import stereo_vision`
z = np.load(npz_path)
H, W = z["L0"].shape
bgr = lambda g: np.ascontiguousarray(np.stack([g, g, g], -1))   # what cv2.imread hands the reference's callers
s = stereo_vision(so_lib_path=lib_path, width=W, height=H, objectTracking=False, graphics=False, display=False, CAMERA_CALIBRATION_YAML=yaml_path)
pts = np.array(s.generatePointCloud(bgr(z["L0"]), bgr(z["R0"])))
pts7 = np.array(s.generatePointCloud(bgr(z["L7"]), bgr(z["R7"])))
col = np.array(s.getColor())
np.savez(out_path, pts=pts, pts7=pts7, col=col)
print("client done", flush=True)
del s                                               # __del__ -> clean() -> exit(0) (sv.py:191-192)
print("not reached")
'''


@pytest.mark.gpu
def test_stereo_vision_class(tmp_path, golden, kitti_gray):
    """The package's mirror of the reference's Python plugin class (elas_b200.sv.stereo_vision = stereo_vision/sv.py:154-192) on real
    frames: BGR images in, the aliased double3 cloud out, against the oracle; teardown through __del__ like the reference."""
    import parity
    from test_calibration import write_yaml
    from test_gpu_dropin import gray_to_bgra, kitti_case

    LIB = os.path.join(PKG_DIR, "lib", "libelas_b200.so")
    case, _ = kitti_case()
    yaml_path = tmp_path / "kitti.yml"
    write_yaml(yaml_path, case)
    out_path = tmp_path / "out.npz"
    script = tmp_path / "client.py"
    script.write_text(SV_CLASS_CLIENT)
    r = subprocess.run([sys.executable, str(script), ROOT, str(yaml_path), os.path.join(GOLDEN, "kitti_gray.npz"), str(out_path), LIB],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "client done" in r.stdout and "not reached" not in r.stdout and "Program exitted successfully!" in r.stdout
    z = np.load(out_path)
    Q = np.array(case["Q"]).reshape(4, 4)
    XR, XT = np.array(case["XR"]), np.array(case["XT"])
    for key, gold in (("pts", "pipeline_0_D1"), ("pts7", "pipeline_7_D1")):
        _, want = parity.reproject_oracle(golden[gold], Q, XR, XT)
        got = z[key]
        fin = np.isfinite(want).all(1)
        assert np.array_equal(np.isfinite(got).all(1), fin)
        rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300)
        assert rel.max() <= 1e-4  # north_star: point cloud within 1e-4 relative (Q comes from the library's own stereoRectify)
    assert np.array_equal(z["col"], gray_to_bgra(kitti_gray["L7"]))
