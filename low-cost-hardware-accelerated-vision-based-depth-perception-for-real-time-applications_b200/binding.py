"""ctypes binding of lib/libelas_b200.so (the C-ABI of include/elas_b200.h).

This is the Python-side host mirror used by tests/ and bench.py.  It contains no arithmetic of its own: every
method forwards to the CUDA library.  If the library is missing, importing `load()` raises -- there is no CPU
fallback anywhere in the product path.
"""
import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# SVB_LIB_DIR selects another build of the same library (lib_guard: the bounds-asserting build, csrc/Makefile GUARD=1)
LIB_PATH = os.path.join(os.environ.get("SVB_LIB_DIR") or os.path.join(PKG_DIR, "lib"), "libelas_b200.so")

ROBOTICS, MIDDLEBURY, PIPELINE = 0, 1, 2
OUT_DISPARITY, OUT_POINTS, OUT_POINTS_FLOATDISP = 1, 2, 4
ERR_FEW_SUPPORT = -5

STAGE_NAMES = None


class Params(C.Structure):
    """svb_params == POD mirror of Elas::parameters (src/parallel_includes/elas/elas.h:58-83)."""

    _fields_ = [
        ("disp_min", C.c_int32),
        ("disp_max", C.c_int32),
        ("support_threshold", C.c_float),
        ("support_texture", C.c_int32),
        ("candidate_stepsize", C.c_int32),
        ("incon_window_size", C.c_int32),
        ("incon_threshold", C.c_int32),
        ("incon_min_support", C.c_int32),
        ("add_corners", C.c_int32),
        ("grid_size", C.c_int32),
        ("beta", C.c_float),
        ("gamma", C.c_float),
        ("sigma", C.c_float),
        ("sradius", C.c_float),
        ("match_texture", C.c_int32),
        ("lr_threshold", C.c_int32),
        ("speckle_sim_threshold", C.c_float),
        ("speckle_size", C.c_int32),
        ("ipol_gap_width", C.c_int32),
        ("filter_median", C.c_int32),
        ("filter_adaptive_mean", C.c_int32),
        ("postprocess_only_left", C.c_int32),
        ("subsampling", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("gpu_ms_total", C.c_double),
        ("delaunay_ms_total", C.c_double),
        ("delaunay_ms_wall", C.c_double),
        ("kernel_launches", C.c_int64),
        ("support_points", C.c_int64),
        ("triangles", C.c_int64),
        ("frames", C.c_int64),
        ("frames_failed", C.c_int64),
        ("stage_ms", C.c_double * 24),
        ("delaunay_lists_device", C.c_int64),
        ("delaunay_lists_host", C.c_int64),
    ]


class Calibration(C.Structure):
    """svb_calibration: K1 K2 D1 D2 R T XR XT of an OpenCV-YAML calibration file."""

    _fields_ = [("K1", C.c_double * 9), ("K2", C.c_double * 9), ("D1", C.c_double * 14), ("D2", C.c_double * 14), ("n_d1", C.c_int32),
                ("n_d2", C.c_int32), ("R", C.c_double * 9), ("T", C.c_double * 3), ("XR", C.c_double * 9), ("XT", C.c_double * 3)]


class BandStats(C.Structure):
    _fields_ = [("gpu_ms", C.c_double), ("wall_ms", C.c_double), ("delaunay_ms", C.c_double), ("p2p_bytes", C.c_int64), ("p2p_copies", C.c_int64),
                ("peer_links", C.c_int32), ("bands", C.c_int32), ("support_points", C.c_int64), ("triangles", C.c_int64)]


class SvbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("svb error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load libelas_b200.so; raise loudly if it has not been built (no fallback)."""
    global _lib, STAGE_NAMES
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(LIB_PATH + " is missing: run __graft_entry__.build() (make -C csrc). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.svb_last_error.restype = C.c_char_p
    lib.svb_version.restype = C.c_char_p
    lib.svb_stage_name.restype = C.c_char_p
    lib.svb_create.restype = C.c_void_p
    lib.svb_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.svb_destroy.argtypes = [C.c_void_p]
    lib.svb_band_create.restype = C.c_void_p
    lib.svb_band_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.svb_band_destroy.argtypes = [C.c_void_p]
    lib.svb_tap.restype = C.c_int64
    lib.svb_tap.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    lib.svb_host_alloc.restype = C.c_void_p
    lib.svb_host_alloc.argtypes = [C.c_size_t]
    lib.svb_host_free.argtypes = [C.c_void_p]
    vp = C.c_void_p
    for name, args in {
        "svb_set_mean_mode": [vp, C.c_int],
        "svb_set_delaunay_threads": [vp, C.c_int],
        "svb_set_stage_timing": [vp, C.c_int],
        "svb_set_single_stream": [vp, C.c_int],
        "svb_set_eval_counting": [vp, C.c_int],
        "svb_get_eval_counts": [vp, vp, vp],
        "svb_set_tap_mode": [vp, C.c_int],
        "svb_process": [vp, vp, vp, C.c_int, vp, vp],
        "svb_inject_triangles": [vp, C.c_int, vp, C.c_int],
        "svb_stage_descriptor": [vp, vp, C.c_int, vp],
        "svb_stage_support": [vp, vp, vp, vp, vp, vp, C.c_int, vp],
        "svb_stage_delaunay": [vp, C.c_int, C.c_int, vp, C.c_int, vp],
        "svb_stage_delaunay_pipeline": [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp, vp],
        "svb_stage_delaunay_ordered": [vp, C.c_int, C.c_int, vp, vp, C.c_int, vp],
        "svb_stage_delaunay_levels": [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp],
        "svb_stage_planes": [vp, vp, C.c_int, vp, C.c_int, vp],
        "svb_stage_grid": [vp, vp, C.c_int, C.c_int, vp],
        "svb_stage_disparity": [vp, vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, vp],
        "svb_stage_lr_check": [vp, vp, vp],
        "svb_stage_remove_small_segments": [vp, vp],
        "svb_stage_gap_interpolation": [vp, vp],
        "svb_stage_adaptive_mean": [vp, vp],
        "svb_stage_median": [vp, vp],
        "svb_stage_reproject": [vp, vp, vp, vp, vp, vp, vp],
        "svb_set_calibration": [vp, vp, vp, vp],
        "svb_batch_upload": [vp, vp, vp, C.c_int],
        "svb_batch_upload_bgra": [vp, vp, vp, C.c_int],
        "svb_batch_device_ptrs": [vp, vp, vp, vp, vp],
        "svb_batch_run": [vp, C.c_int, C.c_int],
        "svb_batch_frame_support": [vp, vp, C.c_int],
        "svb_batch_download_disparity": [vp, C.c_int, vp],
        "svb_batch_download_points": [vp, C.c_int, vp],
        "svb_batch_run_host": [vp, vp, vp, C.c_int, C.c_int, vp, vp],
        "svb_get_stats": [vp, C.POINTER(Stats)],
        "svb_default_params": [C.c_int, C.POINTER(Params)],
        "svb_synth_pair": [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp],
        "svb_point_cloud_bgra": [vp, vp, vp, vp, vp, vp, vp],
        "svb_stage_bgra_to_gray": [vp, vp, vp],
        "svb_band_process": [vp, vp, vp, C.c_int, vp, vp],
        "svb_band_get_stats": [vp, C.POINTER(BandStats)],
        "svb_resize_bgra": [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int],
        "svb_resize_gray": [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int],
        "svb_reproject_u8": [vp, C.c_int, C.c_int, vp, vp, vp, vp],
        "svb_image_read": [C.c_char_p, vp, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "svb_calib_load_yaml": [C.c_char_p, C.POINTER(Calibration)],
        "svb_stereo_rectify": [C.POINTER(Calibration), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, vp, vp, vp],
    }.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    STAGE_NAMES = []
    i = 0
    while True:
        s = lib.svb_stage_name(i).decode()
        if not s:
            break
        STAGE_NAMES.append(s)
        i += 1
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(setting=ROBOTICS, **over):
    p = Params()
    load().svb_default_params(setting, C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    return p


def device_count():
    return load().svb_device_count()


def delaunay(support, right):
    """Host stage on its own (no GPU needed): Elas::computeDelaunayTriangulation (elas.cpp:442-501)."""
    lib = load()
    support = np.ascontiguousarray(support, np.int32)
    n = len(support)
    cap = 2 * n + 16
    tri = np.zeros((cap, 3), np.int32)
    m = C.c_int(0)
    rc = lib.svb_stage_delaunay(_ptr(support), n, int(right), _ptr(tri), cap, C.byref(m))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return tri[: m.value].copy()


def delaunay_ordered(support, right, order):
    """Host half of the pipeline's Delaunay stage (no GPU needed): the recursion on a given vertex order."""
    lib = load()
    support = np.ascontiguousarray(support, np.int32)
    order = np.ascontiguousarray(order, np.int32)
    n = len(support)
    cap = 2 * n + 16
    tri = np.zeros((cap, 3), np.int32)
    m = C.c_int(0)
    rc = lib.svb_stage_delaunay_ordered(_ptr(support), n, int(right), _ptr(order), _ptr(tri), cap, C.byref(m))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return tri[: m.value].copy()


def delaunay_levels(support, right, order, host_levels):
    """The device's share of the Delaunay stage restated on the host (level-synchronous, 16-bit records) + host finish."""
    lib = load()
    support = np.ascontiguousarray(support, np.int32)
    order = np.ascontiguousarray(order, np.int32)
    n = len(support)
    cap = 2 * n + 16
    tri = np.zeros((cap, 3), np.int32)
    m = C.c_int(0)
    rc = lib.svb_stage_delaunay_levels(_ptr(support), n, int(right), _ptr(order), int(host_levels), _ptr(tri), cap, C.byref(m))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return tri[: m.value].copy()


def image_read(path):
    """PNG / PGM reader of the sequence driver (replaces cv::imread / loadPGM): HxWx4 BGRA or HxWx1 gray, uint8."""
    lib = load()
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    rc = lib.svb_image_read(str(path).encode(), None, 0, C.byref(w), C.byref(h), C.byref(c))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    buf = np.zeros((h.value, w.value, c.value), np.uint8)
    rc = lib.svb_image_read(str(path).encode(), _ptr(buf), buf.size, C.byref(w), C.byref(h), C.byref(c))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return buf


def resize_bgra(img, dsize):
    """cv::resize(img, dsize) with INTER_LINEAR on 8-bit BGRA, on the GPU (stereo_vision.cu:599-600,665,676)."""
    lib = load()
    img = np.ascontiguousarray(img, np.uint8)
    assert img.ndim == 3 and img.shape[2] == 4
    out = np.zeros((dsize[1], dsize[0], 4), np.uint8)
    rc = lib.svb_resize_bgra(_ptr(img), img.shape[1], img.shape[0], _ptr(out), dsize[0], dsize[1])
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return out


def resize_gray(img, dsize):
    """cv::resize(img, dsize) with INTER_LINEAR on a single-channel 8-bit image (publishPointCloud's map, stereo_vision.cu:249)."""
    lib = load()
    img = np.ascontiguousarray(img, np.uint8)
    assert img.ndim == 2
    out = np.zeros((dsize[1], dsize[0]), np.uint8)
    rc = lib.svb_resize_gray(_ptr(img), img.shape[1], img.shape[0], _ptr(out), dsize[0], dsize[1])
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return out


def reproject_u8(dmap, Q, XR=None, XT=None):
    """projectParallel on a u8 disparity map of any size (stereo_vision.cu:188-212,245-265)."""
    lib = load()
    dmap = np.ascontiguousarray(dmap, np.uint8)
    H, W = dmap.shape
    Q = np.ascontiguousarray(Q, np.float64).reshape(16)
    XR = None if XR is None else np.ascontiguousarray(XR, np.float64).reshape(9)
    XT = None if XT is None else np.ascontiguousarray(XT, np.float64).reshape(3)
    pts = np.zeros((H * W, 3), np.float64)
    rc = lib.svb_reproject_u8(_ptr(dmap), W, H, _ptr(Q), None if XR is None else _ptr(XR), None if XT is None else _ptr(XT), _ptr(pts))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return pts


def load_calibration(path):
    """K1 K2 D1 D2 R T XR XT from an OpenCV-YAML file (replaces cv::FileStorage, stereo_vision.cu:536-545)."""
    lib = load()
    cal = Calibration()
    rc = lib.svb_calib_load_yaml(str(path).encode(), C.byref(cal))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return cal


def stereo_rectify(cal, calib_size, new_size=None, scale_factor=1.0, alpha=0.0):
    """findRectificationMap() without OpenCV (stereo_vision.cu:368-447): returns dict R1 R2 P1 P2 Q."""
    lib = load()
    new_size = new_size or calib_size
    out = {"R1": np.zeros((3, 3)), "R2": np.zeros((3, 3)), "P1": np.zeros((3, 4)), "P2": np.zeros((3, 4)), "Q": np.zeros((4, 4))}
    rc = lib.svb_stereo_rectify(C.byref(cal), calib_size[0], calib_size[1], new_size[0], new_size[1], float(scale_factor), float(alpha),
                                _ptr(out["R1"]), _ptr(out["R2"]), _ptr(out["P1"]), _ptr(out["P2"]), _ptr(out["Q"]))
    if rc != 0:
        raise SvbError(rc, lib.svb_last_error().decode())
    return out


def synth_pair(frame_index, W=1242, H=375, slanted=0, left=None, right=None):
    """Deterministic synthetic stereo pair (host-side input generator)."""
    lib = load()
    if left is None:
        left = np.zeros((H, W), np.uint8)
    if right is None:
        right = np.zeros((H, W), np.uint8)
    rc = lib.svb_synth_pair(int(frame_index), W, H, int(slanted), _ptr(left), _ptr(right))
    if rc != 0:
        raise SvbError(rc, "svb_synth_pair")
    return left, right


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (for the e2e path at full PCIe rate)."""

    def __init__(self, shape, dtype):
        lib = load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib.svb_host_alloc(self.nbytes)
        if not self.ptr:
            raise SvbError(-2, lib.svb_last_error().decode())
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            load().svb_host_free(self.ptr)
            self.ptr = None


class BandGroup:
    """One frame split into row bands over several GPUs (svb_band_*): Elas::process semantics, P2P halo exchange."""

    def __init__(self, params, width, height, devices):
        self.lib = load()
        self.W, self.H = width, height
        arr = (C.c_int * len(devices))(*devices)
        self.h = self.lib.svb_band_create(C.byref(params), width, height, arr, len(devices))
        if not self.h:
            raise SvbError(-1, self.lib.svb_last_error().decode())

    def process(self, I1, I2, D1=None, D2=None):
        """D1 / D2 may be preallocated (e.g. PinnedArray views, so that the transfers run at full PCIe rate)."""
        I1 = np.ascontiguousarray(I1, np.uint8)
        I2 = np.ascontiguousarray(I2, np.uint8)
        D1 = np.zeros((self.H, self.W), np.float32) if D1 is None else D1
        D2 = np.zeros((self.H, self.W), np.float32) if D2 is None else D2
        rc = self.lib.svb_band_process(self.h, _ptr(I1), _ptr(I2), self.W, _ptr(D1), _ptr(D2))
        if rc != 0:
            raise SvbError(rc, self.lib.svb_last_error().decode())
        return D1, D2

    def stats(self):
        s = BandStats()
        self.lib.svb_band_get_stats(self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in BandStats._fields_}

    def close(self):
        if self.h:
            self.lib.svb_band_destroy(self.h)
            self.h = None


class Context:
    """One svb_context: the B200 replacement of `ElasGPU elas(param)` for frames of a fixed size."""

    def __init__(self, params, width, height, chunk=1, device=-1):
        self.lib = load()
        self.p = params
        self.W, self.H = width, height
        self.h = self.lib.svb_create(C.byref(params), width, height, chunk, device)
        if not self.h:
            raise SvbError(-1, self.lib.svb_last_error().decode())
        step = params.candidate_stepsize
        if params.subsampling:
            step += step % 2  # elas.cpp:376-378
        # disparity maps are (W/2) x (H/2) with subsampling (elas.h:81-83,157-160)
        self.Dw, self.Dh = (width // 2, height // 2) if params.subsampling else (width, height)
        self.cw = (width + step - 1) // step
        self.ch = (height + step - 1) // step
        self.gw = -(-width // params.grid_size)
        self.gh = -(-height // params.grid_size)
        self.maxS = (self.cw - 1) * (self.ch - 1) + 6

    def close(self):
        if self.h:
            self.lib.svb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise SvbError(rc, self.lib.svb_last_error().decode())

    # ---- configuration -----------------------------------------------------------------------
    def set_tap_mode(self, on=True):
        self._chk(self.lib.svb_set_tap_mode(self.h, int(on)))

    def set_mean_mode(self, mode):
        self._chk(self.lib.svb_set_mean_mode(self.h, int(mode)))

    def set_stage_timing(self, on=True):
        self._chk(self.lib.svb_set_stage_timing(self.h, int(on)))

    def set_single_stream(self, on=True):
        self._chk(self.lib.svb_set_single_stream(self.h, int(on)))

    def delaunay_pipeline(self, support, right):
        """The Delaunay stage as the pipeline runs it.  Returns (triangles, used): 2 = the device made the whole list
        (k_delaunay.cu), 1 = device vertex order + host recursion, 0 = complete host path."""
        support = np.ascontiguousarray(support, np.int32)
        n = len(support)
        cap = 2 * n + 16
        tri = np.zeros((cap, 3), np.int32)
        m, used = C.c_int(0), C.c_int(0)
        self._chk(self.lib.svb_stage_delaunay_pipeline(self.h, _ptr(support), n, int(right), _ptr(tri), cap, C.byref(m), C.byref(used)))
        return tri[: m.value].copy(), int(used.value)

    def set_eval_counting(self, on=True):
        self._chk(self.lib.svb_set_eval_counting(self.h, int(on)))

    def eval_counts(self):
        """(support hypotheses, dense hypotheses) counted since the last call; see svb_set_eval_counting."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._chk(self.lib.svb_get_eval_counts(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_delaunay_threads(self, n):
        self._chk(self.lib.svb_set_delaunay_threads(self.h, int(n)))

    def inject_triangles(self, side, tri):
        if tri is None:
            self._chk(self.lib.svb_inject_triangles(self.h, side, None, -1))
        else:
            tri = np.ascontiguousarray(tri, np.int32)
            self._chk(self.lib.svb_inject_triangles(self.h, side, _ptr(tri), len(tri)))

    def set_calibration(self, Q, XR=None, XT=None):
        Q = np.ascontiguousarray(Q, np.float64)
        XR = None if XR is None else np.ascontiguousarray(XR, np.float64)
        XT = None if XT is None else np.ascontiguousarray(XT, np.float64)
        self._chk(self.lib.svb_set_calibration(self.h, _ptr(Q), _ptr(XR), _ptr(XT)))

    # ---- Elas::process -----------------------------------------------------------------------------
    def process(self, I1, I2):
        I1 = np.ascontiguousarray(I1, np.uint8)
        I2 = np.ascontiguousarray(I2, np.uint8)
        assert I1.shape == (self.H, self.W) and I2.shape == (self.H, self.W)
        D1 = np.zeros((self.Dh, self.Dw), np.float32)
        D2 = np.zeros((self.Dh, self.Dw), np.float32)
        self._chk(self.lib.svb_process(self.h, _ptr(I1), _ptr(I2), self.W, _ptr(D1), _ptr(D2)))
        return D1, D2

    _TAP_SPECS = {
        "desc1": (np.uint8, lambda s: (s.H, s.W, 16)),
        "desc2": (np.uint8, lambda s: (s.H, s.W, 16)),
        "dcan_raw": (np.int16, lambda s: (s.ch, s.cw)),
        "dcan": (np.int16, lambda s: (s.ch, s.cw)),
        "support": (np.int32, lambda s: (-1, 3)),
        "tri1": (np.int32, lambda s: (-1, 3)),
        "tri2": (np.int32, lambda s: (-1, 3)),
        "planes1": (np.float32, lambda s: (-1, 6)),
        "planes2": (np.float32, lambda s: (-1, 6)),
        "grid1": (np.int32, lambda s: (s.gh, s.gw, s.p.disp_max + 2)),
        "grid2": (np.int32, lambda s: (s.gh, s.gw, s.p.disp_max + 2)),
        "owner1": (np.int32, lambda s: (s.Dh, s.Dw)),
        "owner2": (np.int32, lambda s: (s.Dh, s.Dw)),
    }

    def tap(self, name):
        if name in self._TAP_SPECS:
            dt, shp = self._TAP_SPECS[name]
            shape = shp(self)
        else:
            dt, shape = np.float32, (self.Dh, self.Dw)
        cap = max(self.W * self.H * 16, self.gw * self.gh * (self.p.disp_max + 2) * 4, 1 << 20)
        buf = np.zeros(cap, np.uint8)
        n = self.lib.svb_tap(self.h, name.encode(), _ptr(buf), cap)
        if n < 0:
            raise SvbError(int(n), self.lib.svb_last_error().decode())
        return buf[:n].view(dt).reshape(shape).copy()

    # ---- stage-isolated entry points ------------------------------------------------------------------
    def descriptor(self, I):
        I = np.ascontiguousarray(I, np.uint8)
        out = np.zeros((self.H, self.W, 16), np.uint8)
        self._chk(self.lib.svb_stage_descriptor(self.h, _ptr(I), self.W, _ptr(out)))
        return out

    def support(self, desc1, desc2):
        desc1 = np.ascontiguousarray(desc1, np.uint8)
        desc2 = np.ascontiguousarray(desc2, np.uint8)
        raw = np.zeros((self.ch, self.cw), np.int16)
        fin = np.zeros((self.ch, self.cw), np.int16)
        pts = np.zeros((self.maxS, 3), np.int32)
        n = C.c_int(0)
        self._chk(self.lib.svb_stage_support(self.h, _ptr(desc1), _ptr(desc2), _ptr(raw), _ptr(fin), _ptr(pts), self.maxS, C.byref(n)))
        return raw, fin, pts[: n.value].copy()

    def planes(self, support, tri):
        support = np.ascontiguousarray(support, np.int32)
        tri = np.ascontiguousarray(tri, np.int32)
        out = np.zeros((len(tri), 6), np.float32)
        self._chk(self.lib.svb_stage_planes(self.h, _ptr(support), len(support), _ptr(tri), len(tri), _ptr(out)))
        return out

    def grid(self, support, right):
        support = np.ascontiguousarray(support, np.int32)
        g = np.zeros((self.gh, self.gw, self.p.disp_max + 2), np.int32)
        self._chk(self.lib.svb_stage_grid(self.h, _ptr(support), len(support), int(right), _ptr(g)))
        return g

    def disparity(self, support, tri, desc1, desc2, right):
        support = np.ascontiguousarray(support, np.int32)
        tri = np.ascontiguousarray(tri, np.int32)
        desc1 = np.ascontiguousarray(desc1, np.uint8)
        desc2 = np.ascontiguousarray(desc2, np.uint8)
        D = np.zeros((self.Dh, self.Dw), np.float32)
        self._chk(self.lib.svb_stage_disparity(self.h, _ptr(support), len(support), _ptr(tri), len(tri), _ptr(desc1), _ptr(desc2), int(right),
                                               _ptr(D)))
        return D

    def lr_check(self, D1, D2):
        a = np.ascontiguousarray(D1, np.float32).copy()
        b = np.ascontiguousarray(D2, np.float32).copy()
        self._chk(self.lib.svb_stage_lr_check(self.h, _ptr(a), _ptr(b)))
        return a, b

    def _inplace(self, fn, D):
        a = np.ascontiguousarray(D, np.float32).copy()
        self._chk(fn(self.h, _ptr(a)))
        return a

    def remove_small_segments(self, D):
        return self._inplace(self.lib.svb_stage_remove_small_segments, D)

    def gap_interpolation(self, D):
        return self._inplace(self.lib.svb_stage_gap_interpolation, D)

    def adaptive_mean(self, D):
        return self._inplace(self.lib.svb_stage_adaptive_mean, D)

    def median(self, D):
        return self._inplace(self.lib.svb_stage_median, D)

    def reproject(self, D, Q, XR=None, XT=None):
        D = np.ascontiguousarray(D, np.float32)
        Q = np.ascontiguousarray(Q, np.float64)
        XR = None if XR is None else np.ascontiguousarray(XR, np.float64)
        XT = None if XT is None else np.ascontiguousarray(XT, np.float64)
        dmap = np.zeros((self.H, self.W), np.uint8)
        pts = np.zeros((self.H * self.W, 3), np.float64)
        self._chk(self.lib.svb_stage_reproject(self.h, _ptr(D), _ptr(Q), _ptr(XR), _ptr(XT), _ptr(dmap), _ptr(pts)))
        return dmap, pts

    def bgra_to_gray(self, bgra):
        """cv::cvtColor(BGRA2GRAY) (stereo_vision.cu:346-347) on the device."""
        bgra = np.ascontiguousarray(bgra, np.uint8)
        assert bgra.shape == (self.H, self.W, 4)
        out = np.zeros((self.H, self.W), np.uint8)
        self._chk(self.lib.svb_stage_bgra_to_gray(self.h, _ptr(bgra), _ptr(out)))
        return out

    def point_cloud_bgra(self, left_bgra, right_bgra):
        """Body of generatePointCloud(): returns (points[N,3] f64, dmap u8, D1 f32, (dmap_ms, pc_ms))."""
        left_bgra = np.ascontiguousarray(left_bgra, np.uint8)
        right_bgra = np.ascontiguousarray(right_bgra, np.uint8)
        pts = np.zeros((self.H * self.W, 3), np.float64)
        dmap = np.zeros((self.H, self.W), np.uint8)
        D1 = np.zeros((self.H, self.W), np.float32)
        t = np.zeros(2, np.float64)
        rc = self.lib.svb_point_cloud_bgra(self.h, _ptr(left_bgra), _ptr(right_bgra), _ptr(pts), _ptr(dmap), _ptr(D1), _ptr(t))
        if rc not in (0, ERR_FEW_SUPPORT):
            self._chk(rc)
        return pts, dmap, D1, (float(t[0]), float(t[1]))

    # ---- batch pipeline ------------------------------------------------------------------------------------
    def batch_upload(self, left, right):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        n = left.shape[0]
        self._chk(self.lib.svb_batch_upload(self.h, _ptr(left), _ptr(right), n))
        return n

    def batch_upload_bgra(self, left_bgra, right_bgra):
        left_bgra = np.ascontiguousarray(left_bgra, np.uint8)
        right_bgra = np.ascontiguousarray(right_bgra, np.uint8)
        n = left_bgra.shape[0]
        self._chk(self.lib.svb_batch_upload_bgra(self.h, _ptr(left_bgra), _ptr(right_bgra), n))
        return n

    def batch_device_ptrs(self):
        """(D1 device pointer or 0, points device pointer or 0, frames, device) of the last batch call's resident results."""
        d1, pts, n, dev = C.c_void_p(0), C.c_void_p(0), C.c_int(0), C.c_int(0)
        self._chk(self.lib.svb_batch_device_ptrs(self.h, C.byref(d1), C.byref(pts), C.byref(n), C.byref(dev)))
        return d1.value or 0, pts.value or 0, n.value, dev.value

    def batch_run(self, n, flags=OUT_POINTS):
        self._chk(self.lib.svb_batch_run(self.h, n, flags))

    def batch_run_host(self, left, right, flags=OUT_POINTS, D1_out=None, points_out=None):
        n = left.shape[0]
        self._chk(self.lib.svb_batch_run_host(self.h, _ptr(left), _ptr(right), n, flags, _ptr(D1_out), _ptr(points_out)))

    def batch_frame_support(self, n):
        """Support points of every frame of the last batch call (< 3: the frame failed and its disparity is 0 everywhere)."""
        out = np.zeros(n, np.int32)
        self._chk(self.lib.svb_batch_frame_support(self.h, _ptr(out), n))
        return out

    def batch_disparity(self, frame):
        out = np.zeros((self.Dh, self.Dw), np.float32)
        self._chk(self.lib.svb_batch_download_disparity(self.h, frame, _ptr(out)))
        return out

    def batch_points(self, frame):
        out = np.zeros((self.H * self.W, 3), np.float64)
        self._chk(self.lib.svb_batch_download_points(self.h, frame, _ptr(out)))
        return out

    def stats(self):
        s = Stats()
        self._chk(self.lib.svb_get_stats(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in Stats._fields_ if k != "stage_ms"}
        d["stage_ms"] = {STAGE_NAMES[i]: s.stage_ms[i] for i in range(len(STAGE_NAMES))}
        return d
