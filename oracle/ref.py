"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libelas_ref.so.

That library is the reference's own serial ELAS (src/serial_includes/elas/elas.cpp and
src/common_includes/elas/*.cpp) compiled by oracle/build_ref.sh with strict-IEEE flags, wrapped by
oracle/ref_taps.cpp so that every stage can be called on its own.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Params(C.Structure):
    """POD mirror of Elas::parameters (src/parallel_includes/elas/elas.h:58-83); shared layout with
    svb_params in include/elas_b200.h."""

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

    def copy(self):
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


ROBOTICS, MIDDLEBURY = 0, 1


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefElas:
    def __init__(self, fast=False, omp=False):
        # omp: the reference's OpenMP variant (timing only; it has no stage taps, only process())
        name = "libelas_ref_omp.so" if omp else ("libelas_ref_fast.so" if fast else "libelas_ref.so")
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " missing: run oracle/build_ref.sh where /root/reference exists")
        self.lib = C.CDLL(path)
        self.lib.ref_process.restype = C.c_double
        self.lib.ref_build_flags.restype = C.c_char_p
        self.flags = self.lib.ref_build_flags().decode()

    # ---- parameters -------------------------------------------------------------------------
    def params(self, setting=ROBOTICS, **over):
        p = Params()
        self.lib.ref_default_params(C.c_int(setting), C.byref(p))
        for k, v in over.items():
            setattr(p, k, v)
        return p

    def pipeline_params(self, subsampling=0):
        """The preset generateDisparityMap() uses (src/parallel_includes/main/stereo_vision.cu:315-319):
        MIDDLEBURY + postprocess_only_left + filter_adaptive_mean."""
        return self.params(MIDDLEBURY, postprocess_only_left=1, filter_adaptive_mean=1, subsampling=subsampling)

    # ---- whole pipeline ---------------------------------------------------------------------
    def process(self, p, I1, I2):
        I1 = np.ascontiguousarray(I1, np.uint8)
        I2 = np.ascontiguousarray(I2, np.uint8)
        H, W = I1.shape
        D1 = np.zeros((H, W), np.float32)
        D2 = np.zeros((H, W), np.float32)
        sec = self.lib.ref_process(C.byref(p), _p(I1, C.c_uint8), _p(I2, C.c_uint8), W, H, W, _p(D1, C.c_float), _p(D2, C.c_float))
        return D1, D2, sec

    # ---- stages -----------------------------------------------------------------------------
    def descriptor(self, I, subsampling=0):
        I = np.ascontiguousarray(I, np.uint8)
        H, W = I.shape
        out = np.zeros((H, W, 16), np.uint8)
        self.lib.ref_descriptor(_p(I, C.c_uint8), W, H, W, int(subsampling), _p(out, C.c_uint8))
        return out

    def dcan_dims(self, p, W, H):
        cw, ch, s = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_dcan_dims(C.byref(p), W, H, C.byref(cw), C.byref(ch), C.byref(s))
        return cw.value, ch.value, s.value

    def support_raw(self, p, desc1, desc2):
        H, W = desc1.shape[:2]
        cw, ch, _ = self.dcan_dims(p, W, H)
        dcan = np.zeros((ch, cw), np.int16)
        self.lib.ref_support_raw(C.byref(p), _p(desc1, C.c_uint8), _p(desc2, C.c_uint8), W, H, _p(dcan, C.c_int16))
        return dcan

    def support_filter(self, p, dcan, which):
        d = np.ascontiguousarray(dcan, np.int16).copy()
        ch, cw = d.shape
        self.lib.ref_support_filter(C.byref(p), _p(d, C.c_int16), cw, ch, int(which))
        return d

    def support(self, p, desc1, desc2):
        H, W = desc1.shape[:2]
        cw, ch, _ = self.dcan_dims(p, W, H)
        cap = cw * ch + 8
        pts = np.zeros((cap, 3), np.int32)
        n = self.lib.ref_support(C.byref(p), _p(desc1, C.c_uint8), _p(desc2, C.c_uint8), W, H, _p(pts, C.c_int32), cap)
        return pts[:n].copy()

    def delaunay(self, pts, right):
        pts = np.ascontiguousarray(pts, np.int32)
        n = len(pts)
        cap = 2 * n + 16
        tri = np.zeros((cap, 3), np.int32)
        m = self.lib.ref_delaunay(_p(pts, C.c_int32), n, int(right), _p(tri, C.c_int32), cap)
        return tri[:m].copy()

    def triangulate_xy(self, xy):
        xy = np.ascontiguousarray(xy, np.float32)
        n = len(xy)
        cap = 2 * n + 16
        tri = np.zeros((cap, 3), np.int32)
        m = self.lib.ref_triangulate_xy(_p(xy, C.c_float), n, _p(tri, C.c_int32), cap)
        return tri[:m].copy()

    def planes(self, pts, tri):
        pts = np.ascontiguousarray(pts, np.int32)
        tri = np.ascontiguousarray(tri, np.int32)
        out = np.zeros((len(tri), 6), np.float32)
        self.lib.ref_planes(_p(pts, C.c_int32), len(pts), _p(tri, C.c_int32), len(tri), _p(out, C.c_float))
        return out

    def grid_dims(self, p, W, H):
        gw, gh = C.c_int(), C.c_int()
        self.lib.ref_grid_dims(C.byref(p), W, H, C.byref(gw), C.byref(gh))
        return gw.value, gh.value

    def grid(self, p, pts, W, H, right):
        pts = np.ascontiguousarray(pts, np.int32)
        gw, gh = self.grid_dims(p, W, H)
        g = np.zeros((gh, gw, p.disp_max + 2), np.int32)
        self.lib.ref_grid(C.byref(p), _p(pts, C.c_int32), len(pts), W, H, int(right), _p(g, C.c_int32))
        return g

    def disparity(self, p, pts, tri, planes, grid, desc1, desc2, right):
        H, W = desc1.shape[:2]
        pts = np.ascontiguousarray(pts, np.int32)
        tri = np.ascontiguousarray(tri, np.int32)
        planes = np.ascontiguousarray(planes, np.float32)
        grid = np.ascontiguousarray(grid, np.int32)
        D = np.zeros((H, W), np.float32)
        self.lib.ref_disparity(C.byref(p), _p(pts, C.c_int32), len(pts), _p(tri, C.c_int32), _p(planes, C.c_float), len(tri),
                               _p(grid, C.c_int32), _p(desc1, C.c_uint8), _p(desc2, C.c_uint8), W, H, int(right), _p(D, C.c_float))
        return D

    def lr_check(self, p, D1, D2):
        a = np.ascontiguousarray(D1, np.float32).copy()
        b = np.ascontiguousarray(D2, np.float32).copy()
        H, W = a.shape
        self.lib.ref_lr_check(C.byref(p), W, H, _p(a, C.c_float), _p(b, C.c_float))
        return a, b

    def _inplace(self, fn, p, D):
        a = np.ascontiguousarray(D, np.float32).copy()
        H, W = a.shape
        fn(C.byref(p), W, H, _p(a, C.c_float))
        return a

    def remove_small_segments(self, p, D):
        return self._inplace(self.lib.ref_remove_small_segments, p, D)

    def gap_interpolation(self, p, D):
        return self._inplace(self.lib.ref_gap_interpolation, p, D)

    def adaptive_mean(self, p, D):
        return self._inplace(self.lib.ref_adaptive_mean, p, D)

    def median(self, p, D):
        return self._inplace(self.lib.ref_median, p, D)

    # ---- every stage in Elas::process order (elas.cpp:31-150), keeping each intermediate ------
    def staged(self, p, I1, I2):
        t = {}
        H, W = I1.shape
        t["desc1"] = self.descriptor(I1, p.subsampling)
        t["desc2"] = self.descriptor(I2, p.subsampling)
        t["dcan_raw"] = self.support_raw(p, t["desc1"], t["desc2"])
        d = self.support_filter(p, t["dcan_raw"], 0)
        t["dcan_incon"] = d
        d = self.support_filter(p, d, 1)
        t["dcan_redv"] = d
        d = self.support_filter(p, d, 2)
        t["dcan"] = d
        t["support"] = self.support(p, t["desc1"], t["desc2"])
        if len(t["support"]) < 3:
            return t
        for side, name in ((0, "1"), (1, "2")):
            t["tri" + name] = self.delaunay(t["support"], side)
            t["planes" + name] = self.planes(t["support"], t["tri" + name])
            t["grid" + name] = self.grid(p, t["support"], W, H, side)
            t["D%sraw" % name] = self.disparity(p, t["support"], t["tri" + name], t["planes" + name], t["grid" + name], t["desc1"], t["desc2"], side)
        t["D1lr"], t["D2lr"] = self.lr_check(p, t["D1raw"], t["D2raw"])
        chain = [("seg", self.remove_small_segments), ("gap", self.gap_interpolation)]
        if p.filter_adaptive_mean:
            chain.append(("mean", self.adaptive_mean))
        if p.filter_median:
            chain.append(("med", self.median))
        sides = ["1"] if p.postprocess_only_left else ["1", "2"]
        for s in ("1", "2"):
            cur = t["D%slr" % s]
            if s in sides:
                for nm, fn in chain:
                    cur = fn(p, cur)
                    t["D%s%s" % (s, nm)] = cur
            t["D" + s] = cur
        return t


class RefProject:
    """TEST INFRASTRUCTURE ONLY: the reference's own `projectParallel` CUDA kernel (stereo_vision.cu:188-212), cut out of the
    reference's driver and compiled by oracle/build_ref.sh into oracle/_ref/libproject_ref.so.  Needs a CUDA device (GPU box)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libproject_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " missing: run oracle/build_ref.sh where /root/reference and nvcc exist")
        self.lib = C.CDLL(path)
        self.lib.ref_project_flags.restype = C.c_char_p
        self.flags = self.lib.ref_project_flags().decode()

    def project(self, dmap, Q, XR, XT):
        """dmap: u8 (H, W) -> (H*W, 3) float64, exactly what publishPointCloud copies back into `points`."""
        dmap = np.ascontiguousarray(dmap, np.uint8)
        H, W = dmap.shape
        Q = np.ascontiguousarray(Q, np.float64).reshape(16)
        XR = np.ascontiguousarray(XR, np.float64).reshape(9)
        XT = np.ascontiguousarray(XT, np.float64).reshape(3)
        pts = np.zeros((H * W, 3), np.float64)
        rc = self.lib.ref_project_parallel(_p(dmap, C.c_uint8), _p(pts, C.c_double), H, W, _p(XT, C.c_double), _p(XR, C.c_double),
                                           _p(Q, C.c_double))
        if rc != 0:
            raise RuntimeError("ref_project_parallel: cudaError %d" % rc)
        return pts


_PORT = None


def port_project(d, H, W, Q, XR, XT):
    """TEST INFRASTRUCTURE ONLY: oracle/project_port.c -- the CPU restatement of `projectParallel` (stereo_vision.cu:188-212) with
    the fused multiply-adds of the reference's own nvcc build (see the header of that file).  d: H*W float64 (the u8 map's values,
    or the float disparity for SVB_OUT_POINTS_FLOATDISP) -> (H*W, 3) float64.  Built by oracle/Makefile (gcc, no reference source
    needed); compiled on the spot when the library is missing."""
    global _PORT
    if _PORT is None:
        path = os.path.join(HERE, "_port", "libproject_port.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.run(["make", "-C", HERE, "-s"], check=True)
        _PORT = C.CDLL(path)
        _PORT.port_project_parallel.restype = None
    d = np.ascontiguousarray(d, np.float64).reshape(-1)
    assert d.size == H * W
    Q = np.ascontiguousarray(Q, np.float64).reshape(16)
    XR = np.ascontiguousarray(XR, np.float64).reshape(9)
    XT = np.ascontiguousarray(XT, np.float64).reshape(3)
    pts = np.empty((H * W, 3), np.float64)
    _PORT.port_project_parallel(_p(d, C.c_double), _p(pts, C.c_double), C.c_int(H), C.c_int(W), _p(XT, C.c_double), _p(XR, C.c_double),
                                _p(Q, C.c_double))
    return pts
