"""The reference's Python plugin class on top of this library: `stereo_vision` of stereo_vision/sv.py:154-192.

Same class name, constructor arguments (names, order, defaults), `generatePointCloud(left, right)` and `__del__` as the reference's
module, bound to the same three C symbols (`generatePointCloud`, `clean`, `getColor`: include/stereo_vision_c.h) with the same
14-argument ctypes prototype (sv.py:180), so code written against `from stereo_vision.sv import stereo_vision` runs against this
class unchanged.  What differs, all of it forced by the environment rather than chosen:

 * the reference module searches site-packages for `stereo_vision*.so` at import time and raises IndexError when there is none
   (sv.py:139-149); here the default `so_lib_path` is the library built in this tree (build/bin/stereo_vision_parallel.so, else
   the package's lib/libelas_b200.so);
 * `ndarray.tostring()` (sv.py:187-188) no longer exists in numpy 2: `tobytes()` is the same bytes;
 * BGR -> BGRA (cv2.cvtColor(..., COLOR_BGR2BGRA), sv.py:185-186) is done in numpy -- alpha 255, as OpenCV sets it -- so that
   OpenCV is not needed; 4-channel input is passed through, a single-channel image is replicated.

Kept from the reference, including its sharp edges: the returned array ALIASES the library-owned point buffer and is overwritten
by the next call (sv.py:167); the first call latches width / height / calibration (stereo_vision.cu:591); `__del__` calls `clean()`,
which prints "Program exitted successfully!" and ends the process with exit(0) (stereo_vision.cu:114-126) unless
SVB_CLEAN_NO_EXIT=1 is set; `subsampling` is accepted and never passed on (the reference's prototype has 14 of the 16 arguments).

No computation happens here: every byte of the result comes from the CUDA library; without a CUDA device the library logs the
error and the buffer stays zero (there is no CPU path).
"""
import ctypes
import os

import numpy as np
from numpy.ctypeslib import ndpointer

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG_DIR)


def _default_so_path():
    for path in (os.path.join(_ROOT, "build", "bin", "stereo_vision_parallel.so"), os.path.join(_PKG_DIR, "lib", "libelas_b200.so")):
        if os.path.exists(path):
            return path
    raise FileNotFoundError("no built library: run `make shared_library` (build/bin/stereo_vision_parallel.so)")


def _to_bgra(img):
    """cv2.cvtColor(img, cv2.COLOR_BGR2BGRA) for a u8 H x W x 3 image (sv.py:185-186); BGRA passes, gray is replicated."""
    img = np.asarray(img)
    if img.dtype != np.uint8:
        raise TypeError("stereo_vision.generatePointCloud: images must be uint8, got %s" % img.dtype)
    if img.ndim == 2:
        img = np.stack([img, img, img], -1)
    if img.ndim != 3 or img.shape[2] not in (3, 4):
        raise ValueError("stereo_vision.generatePointCloud: expected an H x W x 3 BGR image, got shape %s" % (img.shape,))
    if img.shape[2] == 4:
        return np.ascontiguousarray(img)
    out = np.empty(img.shape[:2] + (4,), np.uint8)
    out[..., :3] = img
    out[..., 3] = 255
    return out


class stereo_vision:  # noqa: N801 -- the reference's name
    def __init__(self, so_lib_path=None, width=1242, height=375, defaultCalibFile=True, objectTracking=True, graphics=False, display=False,
                 scale=1, pc_extrapolation=1, YOLO_CFG='src/yolo/yolov4-tiny.cfg', YOLO_WEIGHTS='src/yolo/yolov4-tiny.weights',
                 YOLO_CLASSES='src/yolo/classes.txt', CAMERA_CALIBRATION_YAML='data/calibration/kitti_2011_09_26.yml', subsampling=False):
        self.sv = ctypes.CDLL(so_lib_path or _default_so_path())
        self.width = width
        self.height = height
        self.sv.generatePointCloud.restype = ndpointer(dtype=ctypes.c_double, shape=(width * height, 3))  # sv.py:167

        self.defaultCalibFile = defaultCalibFile
        self.objectTracking = objectTracking
        self.graphics = graphics
        self.display = display
        self.scale = scale
        self.pc_extrapolation = pc_extrapolation

        self.YOLO_CFG = YOLO_CFG
        self.YOLO_WEIGHTS = YOLO_WEIGHTS
        self.YOLO_CLASSES = YOLO_CLASSES
        self.CAMERA_CALIBRATION_YAML = CAMERA_CALIBRATION_YAML
        self.sv.generatePointCloud.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_bool,
                                               ctypes.c_bool, ctypes.c_bool, ctypes.c_bool, ctypes.c_int, ctypes.c_int, ctypes.c_char_p,
                                               ctypes.c_char_p, ctypes.c_char_p]  # sv.py:180: 14 of the 16 parameters
        self.sv.clean.restype = None
        self.sv.clean.argtypes = []

    def generatePointCloud(self, left, right):  # noqa: N802 -- the reference's name
        left = _to_bgra(left)
        right = _to_bgra(right)
        for img in (left, right):
            if img.shape[0] != self.height or img.shape[1] != self.width:
                # the reference wraps the bytes in a cv::Mat of the latched size without looking (stereo_vision.cu:596-597) and reads
                # out of bounds; refuse instead
                raise ValueError("stereo_vision.generatePointCloud: image is %dx%d, the object was made for %dx%d" %
                                 (img.shape[1], img.shape[0], self.width, self.height))
        return self.sv.generatePointCloud(left.tobytes(), right.tobytes(), self.CAMERA_CALIBRATION_YAML.encode('utf-8'), self.width, self.height,
                                          self.defaultCalibFile, self.objectTracking, self.graphics, self.display, self.scale, self.pc_extrapolation,
                                          self.YOLO_CFG.encode('utf-8'), self.YOLO_WEIGHTS.encode('utf-8'), self.YOLO_CLASSES.encode('utf-8'))

    def getColor(self):  # noqa: N802
        """The BGRA image behind the last cloud (`getColor`, stereo_vision.cu:634-636), aliased like the points."""
        self.sv.getColor.restype = ctypes.POINTER(ctypes.c_ubyte)
        ptr = self.sv.getColor()
        if not ptr:
            return None
        return np.ctypeslib.as_array(ptr, shape=(self.height, self.width, 4))

    def __del__(self):
        sv = getattr(self, "sv", None)
        if sv is not None:
            sv.clean()  # sv.py:191-192
