"""ctypes wrapper around oracle/_ref/libvsg_ref.so — the REFERENCE's own ORBextractor.cc (and ORBmatcher.cc),
compiled unmodified by oracle/ref_build/Makefile against the cv2-pinned OpenCV compat layer.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
/root/reference exists only in the build container; the GPU box uses the prebuilt library that travels with
the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as _orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libvsg_ref.so")
_REF_ROOT = os.environ.get("VSG_REFERENCE_ROOT", "/root/reference/orb_slam3")
KEYPOINT_DTYPE = _orc.KEYPOINT_DTYPE


def available():
    return os.path.exists(_LIB_PATH) or os.path.isdir(_REF_ROOT)


def build(force=False):
    """(Re)build from the reference sources when they are present; otherwise use the prebuilt library."""
    _orc.build()
    if os.path.isdir(_REF_ROOT):
        # the library and the CPU matcher test; ref_matcher_test_gpu (needs libvsg_cuda.so) is built by `make all`
        cmd = ["make", "-C", os.path.join(_HERE, "ref_build"), "REF=" + _REF_ROOT, "../_ref/libvsg_ref.so",
               "../_ref/ref_matcher_test_cpu"]
        if force:
            cmd.append("-B")
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError("oracle/_ref/libvsg_ref.so is missing and %s is not present to build it" % _REF_ROOT)
    return _LIB_PATH


_lib = None


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        build()
        _orc.lib()  # liborb_oracle.so first: libvsg_ref.so resolves the cv2-pinned primitives from it
        L = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        L.ref_extractor_create.restype = vp
        L.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_extractor_destroy.argtypes = [vp]
        L.ref_levels.argtypes = [vp]
        L.ref_tables.argtypes = [vp] * 7
        L.ref_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_num_keypoints.argtypes = [vp]
        L.ref_get_keypoints.argtypes = [vp, vp, vp]
        L.ref_level_size.argtypes = [vp, C.c_int, vp, vp]
        L.ref_get_level.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ref_level_keypoints.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.ref_distribute_octree.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.ref_bench_extract.restype = C.c_double
        L.ref_bench_extract.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_int, vp]
        _lib = L
    return _lib


class RefExtractor:
    """VS_GRAPHS::ORBextractor itself (orb_slam3/include/ORBextractor.h:42-119)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7):
        self._L = lib()
        self._h = self._L.ref_extractor_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast)
        self.nlevels = nlevels

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_extractor_destroy(self._h)
            self._h = None

    def tables(self):
        n = self.nlevels
        s, i, s2, i2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        u = np.zeros(16, np.int32)
        self._L.ref_tables(self._h, _ptr(s), _ptr(i), _ptr(s2), _ptr(i2), _ptr(q), _ptr(u))
        return dict(scale=s, inv_scale=i, sigma2=s2, inv_sigma2=i2, quota=q, umax=u)

    def __call__(self, image, lapping=(0, 0)):
        """operator(): returns (mono_index, keypoints structured array, descriptors n x 32)."""
        if image is None or image.size == 0:
            z = np.zeros(1, np.uint8)
            mono = self._L.ref_extract(self._h, _ptr(z), 0, 0, 0, int(lapping[0]), int(lapping[1]))
            return mono, np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        h, w = image.shape
        mono = self._L.ref_extract(self._h, _ptr(image), w, h, image.strides[0], int(lapping[0]), int(lapping[1]))
        n = self._L.ref_num_keypoints(self._h)
        kps = np.zeros(n, KEYPOINT_DTYPE)
        desc = np.zeros((n, 32), np.uint8)
        if n:
            self._L.ref_get_keypoints(self._h, _ptr(kps), _ptr(desc))
        return mono, kps, desc

    def level_size(self, level):
        w, h = C.c_int32(), C.c_int32()
        self._L.ref_level_size(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self._L.ref_get_level(self._h, level, 0, _ptr(out))
        return out

    def level_padded(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self._L.ref_get_level(self._h, level, 19, _ptr(out))
        return out

    def level_keypoints(self, image, level):
        """ComputePyramid + ComputeKeyPointsOctTree alone: one level's keypoints in level coordinates."""
        h, w = image.shape
        out = np.zeros(1 << 16, KEYPOINT_DTYPE)
        n = self._L.ref_level_keypoints(self._h, _ptr(image), w, h, image.strides[0], level, _ptr(out), len(out))
        return out[:n].copy()

    def distribute_octree(self, xyr, min_x, max_x, min_y, max_y, quota):
        xyr = np.ascontiguousarray(xyr, np.float32)
        out = np.zeros((len(xyr) + 8, 3), np.float32)
        n = self._L.ref_distribute_octree(self._h, _ptr(xyr), len(xyr), min_x, max_x, min_y, max_y, quota, _ptr(out),
                                          len(out))
        return out[:n].copy()


def matcher_test_binary(kind):
    """Path of oracle/_ref/ref_matcher_test_{cpu,gpu} (tests/cpp/ref_matcher_test.cpp), or None if it was not built."""
    p = os.path.join(_HERE, "_ref", "ref_matcher_test_" + kind)
    return p if os.path.exists(p) else None


def bench_extract(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, threads=1):
    """Seconds the reference extractor needs for `frames` (n x h x w uint8) on `threads` host threads."""
    frames = np.ascontiguousarray(frames)
    n, h, w = frames.shape
    total = C.c_int64()
    secs = lib().ref_bench_extract(_ptr(frames), n, w, h, nfeatures, scale_factor, nlevels, ini_th, min_th, threads,
                                   C.byref(total))
    return secs, total.value
