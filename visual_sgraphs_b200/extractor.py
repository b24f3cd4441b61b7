"""Python mirror of VS_GRAPHS::ORBextractor (reference orb_slam3/include/ORBextractor.h:42-119) over the
C ABI.  Same constructor arguments, same call semantics (returns monoIndex, -1 for an empty image),
same getters; the work happens in libvsg_cuda.so on the GPU.  Used by tests and bench.py; C++ hosts
use visual_sgraphs_b200/shim/ORBextractor.h instead.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KEYPOINT_DTYPE, OrbParams, check, ptr


class ORBextractor:
    HARRIS_SCORE = 0  # ORBextractor.h:45-49 (unused by the reference, kept for source compatibility)
    FAST_SCORE = 1

    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, device=0, max_batch=1):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.nlevels = nlevels
        self.max_batch = max_batch
        p = OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
        check(self._L.vsg_extractor_create(C.byref(p), device, max_batch, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsg_extractor_destroy(self._h)
            self._h = None

    __del__ = close

    # -- getters (ORBextractor.h:63-91) --
    def _tables(self):
        n = self.nlevels
        s, i, s2, i2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        check(self._L.vsg_extractor_tables(self._h, ptr(s), ptr(i), ptr(s2), ptr(i2), ptr(q)))
        return s, i, s2, i2, q

    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        return self._tables()[4]

    def max_keypoints(self, width, height):
        n = self._L.vsg_extractor_max_keypoints(self._h, width, height)
        if n < 0:
            check(n)
        return n

    # -- operator() (ORBextractor.cc:1083-1169) --
    def __call__(self, image, lapping=(0, 0)):
        """Returns (monoIndex, keypoints[KEYPOINT_DTYPE], descriptors n x 32 uint8); monoIndex -1 for an empty image."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1  # CV_8UC1 (:1091)
        h, w = image.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        check(self._L.vsg_extract(self._h, ptr(image), w, h, image.strides[0], int(lapping[0]), int(lapping[1]),
                                  ptr(kps), ptr(desc), cap, C.byref(n), C.byref(mono)))
        return mono.value, kps[: n.value].copy(), desc[: n.value].copy()

    def extract_batch(self, frames, lapping=(0, 0)):
        """frames: (n, h, w) uint8 host array (pinned or pageable). Returns list of (mono, kps, desc)."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.strides[2] == 1
        nf, h, w = frames.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros((nf, cap), KEYPOINT_DTYPE)
        desc = np.zeros((nf, cap, 32), np.uint8)
        n = np.zeros(nf, np.int32)
        mono = np.zeros(nf, np.int32)
        check(self._L.vsg_extract_batch(self._h, ptr(frames), nf, w, h, frames.strides[1], frames.strides[0],
                                        int(lapping[0]), int(lapping[1]), ptr(kps), ptr(desc), cap, ptr(n), ptr(mono)))
        return [(int(mono[f]), kps[f, : n[f]].copy(), desc[f, : n[f]].copy()) for f in range(nf)]

    def extract_batch_color(self, frames, rgb=True, lapping=(0, 0)):
        """frames: (n, h, w, 3|4) uint8 interleaved colour frames; rgb: channel order as Tracking's mbRGB flag.
        cvtColor(..., COLOR_{RGB,BGR}[A]2GRAY) runs on the device. Returns list of (mono, kps, desc)."""
        assert frames.dtype == np.uint8 and frames.ndim == 4 and frames.shape[3] in (3, 4) and frames.strides[3] == 1
        nf, h, w, ch = frames.shape
        assert frames.strides[2] == ch
        cap = self.max_keypoints(w, h)
        kps = np.zeros((nf, cap), KEYPOINT_DTYPE)
        desc = np.zeros((nf, cap, 32), np.uint8)
        n = np.zeros(nf, np.int32)
        mono = np.zeros(nf, np.int32)
        check(self._L.vsg_extract_batch_color(self._h, ptr(frames), nf, w, h, frames.strides[1], frames.strides[0], ch,
                                              int(bool(rgb)), int(lapping[0]), int(lapping[1]), ptr(kps), ptr(desc), cap,
                                              ptr(n), ptr(mono)))
        return [(int(mono[f]), kps[f, : n[f]].copy(), desc[f, : n[f]].copy()) for f in range(nf)]

    def set_rectify_map(self, slot, map_x, map_y):
        """Rectification map of camera `slot` (0 = left / only, 1 = right): the CV_32FC1 pair of
        cv::initUndistortRectifyMap (Settings.cc:571-574); the maps' shape is the rectified image size."""
        map_x = np.ascontiguousarray(map_x, np.float32)
        map_y = np.ascontiguousarray(map_y, np.float32)
        assert map_x.ndim == 2 and map_x.shape == map_y.shape
        check(self._L.vsg_extractor_set_rectify_map(self._h, int(slot), ptr(map_x), ptr(map_y), map_x.shape[1], map_x.shape[0]))
        self._rect_shape = map_x.shape

    def extract_batch_rectify(self, frames, ncameras=1, lapping=(0, 0)):
        """frames: (n, h, w) uint8 unrectified gray frames; frame f is remapped (cv::remap INTER_LINEAR, System.cc:284-292)
        with the map of camera f % ncameras on the device and extracted. Returns list of (mono, kps, desc)."""
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.strides[2] == 1
        nf, h, w = frames.shape
        rh, rw = self._rect_shape
        cap = self.max_keypoints(rw, rh)
        kps = np.zeros((nf, cap), KEYPOINT_DTYPE)
        desc = np.zeros((nf, cap, 32), np.uint8)
        n = np.zeros(nf, np.int32)
        mono = np.zeros(nf, np.int32)
        check(self._L.vsg_extract_batch_rectify(self._h, ptr(frames), nf, w, h, frames.strides[1], frames.strides[0],
                                                int(ncameras), int(lapping[0]), int(lapping[1]), ptr(kps), ptr(desc), cap,
                                                ptr(n), ptr(mono)))
        return [(int(mono[f]), kps[f, : n[f]].copy(), desc[f, : n[f]].copy()) for f in range(nf)]

    def extract_batch_dev(self, frames_dev, kps_dev, desc_dev, n_dev, mono_dev, lapping=(0, 0)):
        """Device-resident variant: torch uint8 CUDA tensors. frames (n,h,w); kps (n,cap,28) u8; desc (n,cap,32)
        u8; n/mono int32 (n). Asynchronous on the handle's stream."""
        nf, h, w = frames_dev.shape
        cap = kps_dev.shape[1]
        check(self._L.vsg_extract_batch_dev(self._h, ptr(frames_dev), nf, w, h, frames_dev.stride(1),
                                            frames_dev.stride(0), int(lapping[0]), int(lapping[1]), ptr(kps_dev),
                                            ptr(desc_dev), cap, ptr(n_dev), ptr(mono_dev)))

    def sync(self):
        check(self._L.vsg_extractor_sync(self._h))

    def stream(self):
        return self._L.vsg_extractor_stream(self._h)

    STAGES = ("pyramid", "fast", "octree", "blur", "describe")

    def profile(self, enable=True):
        check(self._L.vsg_extractor_profile(self._h, int(enable)))

    def stage_ms(self):
        """(dict stage -> accumulated ms, runs) since the last call; waits for the stream."""
        ms = np.zeros(len(self.STAGES), np.float64)
        runs = C.c_int64(0)
        check(self._L.vsg_extractor_stage_ms(self._h, ptr(ms), C.byref(runs)))
        return dict(zip(self.STAGES, ms.tolist())), runs.value

    # -- mvImagePyramid and the diagnostic taps --
    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        check(self._L.vsg_pyramid_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def pyramid_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        check(self._L.vsg_pyramid_download(self._h, frame, level, ptr(out), w))
        return out

    def blurred_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        check(self._L.vsg_blurred_download(self._h, frame, level, ptr(out), w))
        return out

    def candidates(self, level, frame=0):
        w, h = self.level_size(level)
        cap = max(16, (w * h) // 3)
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        check(self._L.vsg_candidates_download(self._h, frame, level, ptr(out), cap, C.byref(n)))
        return out[: n.value].copy()

    def level_keypoints(self, level, frame=0):
        cap = 65536
        out = np.zeros((cap, 3), np.int32)
        n = C.c_int()
        check(self._L.vsg_level_keypoints_download(self._h, frame, level, ptr(out), cap, C.byref(n)))
        return out[: n.value].copy()
