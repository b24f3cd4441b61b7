"""Flattened Frame / KeyFrame view (what the reference's Search* methods read, Frame.h:254-290,363-381) and
its device-side handle.  Mirrors Frame's grid constants: FRAME_GRID_COLS 64, FRAME_GRID_ROWS 48 (Frame.h:49-50),
mfGridElementWidthInv = COLS / (mnMaxX - mnMinX) (Frame.cc:171-172)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KEYPOINT_DTYPE, FrameView, check

FRAME_GRID_ROWS = 48
FRAME_GRID_COLS = 64


class FrameData:
    """Host arrays of a frame + the C view struct pointing at them (kept alive together)."""

    def __init__(self, keys, descriptors, u_right=None, bounds=None, scale_factors=None, width=640, height=480,
                 grid_cols=FRAME_GRID_COLS, grid_rows=FRAME_GRID_ROWS):
        self.keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
        self.descriptors = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        assert len(self.keys) == len(self.descriptors)
        self.u_right = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        if bounds is None:
            bounds = (0.0, 0.0, float(width), float(height))   # mnMinX, mnMinY, mnMaxX, mnMaxY without distortion
        self.bounds = tuple(np.float32(b) for b in bounds)
        if scale_factors is None:
            s = [np.float32(1.0)]
            for _ in range(7):
                s.append(np.float32(float(s[-1]) * float(np.float32(1.2))))
            scale_factors = np.array(s, np.float32)
        self.scale_factors = np.ascontiguousarray(scale_factors, np.float32)
        v = FrameView()
        v.n = len(self.keys)
        v.keys = self.keys.ctypes.data
        v.descriptors = self.descriptors.ctypes.data
        v.u_right = None if self.u_right is None else self.u_right.ctypes.data
        v.min_x, v.min_y, v.max_x, v.max_y = self.bounds
        v.grid_inv_w = np.float32(grid_cols) / (self.bounds[2] - self.bounds[0])
        v.grid_inv_h = np.float32(grid_rows) / (self.bounds[3] - self.bounds[1])
        v.grid_cols, v.grid_rows = grid_cols, grid_rows
        v.scale_factors = self.scale_factors.ctypes.data
        v.n_levels = len(self.scale_factors)
        self.view = v

    @property
    def n(self):
        return len(self.keys)


class DeviceFrame:
    """vsg_frame handle: the view uploaded once, with its keypoint grid built (Frame::AssignFeaturesToGrid)."""

    def __init__(self, matcher, data):
        self._L = _lib.load()
        self.data = data
        self._matcher = matcher   # keep the uploading matcher (its stream) alive as long as the frame
        self._h = C.c_void_p()
        check(self._L.vsg_frame_create(matcher._h, C.byref(data.view), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsg_frame_destroy(self._h)
            self._h = None

    __del__ = close
