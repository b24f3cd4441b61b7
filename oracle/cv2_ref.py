"""Real-OpenCV (python cv2) restatement of the OpenCV-facing half of the reference extractor.

TEST INFRASTRUCTURE ONLY.  This is the authority for the un-vendored OpenCV primitives the
reference delegates to (cv::resize, cv::copyMakeBorder, cv::GaussianBlur, cv::FAST, cv::fastAtan2 —
call sites ORBextractor.cc:99,832,850,1130,1184,1186,1191): it calls the very same OpenCV functions
through cv2 so the C++ oracle's restated arithmetic can be pinned against them
(tests/test_oracle_cv2.py) and golden vectors can be generated (tests/golden/make_golden.py).
The control flow mirrors ORBextractor.cc:787-900 (per-cell FAST + retry) and :1171-1195 (pyramid).
"""
import ctypes
import math

import cv2
import numpy as np

EDGE = 19
_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.restype = ctypes.c_float
_libm.cosf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]


def scale_tables(nlevels, scale_factor):
    """mvScaleFactor / mvInvScaleFactor (ORBextractor.cc:415-431): float * double member -> float."""
    sf = float(np.float32(scale_factor))  # the double member holds the float argument
    scale = [np.float32(1.0)]
    for _ in range(1, nlevels):
        scale.append(np.float32(float(scale[-1]) * sf))
    inv = [np.float32(1.0) / s for s in scale]
    return scale, inv


def cv_round(v):
    return int(np.rint(v))  # half-to-even


def pyramid(image, nlevels=8, scale_factor=1.2):
    """ComputePyramid (ORBextractor.cc:1171-1195) with cv2. Returns list of padded buffers (h+38, w+38)."""
    _, inv = scale_tables(nlevels, scale_factor)
    h0, w0 = image.shape
    padded = []
    for level in range(nlevels):
        w = cv_round(np.float32(w0) * inv[level])
        h = cv_round(np.float32(h0) * inv[level])
        if level == 0:
            buf = cv2.copyMakeBorder(image, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
        else:
            prev = padded[-1][EDGE:-EDGE, EDGE:-EDGE]
            cur = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR)
            buf = cv2.copyMakeBorder(cur, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101 | cv2.BORDER_ISOLATED)
        assert buf.shape == (h + 2 * EDGE, w + 2 * EDGE)
        padded.append(buf)
    return padded


def cell_grid(w, h):
    """Cell layout of ComputeKeyPointsOctTree (ORBextractor.cc:795-828). Yields (i, j, x0, y0, x1, y1)."""
    min_b = EDGE - 3
    max_bx, max_by = w - EDGE + 3, h - EDGE + 3
    width, height = np.float32(max_bx - min_b), np.float32(max_by - min_b)
    n_cols, n_rows = int(width / np.float32(35)), int(height / np.float32(35))
    w_cell, h_cell = int(math.ceil(width / n_cols)), int(math.ceil(height / n_rows))
    cells = []
    for i in range(n_rows):
        ini_y = min_b + i * h_cell
        max_y = ini_y + h_cell + 6
        if ini_y >= max_by - 3:
            continue
        max_y = min(max_y, max_by)
        for j in range(n_cols):
            ini_x = min_b + j * w_cell
            max_x = ini_x + w_cell + 6
            if ini_x >= max_bx - 6:
                continue
            max_x = min(max_x, max_bx)
            cells.append((i, j, ini_x, ini_y, max_x, max_y))
    return cells, w_cell, h_cell, n_cols, n_rows


def fast_candidates(level_img, ini_th=20, min_th=7):
    """Per-cell cv::FAST with the ini/min retry. Returns float32 n x 3 (x, y, response), reference order."""
    h, w = level_img.shape
    cells, w_cell, h_cell, _, _ = cell_grid(w, h)
    det_ini = cv2.FastFeatureDetector_create(ini_th, True)
    det_min = cv2.FastFeatureDetector_create(min_th, True)
    out = []
    retries = 0
    for (i, j, x0, y0, x1, y1) in cells:
        roi = level_img[y0:y1, x0:x1]
        kps = det_ini.detect(roi, None)
        if len(kps) == 0:
            kps = det_min.detect(roi, None)
            retries += 1
        for kp in kps:
            out.append((kp.pt[0] + j * w_cell, kp.pt[1] + i * h_cell, kp.response))
    return np.array(out, np.float32).reshape(-1, 3), retries


def blur(level_img):
    """GaussianBlur on a clone (ORBextractor.cc:1129-1130)."""
    m = level_img.copy()
    return cv2.GaussianBlur(m, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)


def ic_angle(level_img, x, y, umax):
    """IC_Angle (ORBextractor.cc:73-100) with cv2.fastAtan2."""
    m01 = 0
    m10 = 0
    for u in range(-15, 16):
        m10 += u * int(level_img[y, x + u])
    for v in range(1, 16):
        d = int(umax[v])
        vsum = 0
        for u in range(-d, d + 1):
            p, m = int(level_img[y + v, x + u]), int(level_img[y - v, x + u])
            vsum += p - m
            m10 += u * (p + m)
        m01 += v * vsum
    return cv2.fastAtan2(float(m01), float(m10))


def orb_descriptor(blurred, x, y, angle_deg, pattern):
    """computeOrbDescriptor (ORBextractor.cc:103-149) in float32 numpy with glibc cosf/sinf."""
    f32 = np.float32
    ang = f32(angle_deg) * f32(np.float64(np.pi) / np.float64(f32(180.0)))
    a = f32(_libm.cosf(ang))
    b = f32(_libm.sinf(ang))
    pat = np.asarray(pattern, np.int32).reshape(512, 2)
    px = pat[:, 0].astype(f32)
    py = pat[:, 1].astype(f32)
    ry = np.rint(px * b + py * a).astype(np.int32)
    rx = np.rint(px * a - py * b).astype(np.int32)
    vals = blurred[y + ry, x + rx].astype(np.int32)
    bits = (vals[0::2] < vals[1::2]).astype(np.uint8)
    return np.packbits(bits, bitorder="little")


def load_pattern(path):
    txt = open(path).read()
    nums = []
    for line in txt.splitlines():
        if line.startswith("//"):
            continue
        nums += [int(t) for t in line.replace(",", " ").split()]
    assert len(nums) == 1024
    return nums
