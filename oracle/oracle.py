"""ctypes wrapper around the CPU ORACLE (oracle/liborb_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package (visual_sgraphs_b200).
The library restates reference orb_slam3/src/{ORBextractor,ORBmatcher,Frame}.cc — see oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liborb_oracle.so")

KEYPOINT_DTYPE = np.dtype(
    [("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
     ("class_id", "<i4")]
)
assert KEYPOINT_DTYPE.itemsize == 28


def build(force=False):
    """Compile the oracle with its Makefile (g++ only)."""
    if force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
        for f in ("orb_oracle.cpp", "match_oracle.cpp", "oracle.h", "orb_pattern.inc", "Makefile")
    ):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborb_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float)
        vp = C.c_void_p
        L.orc_extractor_create.restype = vp
        L.orc_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_extractor_destroy.argtypes = [vp]
        L.orc_levels.argtypes = [vp]
        L.orc_scale_factors.argtypes = [vp, vp, vp, vp, vp]
        L.orc_quotas.argtypes = [vp, vp]
        L.orc_umax.argtypes = [vp, vp]
        L.orc_extract.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_num_keypoints.argtypes = [vp]
        L.orc_get_keypoints.argtypes = [vp, vp, vp]
        L.orc_level_size.argtypes = [vp, C.c_int, i32p, i32p]
        L.orc_get_level.argtypes = [vp, C.c_int, vp]
        L.orc_get_level_padded.argtypes = [vp, C.c_int, vp]
        L.orc_get_blurred.argtypes = [vp, C.c_int, vp]
        L.orc_num_candidates.argtypes = [vp, C.c_int]
        L.orc_get_candidates.argtypes = [vp, C.c_int, vp]
        L.orc_num_level_keypoints.argtypes = [vp, C.c_int]
        L.orc_get_level_keypoints.argtypes = [vp, C.c_int, vp]
        L.orc_resize_linear.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int]
        L.orc_gaussian_blur7.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_border_reflect101.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.orc_fast.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_cv_round_f.argtypes = [C.c_float]
        L.orc_cv_round_d.argtypes = [C.c_double]
        L.orc_ic_angle.restype = C.c_float
        L.orc_ic_angle.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        L.orc_orb_descriptor.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, vp]
        L.orc_distribute_octree.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_cvt_gray.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int]
        L.orc_remap_bilinear.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, vp, C.c_int]
        L.orc_remap_bilinear.restype = None
        L.orc_undistort_points.argtypes = [C.c_int, vp, C.c_double, C.c_double, C.c_double, C.c_double, vp, C.c_int, vp]
        L.orc_undistort_points.restype = None
        L.orc_stereo_from_rgbd.argtypes = [C.c_int, vp, vp, vp, C.c_int, C.c_float, vp, vp]
        L.orc_stereo_from_rgbd.restype = None
        L.orc_descriptor_distance.argtypes = [vp, vp]
        L.orc_get_features_in_area.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, vp, C.c_int]
        L.orc_three_maxima.argtypes = [vp, C.c_int, i32p, i32p, i32p]
        L.orc_search_by_projection_map.argtypes = [vp, vp, C.c_int, vp, vp, C.c_float, C.c_int, C.c_float, C.c_float, vp]
        L.orc_bench_knn2.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_bench_knn2.restype = C.c_double
        L.orc_search_by_projection_last_2cam.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, C.c_int, vp]
        L.orc_search_by_projection_map_2cam.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, C.c_float,
                                                         C.c_float, vp]
        L.orc_search_by_projection_last.argtypes = [vp, vp, C.c_int, vp, vp, C.c_float, C.c_int, C.c_int, vp]
        L.orc_search_for_initialization.argtypes = [vp, vp, vp, C.c_int, C.c_float, C.c_int, vp]
        L.orc_search_by_bow.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, vp]
        L.orc_search_by_bow_2cam.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, vp]
        L.orc_search_by_projection_reloc.argtypes = [vp, vp, C.c_int, vp, vp, C.c_float, C.c_int, C.c_int, vp]
        L.orc_search_by_projection_sim3.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, vp]
        L.orc_fuse_search.argtypes = [vp, C.c_int, vp, vp, C.c_float, vp, C.c_int, vp]
        L.orc_search_by_sim3.argtypes = [vp, vp, vp, vp, vp, vp, C.c_float, vp]
        L.orc_search_by_bow_kf.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, vp]
        L.orc_search_for_triangulation.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, C.c_int,
                                                   C.c_int, vp, vp, vp, C.c_int, vp]
        L.orc_distinctive_descriptor.argtypes = [vp, C.c_int]
        L.orc_bow_transform.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.orc_stereo_matches.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.c_float, vp, vp]
        L.orc_bench_extract.restype = C.c_double
        L.orc_bench_extract.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleExtractor:
    """Mirror of VS_GRAPHS::ORBextractor (ORBextractor.h:42-119) on the CPU oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7):
        self._L = lib()
        self._h = self._L.orc_extractor_create(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast)
        self.nlevels = nlevels

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_extractor_destroy(self._h)
            self._h = None

    def tables(self):
        n = self.nlevels
        s, i, s2, i2 = (np.zeros(n, np.float32) for _ in range(4))
        q = np.zeros(n, np.int32)
        u = np.zeros(16, np.int32)
        self._L.orc_scale_factors(self._h, _ptr(s), _ptr(i), _ptr(s2), _ptr(i2))
        self._L.orc_quotas(self._h, _ptr(q))
        self._L.orc_umax(self._h, _ptr(u))
        return dict(scale=s, inv_scale=i, sigma2=s2, inv_sigma2=i2, quota=q, umax=u)

    def __call__(self, image, lapping=(0, 0)):
        """operator(): returns (mono_index, keypoints structured array, descriptors n x 32)."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1
        h, w = image.shape
        mono = self._L.orc_extract(self._h, _ptr(image), w, h, image.strides[0], int(lapping[0]), int(lapping[1]))
        n = self._L.orc_num_keypoints(self._h)
        kps = np.zeros(n, KEYPOINT_DTYPE)
        desc = np.zeros((n, 32), np.uint8)
        if n:
            self._L.orc_get_keypoints(self._h, _ptr(kps), _ptr(desc))
        return mono, kps, desc

    def level_size(self, level):
        w, h = C.c_int32(), C.c_int32()
        self._L.orc_level_size(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self._L.orc_get_level(self._h, level, _ptr(out))
        return out

    def level_padded(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self._L.orc_get_level_padded(self._h, level, _ptr(out))
        return out

    def blurred(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        ok = self._L.orc_get_blurred(self._h, level, _ptr(out))
        return out if ok else None

    def candidates(self, level):
        n = self._L.orc_num_candidates(self._h, level)
        out = np.zeros((n, 3), np.float32)
        if n:
            self._L.orc_get_candidates(self._h, level, _ptr(out))
        return out

    def level_keypoints(self, level):
        n = self._L.orc_num_level_keypoints(self._h, level)
        out = np.zeros(n, KEYPOINT_DTYPE)
        if n:
            self._L.orc_get_level_keypoints(self._h, level, _ptr(out))
        return out


# ---- primitives -------------------------------------------------------------------------------

def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dw)
    return dst


def gaussian_blur7(src):
    src = np.ascontiguousarray(src)
    dst = np.zeros_like(src)
    lib().orc_gaussian_blur7(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dst.strides[0])
    return dst


def border_reflect101(src, border):
    src = np.ascontiguousarray(src)
    h, w = src.shape
    dst = np.zeros((h + 2 * border, w + 2 * border), np.uint8)
    lib().orc_border_reflect101(_ptr(src), w, h, src.strides[0], _ptr(dst), border, dst.strides[0])
    return dst


def fast(img, threshold):
    """FAST-9/16 + NMS on a (possibly strided) 2-D uint8 view. Returns int32 array n x 3 (x, y, score)."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    h, w = img.shape
    cap = max(1, (w * h) // 2)
    out = np.zeros((cap, 3), np.int32)
    n = lib().orc_fast(_ptr(img), w, h, img.strides[0], threshold, _ptr(out), cap)
    return out[:n].copy()


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def ic_angle(img, x, y):
    assert img.dtype == np.uint8 and img.strides[1] == 1
    return lib().orc_ic_angle(_ptr(img), img.strides[0], int(x), int(y))


def orb_descriptor(img, x, y, angle_deg):
    assert img.dtype == np.uint8 and img.strides[1] == 1
    out = np.zeros(32, np.uint8)
    lib().orc_orb_descriptor(_ptr(img), img.strides[0], int(x), int(y), float(angle_deg), _ptr(out))
    return out


def distribute_octree(xyr, min_x, max_x, min_y, max_y, quota):
    xyr = np.ascontiguousarray(xyr, np.float32)
    n = xyr.shape[0]
    out = np.zeros(max(n, 1), np.int32)
    m = lib().orc_distribute_octree(_ptr(xyr), n, min_x, max_x, min_y, max_y, quota, _ptr(out), out.size)
    return out[:m].copy()


def cvt_gray(img, rgb=True):
    """cv::cvtColor(img, COLOR_{RGB,BGR}[A]2GRAY) for an (h, w, 3|4) uint8 image."""
    img = np.ascontiguousarray(img, np.uint8)
    h, w, ch = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_cvt_gray(_ptr(img), w, h, img.strides[0], ch, int(bool(rgb)), _ptr(out), w)
    return out


def remap_bilinear(img, map_x, map_y):
    """cv::remap(img, map_x, map_y, INTER_LINEAR) for an 8-bit gray image and float32 maps (System.cc:284-292)."""
    img = np.ascontiguousarray(img, np.uint8)
    map_x = np.ascontiguousarray(map_x, np.float32)
    map_y = np.ascontiguousarray(map_y, np.float32)
    h, w = map_x.shape
    out = np.zeros((h, w), np.uint8)
    lib().orc_remap_bilinear(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(map_x), _ptr(map_y), w, h, _ptr(out), w)
    return out


def undistort_points(xy, fx, fy, cx, cy, dist):
    """Frame::UndistortKeyPoints: cv::undistortPoints(xy, K, dist, R=I, P=K) on (n, 2) float32 points."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    dist = np.ascontiguousarray(dist, np.float64).ravel()
    out = np.zeros_like(xy)
    lib().orc_undistort_points(xy.shape[0], _ptr(xy), fx, fy, cx, cy, _ptr(dist), dist.size, _ptr(out))
    return out


def stereo_from_rgbd(xy, xy_un, depth, bf):
    """Frame::ComputeStereoFromRGBD: (u_right, depth) per keypoint from a float32 depth image."""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    xy_un = np.ascontiguousarray(xy_un, np.float32).reshape(-1, 2)
    depth = np.ascontiguousarray(depth, np.float32)
    ur = np.zeros(xy.shape[0], np.float32)
    dz = np.zeros(xy.shape[0], np.float32)
    lib().orc_stereo_from_rgbd(xy.shape[0], _ptr(xy), _ptr(xy_un), _ptr(depth), depth.shape[1], bf, _ptr(ur), _ptr(dz))
    return ur, dz


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_ptr(a), _ptr(b))


def bench_knn2(query, train, threads=1, variant=0):
    """CPU brute-force kNN-2 (seconds, idx (nq, 2), dist (nq, 2)); variant 0 = bit-hack distance, 1 = popcount."""
    query = np.ascontiguousarray(query, np.uint8)
    train = np.ascontiguousarray(train, np.uint8)
    idx = np.zeros((len(query), 2), np.int32)
    dist = np.zeros((len(query), 2), np.int32)
    secs = lib().orc_bench_knn2(_ptr(query), len(query), _ptr(train), len(train), threads, variant, _ptr(idx), _ptr(dist))
    return secs, idx, dist


def bench_extract(frames, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, threads=1):
    """Times the oracle extractor over a stack of frames (n, h, w) uint8. Returns (seconds, total_keypoints)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    total = C.c_int64(0)
    secs = lib().orc_bench_extract(_ptr(frames), n, w, h, nfeatures, scale_factor, nlevels, ini_th, min_th, threads,
                                   C.byref(total))
    return secs, total.value


# ---- matcher methods on flattened views; `view` is a ctypes struct with the orc_frame_view layout -----------
# (visual_sgraphs_b200._lib.FrameView has that layout; tests pass the same buffers to both sides)

def get_features_in_area(view, x, y, r, min_level=-1, max_level=-1):
    out = np.zeros(max(view.n, 1), np.int32)
    n = lib().orc_get_features_in_area(C.addressof(view), x, y, r, min_level, max_level, _ptr(out), out.size)
    return out[:n].copy()


def three_maxima(sizes):
    sizes = np.ascontiguousarray(sizes, np.int32)
    a, b, c = C.c_int32(-1), C.c_int32(-1), C.c_int32(-1)
    lib().orc_three_maxima(_ptr(sizes), len(sizes), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def search_by_projection_map(view, occupied, pts, mp_desc, th, far_points, th_far, nnratio):
    occupied = np.ascontiguousarray(occupied, np.uint8)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    assign = np.zeros(view.n, np.int32)
    nm = lib().orc_search_by_projection_map(C.addressof(view), _ptr(occupied), len(pts), _ptr(pts), _ptr(mp_desc), th,
                                            int(far_points), th_far, nnratio, _ptr(assign))
    return nm, assign


def search_by_projection_map_2cam(view_l, view_r, occupied, l2r, r2l, pts_l, pts_r, mp_desc, th, far_points, th_far, nnratio):
    """ORBmatcher.cc:42-216 on a two-camera frame; slots [0, nL) are the left keypoints, [nL, nL + nR) the right ones."""
    occupied = np.ascontiguousarray(occupied, np.uint8)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    l2r = np.ascontiguousarray(l2r, np.int32)
    r2l = np.ascontiguousarray(r2l, np.int32)
    assign = np.zeros(view_l.n + view_r.n, np.int32)
    nm = lib().orc_search_by_projection_map_2cam(C.addressof(view_l), C.addressof(view_r), _ptr(occupied), _ptr(l2r), _ptr(r2l),
                                                 len(pts_l), _ptr(pts_l), _ptr(pts_r), _ptr(mp_desc), th, int(far_points),
                                                 th_far, nnratio, _ptr(assign))
    return nm, assign


def search_by_projection_last_2cam(view_l, view_r, occupied, pts_l, pts_r, desc, th, mode, check_ori):
    occupied = np.ascontiguousarray(occupied, np.uint8)
    desc = np.ascontiguousarray(desc, np.uint8)
    assign = np.zeros(view_l.n + view_r.n, np.int32)
    nm = lib().orc_search_by_projection_last_2cam(C.addressof(view_l), C.addressof(view_r), _ptr(occupied), len(pts_l),
                                                  _ptr(pts_l), _ptr(pts_r), _ptr(desc), th, mode, int(check_ori), _ptr(assign))
    return nm, assign


def search_by_projection_last(view, occupied, pts, desc, th, mode, check_ori):
    occupied = np.ascontiguousarray(occupied, np.uint8)
    desc = np.ascontiguousarray(desc, np.uint8)
    assign = np.zeros(view.n, np.int32)
    nm = lib().orc_search_by_projection_last(C.addressof(view), _ptr(occupied), len(pts), _ptr(pts), _ptr(desc), th,
                                             mode, int(check_ori), _ptr(assign))
    return nm, assign


def search_for_initialization(view1, view2, prev_matched, window, nnratio, check_ori):
    m12 = np.zeros(view1.n, np.int32)
    nm = lib().orc_search_for_initialization(C.addressof(view1), C.addressof(view2), _ptr(prev_matched), window,
                                             nnratio, int(check_ori), _ptr(m12))
    return nm, m12


def search_by_bow(kf_view, kf_mp_valid, f_view, kf_fv, f_fv, nnratio, check_ori, f_nleft=-1):
    """f_nleft != -1: F is a two-camera frame whose first f_nleft features belong to the left camera."""
    kn, kp, ki = (np.ascontiguousarray(a, np.int32) for a in kf_fv)
    fn, fp, fi = (np.ascontiguousarray(a, np.int32) for a in f_fv)
    valid = np.ascontiguousarray(kf_mp_valid, np.uint8)
    out = np.zeros(f_view.n, np.int32)
    nm = lib().orc_search_by_bow_2cam(C.addressof(kf_view), _ptr(valid), C.addressof(f_view), int(f_nleft), len(kn), _ptr(kn),
                                      _ptr(kp), _ptr(ki), len(fn), _ptr(fn), _ptr(fp), _ptr(fi), nnratio, int(check_ori),
                                      _ptr(out))
    return nm, out


def stereo_matches(ex_left, ex_right, keys_l, desc_l, keys_r, desc_r, mb, mbf):
    """Frame::ComputeStereoMatches on two OracleExtractor objects (their last call's pyramids)."""
    keys_l = np.ascontiguousarray(keys_l, KEYPOINT_DTYPE)
    keys_r = np.ascontiguousarray(keys_r, KEYPOINT_DTYPE)
    desc_l = np.ascontiguousarray(desc_l, np.uint8)
    desc_r = np.ascontiguousarray(desc_r, np.uint8)
    u_right = np.zeros(len(keys_l), np.float32)
    depth = np.zeros(len(keys_l), np.float32)
    lib().orc_stereo_matches(ex_left._h, ex_right._h, _ptr(keys_l), _ptr(desc_l), len(keys_l), _ptr(keys_r), _ptr(desc_r),
                             len(keys_r), mb, mbf, _ptr(u_right), _ptr(depth))
    return u_right, depth


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def search_by_projection_reloc(view, occupied, pts, desc, th, orb_dist, check_ori):
    occupied, desc = _u8(occupied), _u8(desc)
    assign = np.zeros(view.n, np.int32)
    nm = lib().orc_search_by_projection_reloc(C.addressof(view), _ptr(occupied), len(pts), _ptr(pts), _ptr(desc), th,
                                              int(orb_dist), int(check_ori), _ptr(assign))
    return nm, assign


def search_by_projection_sim3(view, matched, pts, desc, th, ratio_hamming):
    matched, desc = _u8(matched), _u8(desc)
    assign = np.zeros(view.n, np.int32)
    nm = lib().orc_search_by_projection_sim3(C.addressof(view), _ptr(matched), len(pts), _ptr(pts), _ptr(desc), int(th),
                                             ratio_hamming, _ptr(assign))
    return nm, assign


def fuse_search(view, pts, desc, th, inv_level_sigma2, variant):
    desc = _u8(desc)
    sig = np.ascontiguousarray(inv_level_sigma2, np.float32)
    best = np.zeros(max(len(pts), 1), np.int32)
    nf = lib().orc_fuse_search(C.addressof(view), len(pts), _ptr(pts), _ptr(desc), th, _ptr(sig), int(variant), _ptr(best))
    return nf, best[:len(pts)]


def search_by_sim3(view1, view2, pts1, desc1, pts2, desc2, th):
    desc1, desc2 = _u8(desc1), _u8(desc2)
    m12 = np.zeros(max(view1.n, 1), np.int32)
    nf = lib().orc_search_by_sim3(C.addressof(view1), C.addressof(view2), _ptr(pts1), _ptr(desc1), _ptr(pts2), _ptr(desc2),
                                  th, _ptr(m12))
    return nf, m12[:view1.n]


def search_by_bow_kf(view1, valid1, view2, valid2, fv1, fv2, nnratio, check_ori):
    n1, p1, i1 = (np.ascontiguousarray(a, np.int32) for a in fv1)
    n2, p2, i2 = (np.ascontiguousarray(a, np.int32) for a in fv2)
    valid1, valid2 = _u8(valid1), _u8(valid2)
    out = np.zeros(max(view1.n, 1), np.int32)
    nm = lib().orc_search_by_bow_kf(C.addressof(view1), _ptr(valid1), C.addressof(view2), _ptr(valid2), len(n1), _ptr(n1),
                                    _ptr(p1), _ptr(i1), len(n2), _ptr(n2), _ptr(p2), _ptr(i2), nnratio, int(check_ori),
                                    _ptr(out))
    return nm, out[:view1.n]


def search_for_triangulation(view1, has1, view2, has2, fv1, fv2, only_stereo, coarse, f12, ep, level_sigma2_2, check_ori):
    n1, p1, i1 = (np.ascontiguousarray(a, np.int32) for a in fv1)
    n2, p2, i2 = (np.ascontiguousarray(a, np.int32) for a in fv2)
    has1, has2 = _u8(has1), _u8(has2)
    f12 = np.ascontiguousarray(f12, np.float32).reshape(9)
    ep = np.ascontiguousarray(ep, np.float32).reshape(2)
    sig = np.ascontiguousarray(level_sigma2_2, np.float32)
    out = np.zeros(max(view1.n, 1), np.int32)
    nm = lib().orc_search_for_triangulation(C.addressof(view1), _ptr(has1), C.addressof(view2), _ptr(has2), len(n1),
                                            _ptr(n1), _ptr(p1), _ptr(i1), len(n2), _ptr(n2), _ptr(p2), _ptr(i2),
                                            int(only_stereo), int(coarse), _ptr(f12), _ptr(ep), _ptr(sig), int(check_ori),
                                            _ptr(out))
    return nm, out[:view1.n]


def distinctive_descriptor(desc):
    """MapPoint::ComputeDistinctiveDescriptors for one point: (n, 32) uint8 -> BestIdx (-1 if n == 0)."""
    desc = _u8(desc).reshape(-1, 32)
    return lib().orc_distinctive_descriptor(_ptr(desc), len(desc))


def bow_transform(child_ptr, child_idx, node_desc, levels, desc, levelsup=4):
    """DBoW2 tree walk per descriptor -> (leaf node, FeatureVector node)."""
    child_ptr = np.ascontiguousarray(child_ptr, np.int32)
    child_idx = np.ascontiguousarray(child_idx, np.int32)
    node_desc, desc = _u8(node_desc).reshape(-1, 32), _u8(desc).reshape(-1, 32)
    leaf = np.zeros(max(len(desc), 1), np.int32)
    nid = np.zeros(max(len(desc), 1), np.int32)
    lib().orc_bow_transform(_ptr(child_ptr), _ptr(child_idx), _ptr(node_desc), int(levels), _ptr(desc), len(desc),
                            int(levelsup), _ptr(leaf), _ptr(nid))
    return leaf[:len(desc)], nid[:len(desc)]
