"""Pins the C++ oracle's restated OpenCV primitives to real OpenCV (python cv2) — SURVEY §8c.

CPU-only.  Every check is bit-exact.
"""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cv2_ref
from visual_sgraphs_b200.synth import synth_frame


def _rand_img(seed, w, h, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)
    if kind == "const":
        return np.full((h, w), 255, np.uint8)
    raise ValueError(kind)


def test_cv_round_half_even(oracle):
    L = oracle.lib()
    for v in (0.5, 1.5, 2.5, -0.5, -1.5, 3.49999, 1e6 + 0.5, 266.5, 399.99997):
        assert L.orc_cv_round_f(v) == int(np.rint(np.float32(v)))
        assert L.orc_cv_round_d(v) == int(np.rint(v))


@pytest.mark.parametrize("shape", [((640, 480), (533, 400)), ((533, 400), (444, 333)), ((444, 333), (370, 278)),
                                   ((752, 480), (627, 400)), ((1280, 720), (1067, 600)), ((214, 161), (179, 134)),
                                   ((101, 77), (84, 64)), ((257, 193), (214, 161))])
@pytest.mark.parametrize("kind", ["uniform", "binary"])
def test_resize_matches_cv2(oracle, shape, kind):
    (sw, sh), (dw, dh) = shape
    src = _rand_img(sw * 31 + dh, sw, sh, kind)
    want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
    got = oracle.resize_linear(src, dw, dh)
    assert np.array_equal(got, want)


def test_resize_from_strided_roi(oracle):
    big = _rand_img(5, 700, 500)
    roi = big[19:19 + 400, 19:19 + 533]
    want = cv2.resize(roi, (444, 333), interpolation=cv2.INTER_LINEAR)
    dst = np.zeros((333, 444), np.uint8)
    oracle.lib().orc_resize_linear(roi.ctypes.data, 533, 400, roi.strides[0], dst.ctypes.data, 444, 333, 444)
    assert np.array_equal(dst, want)


@pytest.mark.parametrize("wh", [(640, 480), (533, 400), (179, 134), (752, 480), (357, 201), (97, 131), (40, 41)])
@pytest.mark.parametrize("kind", ["uniform", "binary", "const"])
def test_gaussian_blur_matches_cv2(oracle, wh, kind):
    w, h = wh
    src = _rand_img(w + h, w, h, kind)
    want = cv2_ref.blur(src)
    got = oracle.gaussian_blur7(src)
    assert np.array_equal(got, want)


def test_border_matches_cv2(oracle):
    src = _rand_img(3, 61, 47)
    want = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(oracle.border_reflect101(src, 19), want)


@pytest.mark.parametrize("threshold", [20, 7, 1, 40])
def test_fast_matches_cv2_on_cells(oracle, threshold):
    frame = synth_frame(77, 320, 240)
    noise = _rand_img(9, 320, 240)
    det = cv2.FastFeatureDetector_create(threshold, True)
    rng = np.random.default_rng(threshold)
    ncorners = 0
    for img in (frame, noise):
        for _ in range(40):
            cw, ch = rng.integers(7, 60, 2)
            x0 = int(rng.integers(0, 320 - cw))
            y0 = int(rng.integers(0, 240 - ch))
            roi = img[y0:y0 + ch, x0:x0 + cw]
            kps = det.detect(roi, None)
            want = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kps], np.int32).reshape(-1, 3)
            got = oracle.fast(roi, threshold)
            assert np.array_equal(got, want), (x0, y0, cw, ch)
            ncorners += len(kps)
    assert ncorners > 50


def test_fast_tiny_images(oracle):
    for w, h in ((6, 30), (30, 6), (7, 7), (3, 3)):
        img = _rand_img(w * h, w, h)
        kps = cv2.FastFeatureDetector_create(7, True).detect(img, None)
        got = oracle.fast(img, 7)
        assert len(got) == len(kps)


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.default_rng(0)
    ys = rng.integers(-200000, 200000, 20000)
    xs = rng.integers(-200000, 200000, 20000)
    for y, x in zip(ys[:5000], xs[:5000]):
        assert oracle.fast_atan2(y, x) == cv2.fastAtan2(float(y), float(x))
    for y, x in ((0, 0), (0, 5), (5, 0), (0, -5), (-5, 0), (7, 7), (-7, 7), (7, -7), (-7, -7), (1, 100000)):
        assert oracle.fast_atan2(y, x) == cv2.fastAtan2(float(y), float(x))


def test_pyramid_matches_cv2_chain(oracle):
    frame = synth_frame(1000)
    ex = oracle.OracleExtractor()
    ex(frame)
    want = cv2_ref.pyramid(frame)
    for level in range(8):
        assert np.array_equal(ex.level_padded(level), want[level]), level


@pytest.mark.parametrize("wh,nfeat", [((640, 480), 1000), ((752, 480), 1200), ((322, 243), 500)])
def test_fast_candidates_match_cv2_per_cell(oracle, wh, nfeat):
    frame = synth_frame(1234, *wh)
    ex = oracle.OracleExtractor(nfeat)
    ex(frame)
    total_retries = 0
    for level in range(8):
        img = ex.level_padded(level)[19:-19, 19:-19]
        # the reference runs FAST on ROIs of the padded level; cv2 on a view of the same memory
        padded = ex.level_padded(level)
        view = padded[19:-19, 19:-19]
        want, retries = cv2_ref.fast_candidates(view)
        total_retries += retries
        got = ex.candidates(level)
        assert got.shape == want.shape, level
        assert np.array_equal(got, want), level
        assert np.array_equal(img, view)
    assert total_retries >= 0


def test_low_contrast_frame_forces_min_threshold(oracle):
    frame = (synth_frame(5, 320, 240).astype(np.float32) * 0.12 + 100).astype(np.uint8)
    ex = oracle.OracleExtractor(500)
    ex(frame)
    want, retries = cv2_ref.fast_candidates(ex.level(0))
    assert retries > 10
    assert np.array_equal(ex.candidates(0), want)


def test_blur_and_angles_and_descriptors_match_cv2(oracle):
    frame = synth_frame(4321)
    ex = oracle.OracleExtractor()
    mono, kps, desc = ex(frame)
    umax = ex.tables()["umax"]
    import os
    pattern = cv2_ref.load_pattern(os.path.join(os.path.dirname(cv2_ref.__file__), "orb_pattern.inc"))
    row = 0
    for level in range(8):
        lk = ex.level_keypoints(level)
        img = ex.level(level)
        b = ex.blurred(level)
        assert np.array_equal(b, cv2_ref.blur(img))
        step = max(1, len(lk) // 12)
        for i in range(0, len(lk), step):
            x, y = int(lk["x"][i]), int(lk["y"][i])
            assert lk["angle"][i] == np.float32(cv2_ref.ic_angle(img, x, y, umax))
            d = cv2_ref.orb_descriptor(b, x, y, lk["angle"][i], pattern)
            assert np.array_equal(d, desc[row + i]), (level, i)
        row += len(lk)
    assert row == len(kps) and mono == len(kps)


@pytest.mark.parametrize("channels", [3, 4])
@pytest.mark.parametrize("rgb", [True, False])
def test_cvt_gray_matches_cv2(oracle, channels, rgb):
    """Tracking::GrabImageRGBD's cvtColor (Tracking.cc:1595-1608) for RGB / BGR / RGBA / BGRA inputs."""
    rng = np.random.default_rng(channels * 2 + rgb)
    code = {(3, True): cv2.COLOR_RGB2GRAY, (3, False): cv2.COLOR_BGR2GRAY, (4, True): cv2.COLOR_RGBA2GRAY,
            (4, False): cv2.COLOR_BGRA2GRAY}[(channels, rgb)]
    for shape in ((480, 640), (97, 131), (5, 3)):
        img = rng.integers(0, 256, shape + (channels,), dtype=np.uint8)
        assert np.array_equal(oracle.cvt_gray(img, rgb), cv2.cvtColor(img, code))
    ramp = np.stack(list(np.meshgrid(np.arange(256), np.arange(256), indexing="ij")) + [np.full((256, 256), 77)], -1).astype(np.uint8)
    if channels == 4:
        ramp = np.concatenate([ramp, ramp[..., :1]], -1)
    assert np.array_equal(oracle.cvt_gray(ramp, rgb), cv2.cvtColor(np.ascontiguousarray(ramp), code))


# Calibrations of the reference's own configs: TUM1 (config/RGB-D/TUM1.yaml:9-19), EuRoC (config/Stereo/EuRoC.yaml), and an
# 8-coefficient rational model; values are float32 like the reference's mK / mDistCoef (Tracking.cc ParseCamParamFile).
CALIBRATIONS = [
    ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
    ((600.0, 601.5, 320.25, 239.75), (0.1, -0.2, 0.001, 0.002, 0.05, 0.01, -0.02, 0.003)),
    ((300.0, 300.0, 320.0, 240.0), (-2.5, 6.0, 0.0, 0.0, -5.0)),       # icdist < 0 far from the centre (regression_14583)
]


@pytest.mark.parametrize("calib", range(len(CALIBRATIONS)))
def test_undistort_points_matches_cv2(oracle, calib):
    """Frame::UndistortKeyPoints (Frame.cc:891-922): cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK)."""
    (fx, fy, cx, cy), dist = CALIBRATIONS[calib]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32).astype(np.float64)
    dist = np.array(dist, np.float32).astype(np.float64)
    rng = np.random.default_rng(40 + calib)
    pts = np.stack([rng.uniform(-20, 780, 20000), rng.uniform(-20, 520, 20000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]          # ComputeImageBounds' corners (:928-931)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
    got = oracle.undistort_points(pts, K[0, 0], K[1, 1], K[0, 2], K[1, 2], dist)
    assert np.array_equal(ref, got)


def test_undistort_points_shortcut_and_rgbd_association(oracle):
    """k1 == 0 copies the keypoints (Frame.cc:893-897); ComputeStereoFromRGBD (:1129-1150) against a direct restatement."""
    rng = np.random.default_rng(9)
    pts = rng.uniform(0, 640, (100, 2)).astype(np.float32)
    assert np.array_equal(oracle.undistort_points(pts, 500, 500, 320, 240, [0.0, 0.3, 0.0, 0.0]), pts)
    assert np.array_equal(oracle.undistort_points(pts, 500, 500, 320, 240, []), pts)
    xy = np.stack([rng.uniform(0, 639.9, 500), rng.uniform(0, 479.9, 500)], 1).astype(np.float32)
    xy_un = (xy + rng.normal(0, 1, xy.shape)).astype(np.float32)
    depth = rng.uniform(-0.5, 6.0, (480, 640)).astype(np.float32)
    ur, dz = oracle.stereo_from_rgbd(xy, xy_un, depth, 40.0)
    for i in range(len(xy)):
        d = depth[int(xy[i, 1]), int(xy[i, 0])]
        if d > 0:
            assert dz[i] == d and ur[i] == np.float32(xy_un[i, 0] - np.float32(40.0) / d)
        else:
            assert dz[i] == -1 and ur[i] == -1


def _euroc_rectify_maps(right=False):
    """cv::initUndistortRectifyMap maps like Settings.cc:571-574 builds them for config/Stereo/EuRoC.yaml (cam0 / cam1
    intrinsics of that file; a small rectifying rotation and a common new projection)."""
    if right:
        K = np.array([[457.587, 0, 379.999], [0, 456.134, 255.238], [0, 0, 1]])
        D = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])
        rvec = np.array([-0.004, 0.012, -0.002])
    else:
        K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]])
        D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
        rvec = np.array([0.003, -0.011, 0.0015])
    R = cv2.Rodrigues(rvec)[0]
    P = np.array([[435.2046959714599, 0, 367.4517211914062], [0, 435.2046959714599, 252.2008514404297], [0, 0, 1]])
    return cv2.initUndistortRectifyMap(K, D, R, P, (752, 480), cv2.CV_32F)


def test_remap_bilinear_matches_cv2(oracle):
    """cv::remap(im, imToFeed, M1, M2, INTER_LINEAR) of System::TrackStereo (System.cc:284-292) on 8-bit frames."""
    from visual_sgraphs_b200.synth import synth_frame
    img = synth_frame(5, 752, 480)
    for right in (False, True):
        mx, my = _euroc_rectify_maps(right)
        assert np.array_equal(oracle.remap_bilinear(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    rng = np.random.default_rng(3)
    mx = rng.uniform(-5, 760, (300, 400)).astype(np.float32)          # taps outside the source on every side
    my = rng.uniform(-5, 490, (300, 400)).astype(np.float32)
    assert np.array_equal(oracle.remap_bilinear(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    mx = (np.arange(641)[None, :] * 1.171875 + 1 / 64).astype(np.float32).repeat(361, 0)    # exact ties of the 1/32 quantiser
    my = (np.arange(361)[:, None] * 1.3 + 0.015625).astype(np.float32).repeat(641, 1)
    assert np.array_equal(oracle.remap_bilinear(img, mx, my), cv2.remap(img, mx, my, cv2.INTER_LINEAR))
    noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    mx = rng.uniform(-40000, 40000, (50, 60)).astype(np.float32)      # saturation of the integer part
    my = rng.uniform(-3, 100, (50, 60)).astype(np.float32)
    assert np.array_equal(oracle.remap_bilinear(noise, mx, my), cv2.remap(noise, mx, my, cv2.INTER_LINEAR))


def test_primitives_match_cv2_on_random_shapes(oracle):
    """Randomised sweep over image sizes and scale factors (seeded): resize at the extractor's level ratios 1.1 .. 1.9,
    the 7x7 blur, the colour conversion and remap, all against cv2 — the shapes the fixed parametrisations do not hit
    (odd sizes, widths that are not multiples of 4, very small levels)."""
    rng = np.random.default_rng(2026)
    for trial in range(40):
        w, h = int(rng.integers(67, 400)), int(rng.integers(67, 300))
        kind = ("uniform", "binary")[trial % 2]
        src = _rand_img(1000 + trial, w, h, kind)
        scale = float(rng.uniform(1.1, 1.9))
        dw, dh = int(round(w / scale)), int(round(h / scale))
        assert np.array_equal(oracle.resize_linear(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)), (w, h, dw, dh)
        assert np.array_equal(oracle.gaussian_blur7(src), cv2_ref.blur(src)), (w, h)
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(oracle.cvt_gray(rgb, True), cv2.cvtColor(rgb, cv2.COLOR_RGB2GRAY))
        mx = (np.arange(dw)[None, :] * scale + rng.uniform(-3, 3)).astype(np.float32).repeat(dh, 0) + rng.normal(0, 0.3, (dh, dw)).astype(np.float32)
        my = (np.arange(dh)[:, None] * scale + rng.uniform(-3, 3)).astype(np.float32).repeat(dw, 1) + rng.normal(0, 0.3, (dh, dw)).astype(np.float32)
        assert np.array_equal(oracle.remap_bilinear(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR)), (w, h)
