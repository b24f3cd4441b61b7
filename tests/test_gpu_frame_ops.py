"""GPU parity tests for the per-keypoint steps Frame's constructors run right after the extractor:
Frame::UndistortKeyPoints / ComputeImageBounds (cv::undistortPoints, Frame.cc:891-955) through the C ABI against the
oracle (itself pinned bit-exact against cv2 in tests/test_oracle_cv2.py), and the RGB-D depth association."""
import numpy as np
import pytest

from visual_sgraphs_b200.synth import synth_frame

pytestmark = pytest.mark.gpu

CALIBRATIONS = [
    ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),   # TUM1.yaml
    ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),              # EuRoC.yaml
    ((600.0, 601.5, 320.25, 239.75), (0.1, -0.2, 0.001, 0.002, 0.05, 0.01, -0.02, 0.003)),
    ((300.0, 300.0, 320.0, 240.0), (-2.5, 6.0, 0.0, 0.0, -5.0)),                                                  # icdist < 0
]


def _calib(i):
    (fx, fy, cx, cy), dist = CALIBRATIONS[i]
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float32).astype(np.float64)
    return K, np.array(dist, np.float32).astype(np.float64)


@pytest.mark.parametrize("calib", range(len(CALIBRATIONS)))
def test_undistort_keypoints_bit_exact(oracle, calib):
    from visual_sgraphs_b200.matcher import ORBmatcher
    K, dist = _calib(calib)
    rng = np.random.default_rng(60 + calib)
    pts = np.stack([rng.uniform(-20, 780, 50000), rng.uniform(-20, 520, 50000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    got = ORBmatcher().UndistortKeyPoints(pts, K, dist)
    want = oracle.undistort_points(pts, K[0, 0], K[1, 1], K[0, 2], K[1, 2], dist)
    assert np.array_equal(got, want)
    assert not np.array_equal(got, pts)


def test_undistort_shortcut_empty_and_errors():
    from visual_sgraphs_b200 import _lib
    from visual_sgraphs_b200.matcher import ORBmatcher
    m = ORBmatcher()
    K, _ = _calib(0)
    pts = np.random.default_rng(1).uniform(0, 600, (300, 2)).astype(np.float32)
    assert np.array_equal(m.UndistortKeyPoints(pts, K, [0.0, 0.5, 0.0, 0.0]), pts)     # Frame.cc:893-897
    assert np.array_equal(m.UndistortKeyPoints(pts, K, []), pts)
    assert m.UndistortKeyPoints(np.zeros((0, 2), np.float32), K, [0.1, 0, 0, 0]).shape == (0, 2)
    bad = K.copy()
    bad[0, 0] = 0.0
    with pytest.raises(_lib.VsgError):
        m.UndistortKeyPoints(pts, bad, [0.1, 0, 0, 0])
    with pytest.raises(_lib.VsgError):
        m.UndistortKeyPoints(pts, K, np.zeros(13))


def test_rgbd_frame_chain(oracle):
    """The RGB-D constructor's chain (Frame.cc:168-190) on a batch: extraction, undistortion of the device-resident
    keypoints, depth association — against the oracle step by step."""
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    K, dist = _calib(0)
    frames = np.stack([synth_frame(800 + i, 640, 480) for i in range(3)] + [np.full((480, 640), 90, np.uint8)])
    ex = ORBextractor(1000, max_batch=len(frames))
    res = ex.extract_batch(frames)
    m = ORBmatcher()
    un = m.UndistortKeyPointsBatch(ex, len(frames), K, dist)
    un_id = m.UndistortKeyPointsBatch(ex, len(frames), K, [0.0, 0.0, 0.0, 0.0])
    rng = np.random.default_rng(4)
    for f, (_, kps, _) in enumerate(res):
        n = len(kps)
        xy = np.stack([kps["x"], kps["y"]], 1) if n else np.zeros((0, 2), np.float32)
        want = oracle.undistort_points(xy, K[0, 0], K[1, 1], K[0, 2], K[1, 2], dist)
        assert np.array_equal(un[f, :n], want)
        assert np.all(un[f, n:] == -1)
        assert np.array_equal(un_id[f, :n], xy)
        assert np.array_equal(m.UndistortKeyPoints(xy, K, dist), want)
        depth = rng.uniform(-0.5, 6.0, (480, 640)).astype(np.float32)
        ur, dz = m.ComputeStereoFromRGBD(xy, want, depth, 40.0)
        wur, wdz = oracle.stereo_from_rgbd(xy, want, depth, 40.0)
        assert np.array_equal(ur, wur) and np.array_equal(dz, wdz)
    assert len(res[-1][1]) == 0
    with pytest.raises(Exception):
        m.UndistortKeyPointsBatch(ex, len(frames) + 1, K, dist)


def _rectify_maps(w, h, sw, sh, k1, shift):
    """Smooth float32 maps of the kind cv::initUndistortRectifyMap produces (radial term + a small rotation / shift),
    partly pointing outside the source so that the constant border is exercised."""
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    xn, yn = (xs - w / 2) / (0.6 * w), (ys - h / 2) / (0.6 * w)
    r2 = xn * xn + yn * yn
    f = 1 + k1 * r2
    mx = (xn * f * np.cos(0.01) - yn * f * np.sin(0.01)) * 0.6 * w + sw / 2 + shift
    my = (xn * f * np.sin(0.01) + yn * f * np.cos(0.01)) * 0.6 * w + sh / 2 - shift / 2
    return mx.astype(np.float32), my.astype(np.float32)


def test_rectified_stereo_chain(oracle):
    """BASELINE config 2 with the rectification System::TrackStereo runs first (System.cc:284-292): unrectified left /
    right frames interleaved in one batch, remapped on the device with each camera's map, extracted, stereo-matched —
    against oracle remap + oracle extraction + oracle stereo matching."""
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    from visual_sgraphs_b200.synth import synth_stereo_pair
    sw, sh, w, h = 752, 480, 736, 464                     # the rectified size may differ from the sensor's (newImSize)
    maps = [_rectify_maps(w, h, sw, sh, 0.12, 2.5), _rectify_maps(w, h, sw, sh, 0.10, -1.5)]
    pairs = [synth_stereo_pair(70 + i, sw, sh) for i in range(2)]
    stack = np.stack([im for pr in pairs for im in pr])
    ex = ORBextractor(1200, max_batch=len(stack))
    for slot, (mx, my) in enumerate(maps):
        ex.set_rectify_map(slot, mx, my)
    res = ex.extract_batch_rectify(stack, ncameras=2)
    m = ORBmatcher()
    ub, db = m.ComputeStereoMatchesBatch(ex, len(pairs), 0.11, 47.9)
    for p, pr in enumerate(pairs):
        rect = [oracle.remap_bilinear(im, *maps[c]) for c, im in enumerate(pr)]
        assert (rect[0] == 0).any() and rect[0].std() > 10           # border pixels present, content preserved
        oxl, oxr = oracle.OracleExtractor(1200), oracle.OracleExtractor(1200)
        _, okl, odl = oxl(rect[0])
        _, okr, odr = oxr(rect[1])
        (_, kl, dl), (_, kr, dr) = res[2 * p], res[2 * p + 1]
        assert kl.tobytes() == okl.tobytes() and np.array_equal(dl, odl)
        assert kr.tobytes() == okr.tobytes() and np.array_equal(dr, odr)
        wu, wd = oracle.stereo_matches(oxl, oxr, okl, odl, okr, odr, 0.11, 47.9)
        assert np.array_equal(ub[p, :len(kl)], wu) and np.array_equal(db[p, :len(kl)], wd)
    # one camera (monocular / RGB-D rectification), a chunk boundary inside the batch is covered by the extractor tests
    ex1 = ORBextractor(1000, max_batch=3)
    ex1.set_rectify_map(0, *maps[1])
    frames = np.stack([pairs[0][0], pairs[1][1], pairs[0][1]])
    for (_, k, d), im in zip(ex1.extract_batch_rectify(frames), frames):
        _, ok, od = oracle.OracleExtractor(1000)(oracle.remap_bilinear(im, *maps[1]))
        assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    with pytest.raises(Exception):
        ex1.extract_batch_rectify(frames, ncameras=2)                 # no map for the second camera
