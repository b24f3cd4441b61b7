"""Threading contract of the boundary: extractor instances run concurrently (left / right images in two
std::threads, Frame.cc:129-132) and ORBmatcher objects are used at the same time from Tracking, LocalMapping and
LoopClosing.  Several host threads, each with its own handles, run extraction and different Search* methods at once;
every result must equal the single-threaded one (and a failing call in one thread must not leak its error text into
another thread's vsg_last_error)."""
import threading

import numpy as np
import pytest

from tests import match_scenarios as sc
from visual_sgraphs_b200.synth import synth_frame, synth_query_train

pytestmark = pytest.mark.gpu


def test_concurrent_handles_match_single_threaded_results(oracle):
    from visual_sgraphs_b200 import _lib
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da)
    pts, desc, occ = sc.track_points(fd, kb, db, (9, 5), 21, False)
    frames = [synth_frame(900 + i, 640, 480) for i in range(4)]
    q, t = synth_query_train(5, 500, 20000)
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32).astype(np.float64)
    dist = np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32).astype(np.float64)
    xy = np.stack([ka["x"], ka["y"]], 1)

    def job_extract(i):
        ex = ORBextractor(1000)
        return [(m, k.tobytes(), d.tobytes()) for m, k, d in (ex(frames[(i + j) % 4]) for j in range(6))]

    def job_batch(i):
        ex = ORBextractor(1000, max_batch=4)
        return [(m, k.tobytes(), d.tobytes()) for m, k, d in ex.extract_batch(np.stack(frames))]

    def job_projection(i):
        m = ORBmatcher(0.8)
        out = []
        for _ in range(8):
            nm, assign = m.SearchByProjectionMap(m.frame(fd), occ, pts, desc, 3.0, False, 40.0)
            out.append((nm, assign.tobytes()))
        return out

    def job_knn(i):
        m = ORBmatcher()
        return [tuple(a.tobytes() for a in m.knn2(q, t)) for _ in range(4)]

    def job_undistort(i):
        m = ORBmatcher()
        return [m.UndistortKeyPoints(xy, K, dist).tobytes() for _ in range(20)]

    def job_errors(i):
        m = ORBmatcher()
        seen = []
        for _ in range(20):
            try:
                m.UndistortKeyPoints(xy, K, np.zeros(13))           # always invalid: 13 coefficients
            except _lib.VsgError as e:
                seen.append("vsg_undistort_keypoints" in str(e))
        return seen

    jobs = [job_extract, job_batch, job_projection, job_knn, job_undistort, job_errors, job_extract, job_projection]
    expected = [job(i) for i, job in enumerate(jobs)]               # single-threaded reference run
    assert all(expected[5]) and len(expected[5]) == 20
    for round_ in range(2):
        results = [None] * len(jobs)
        errors = []

        def run(i):
            try:
                results[i] = jobs[i](i)
            except Exception as e:                                  # noqa: BLE001 - reported below
                errors.append((i, repr(e)))

        threads = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        assert not errors, errors
        for i in range(len(jobs)):
            assert results[i] == expected[i], (round_, i, jobs[i].__name__)
