"""GPU parity tests: the Search* methods behind the C ABI (GPU window query + Hamming, host replay of the
order-dependent bookkeeping) against the CPU oracle's restatement of the reference loops.  Bit-exact: match
indices, counts and updated vbPrevMatched must be identical."""
import numpy as np
import pytest

from tests import match_scenarios as sc

pytestmark = pytest.mark.gpu


def _matcher(nnratio=0.6, check_ori=True):
    from visual_sgraphs_b200.matcher import ORBmatcher
    return ORBmatcher(nnratio, check_ori)


def test_area_search_matches_get_features_in_area(oracle):
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da)
    m = _matcher()
    fr = m.frame(fd)
    rng = np.random.default_rng(8)
    nq = 400
    qx, qy = rng.uniform(-30, 670, nq).astype(np.float32), rng.uniform(-30, 510, nq).astype(np.float32)
    qr = rng.uniform(1, 70, nq).astype(np.float32)
    choices = np.array([(-1, -1), (0, 0), (2, 3), (1, -1), (0, 4)], np.int32)
    pick = rng.integers(0, 5, nq)
    lo, hi = choices[pick, 0].copy(), choices[pick, 1].copy()
    qdesc = db[rng.integers(0, len(db), nq)]
    ptr, idx, dist = m.area_search(fr, qx, qy, qr, lo, hi, qdesc)
    total = 0
    for q in range(nq):
        want = oracle.get_features_in_area(fd.view, float(qx[q]), float(qy[q]), float(qr[q]), int(lo[q]), int(hi[q]))
        got = idx[ptr[q]:ptr[q + 1]]
        assert np.array_equal(got, want), q
        for k, j in enumerate(got):
            assert dist[ptr[q] + k] == oracle.descriptor_distance(qdesc[q], da[j])
        total += len(want)
    assert total > 2000


@pytest.mark.parametrize("stereo", [False, True])
@pytest.mark.parametrize("th,far", [(3.0, False), (1.0, True), (5.0, False)])
def test_search_by_projection_map(oracle, stereo, th, far):
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da, stereo_seed=5 if stereo else None)
    pts, desc, occ = sc.track_points(fd, kb, db, (9, 5), 21, stereo)
    m = _matcher(0.8)
    nm, assign = m.SearchByProjectionMap(m.frame(fd), occ, pts, desc, th, far, 40.0)
    wnm, wassign = oracle.search_by_projection_map(fd.view, occ, pts, desc, th, far, 40.0, float(np.float32(0.8)))
    assert nm == wnm and np.array_equal(assign, wassign)
    assert nm > 100


@pytest.mark.parametrize("th,far", [(3.0, False), (1.0, True), (5.0, False)])
def test_search_by_projection_map_two_cameras(oracle, th, far):
    """F.Nleft != -1 (stereo-fisheye rigs): left- and right-camera searches incl. the cross assignments through
    mvLeftToRightMatch / mvRightToLeftMatch (ORBmatcher.cc:42-216)."""
    sc2 = sc.two_camera_scene(oracle)
    m = _matcher(0.8)
    nm, assign = m.SearchByProjectionMap2Cam(m.frame(sc2["fl"]), m.frame(sc2["fr"]), sc2["occupied"], sc2["l2r"], sc2["r2l"],
                                             sc2["pl"], sc2["pr"], sc2["desc"], th, far, 40.0)
    wnm, wassign = oracle.search_by_projection_map_2cam(sc2["fl"].view, sc2["fr"].view, sc2["occupied"], sc2["l2r"], sc2["r2l"],
                                                        sc2["pl"], sc2["pr"], sc2["desc"], th, far, 40.0, float(np.float32(0.8)))
    assert nm == wnm and np.array_equal(assign, wassign)
    nl = sc2["fl"].n
    assert (assign[:nl] >= 0).sum() > 100 and (assign[nl:] >= 0).sum() > 100


def test_search_by_projection_map_large_map(oracle):
    """Many more map points than keypoints (the C3 shape, scaled): long claim chains on every keypoint."""
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da)
    rng = np.random.default_rng(2)
    reps = 30                                   # > 20 000 queries: the device-ordered lists + entry-walk replay
    kb2 = np.concatenate([kb] * reps)
    db2 = np.concatenate([db] * reps)
    flip = rng.integers(0, 256, db2.shape, dtype=np.uint8) & rng.integers(0, 256, db2.shape, dtype=np.uint8) & \
        rng.integers(0, 256, db2.shape, dtype=np.uint8)
    db2 = db2 ^ flip
    pts, desc, occ = sc.track_points(fd, kb2, db2, (9, 5), 33)
    m = _matcher(0.8)
    nm, assign = m.SearchByProjectionMap(m.frame(fd), occ, pts, desc, 3.0)
    wnm, wassign = oracle.search_by_projection_map(fd.view, occ, pts, desc, 3.0, False, 50.0, float(np.float32(0.8)))
    assert nm == wnm and np.array_equal(assign, wassign)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_projection_last(oracle, mode, check_ori):
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da, stereo_seed=9)
    pts, desc, occ = sc.proj_points(fd, kb, db, (9, 5), 4)
    m = _matcher(0.9, check_ori)
    nm, assign = m.SearchByProjectionLast(m.frame(fd), occ, pts, desc, 15.0, mode)
    wnm, wassign = oracle.search_by_projection_last(fd.view, occ, pts, desc, 15.0, mode, check_ori)
    assert nm == wnm and np.array_equal(assign, wassign)
    assert nm > 50


@pytest.mark.parametrize("check_ori", [True, False])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_search_by_projection_last_two_cameras(oracle, mode, check_ori):
    """CurrentFrame.Nleft != -1: the right-camera search of ORBmatcher.cc:1785-1852, skipped when the left window is empty."""
    sc2 = sc.two_camera_last_scene(oracle)
    m = _matcher(0.9, check_ori)
    nm, assign = m.SearchByProjectionLast2Cam(m.frame(sc2["fl"]), m.frame(sc2["fr"]), sc2["occupied"], sc2["pl"], sc2["pr"],
                                              sc2["desc"], 15.0, mode)
    wnm, wassign = oracle.search_by_projection_last_2cam(sc2["fl"].view, sc2["fr"].view, sc2["occupied"], sc2["pl"], sc2["pr"],
                                                         sc2["desc"], 15.0, mode, check_ori)
    assert nm == wnm and np.array_equal(assign, wassign)
    nl = sc2["fl"].n
    assert (assign[:nl] >= 0).sum() > 100 and (assign[nl:] >= 0).sum() > 100


@pytest.mark.parametrize("window", [10, 100])
def test_search_for_initialization(oracle, window):
    ka, da, kb, db = sc.two_frames(oracle, nfeat=2000)
    f1, f2 = sc.frame_data(ka, da), sc.frame_data(kb, db)
    m = _matcher(0.9, True)
    prev_g = np.stack([ka["x"], ka["y"]], 1).astype(np.float32).copy()
    prev_w = prev_g.copy()
    nm, m12 = m.SearchForInitialization(f1, m.frame(f2), prev_g, window)
    wnm, wm12 = oracle.search_for_initialization(f1.view, f2.view, prev_w, window, float(np.float32(0.9)), True)
    assert nm == wnm and np.array_equal(m12, wm12) and np.array_equal(prev_g, prev_w)
    assert nm > 30


@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_bow(oracle, check_ori):
    ka, da, kb, db = sc.two_frames(oracle, shift=(3, 2))
    kf, f = sc.frame_data(kb, db), sc.frame_data(ka, da)
    rng = np.random.default_rng(6)
    valid = (rng.random(kf.n) < 0.85).astype(np.uint8)
    # a vocabulary-like bucketing: related descriptors mostly share a node
    kfv, ffv = sc.feature_vector(db, 24), sc.feature_vector(da, 24)
    m = _matcher(0.7, check_ori)
    nm, mf = m.SearchByBoW(kf, valid, f, kfv, ffv)
    wnm, wmf = oracle.search_by_bow(kf.view, valid, f.view, kfv, ffv, float(np.float32(0.7)), check_ori)
    assert nm == wnm and np.array_equal(mf, wmf)
    assert nm > 20


@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_bow_two_cameras(oracle, check_ori):
    """F.Nleft != -1: best / second-best per camera, the right-camera match nested in the left one's threshold block with
    its ratio test disabled (ORBmatcher.cc:298-322, :362-390)."""
    ka, da, kb, db = sc.two_frames(oracle, shift=(3, 2))
    keys, desc = np.concatenate([ka, kb]), np.concatenate([da, db])
    kf, f = sc.frame_data(kb, db), sc.frame_data(keys, desc)
    rng = np.random.default_rng(6)
    valid = (rng.random(kf.n) < 0.85).astype(np.uint8)
    kfv, ffv = sc.feature_vector(db, 24), sc.feature_vector(desc, 24)
    m = _matcher(0.7, check_ori)
    nm, mf = m.SearchByBoW(kf, valid, f, kfv, ffv, f_nleft=len(ka))
    wnm, wmf = oracle.search_by_bow(kf.view, valid, f.view, kfv, ffv, float(np.float32(0.7)), check_ori, len(ka))
    assert nm == wnm and np.array_equal(mf, wmf)
    assert (mf[:len(ka)] >= 0).sum() > 20 and (mf[len(ka):] >= 0).sum() > 20


def test_empty_inputs(oracle):
    from visual_sgraphs_b200._lib import KEYPOINT_DTYPE, TRACK_POINT_DTYPE
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da)
    m = _matcher()
    fr = m.frame(fd)
    nm, assign = m.SearchByProjectionMap(fr, np.zeros(fd.n, np.uint8), np.zeros(0, TRACK_POINT_DTYPE), np.zeros((0, 32), np.uint8))
    assert nm == 0 and (assign == -1).all()
    empty = sc.frame_data(np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8))
    efr = m.frame(empty)
    pts, desc, occ = sc.track_points(fd, kb, db, (9, 5), 1)
    nm, assign = m.SearchByProjectionMap(efr, np.zeros(0, np.uint8), pts, desc)
    assert nm == 0 and len(assign) == 0


def test_projection_map_in_two_halves(oracle):
    """vsg_projection_map_candidates (GPU) + vsg_projection_map_resolve (host) == the one-call method == the oracle;
    with the map points split into shards the concatenated lists give the same answer (the multi-GPU C3 path)."""
    from visual_sgraphs_b200 import sharded
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da, stereo_seed=5)
    kb2, db2 = np.concatenate([kb] * 6), np.concatenate([db] * 6)
    pts, desc, occ = sc.track_points(fd, kb2, db2, (9, 5), 21, True)
    m = _matcher(0.8)
    fr = m.frame(fd)
    wnm, wassign = oracle.search_by_projection_map(fd.view, occ, pts, desc, 3.0, False, 50.0, float(np.float32(0.8)))
    cp, ci, cd = m.ProjectionMapCandidates(fr, pts, desc, 3.0)
    nm, assign = m.ProjectionMapResolve(fd, occ, pts, cp, ci, cd)
    assert nm == wnm and np.array_equal(assign, wassign)
    ptrs, idxs, dists = [np.zeros(1, np.int32)], [], []
    for b, e in sharded.shard_bounds(len(pts), 3):
        p, i, d = m.ProjectionMapCandidates(fr, pts[b:e], desc[b:e], 3.0)
        ptrs.append(p[1:] + ptrs[-1][-1])
        idxs.append(i)
        dists.append(d)
    nm3, assign3 = m.ProjectionMapResolve(fd, occ, pts, np.concatenate(ptrs), np.concatenate(idxs), np.concatenate(dists))
    assert nm3 == wnm and np.array_equal(assign3, wassign)
