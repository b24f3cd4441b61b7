"""GPU parity test for Frame::ComputeStereoMatches (BASELINE config 2: 752x480 stereo, 1200 features per side):
extraction of both images on the device + vsg_stereo_match on the device pyramids against the oracle."""
import numpy as np
import pytest

from visual_sgraphs_b200.synth import synth_stereo_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,size,nfeat", [(42, (752, 480), 1200), (7, (376, 240), 500)])
def test_stereo_matches_parity(oracle, seed, size, nfeat):
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    left, right = synth_stereo_pair(seed, *size)
    exl, exr = ORBextractor(nfeat), ORBextractor(nfeat)          # two instances, like Frame.cc:129-132
    _, kl, dl = exl(left)
    _, kr, dr = exr(right)
    oxl, oxr = oracle.OracleExtractor(nfeat), oracle.OracleExtractor(nfeat)
    _, okl, odl = oxl(left)
    _, okr, odr = oxr(right)
    assert kl.tobytes() == okl.tobytes() and np.array_equal(dl, odl) and kr.tobytes() == okr.tobytes()
    mb, mbf = 0.11, 47.9
    u, d = ORBmatcher().ComputeStereoMatches(exl, exr, kl, dl, kr, dr, mb, mbf)
    wu, wd = oracle.stereo_matches(oxl, oxr, okl, odl, okr, odr, mb, mbf)
    assert np.array_equal(u, wu)
    assert np.array_equal(d, wd)
    assert (u >= 0).sum() > 0.2 * len(kl)


def test_stereo_inside_a_batch(oracle):
    """Left and right images extracted as frames 0 and 1 of one batched call (pair i on one GPU)."""
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    left, right = synth_stereo_pair(3, 752, 480)
    ex = ORBextractor(1200, max_batch=2)
    (_, kl, dl), (_, kr, dr) = ex.extract_batch(np.stack([left, right]))
    u, d = ORBmatcher().ComputeStereoMatches(ex, ex, kl, dl, kr, dr, 0.11, 47.9, frame_l=0, frame_r=1)
    oxl, oxr = oracle.OracleExtractor(1200), oracle.OracleExtractor(1200)
    oxl(left)
    oxr(right)
    wu, wd = oracle.stereo_matches(oxl, oxr, kl, dl, kr, dr, 0.11, 47.9)
    assert np.array_equal(u, wu) and np.array_equal(d, wd)


def test_stereo_batch_matches_per_pair_and_oracle(oracle):
    """vsg_stereo_match_batch (pairs 2p / 2p+1 of one extractor batch, device-resident inputs, median rejection on the
    device) against the per-pair entry point and the oracle."""
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    pairs = [synth_stereo_pair(20 + i, 752, 480) for i in range(3)]
    pairs.append((pairs[0][0], np.full_like(pairs[0][1], 128)))        # a right image without keypoints: no matches at all
    stack = np.stack([im for pr in pairs for im in pr])
    ex = ORBextractor(1200, max_batch=len(stack))
    res = ex.extract_batch(stack)
    m = ORBmatcher()
    ub, db = m.ComputeStereoMatchesBatch(ex, len(pairs), 0.11, 47.9)
    for p, (left, right) in enumerate(pairs):
        (_, kl, dl), (_, kr, dr) = res[2 * p], res[2 * p + 1]
        u1, d1 = m.ComputeStereoMatches(ex, ex, kl, dl, kr, dr, 0.11, 47.9, frame_l=2 * p, frame_r=2 * p + 1)
        assert np.array_equal(ub[p, :len(kl)], u1) and np.array_equal(db[p, :len(kl)], d1), p
        assert (ub[p, len(kl):] == -1).all() and (db[p, len(kl):] == -1).all()
        oxl, oxr = oracle.OracleExtractor(1200), oracle.OracleExtractor(1200)
        oxl(left)
        oxr(right)
        wu, wd = oracle.stereo_matches(oxl, oxr, kl, dl, kr, dr, 0.11, 47.9)
        assert np.array_equal(u1, wu) and np.array_equal(d1, wd), p
    assert (ub[0] >= 0).sum() > 200 and (ub[3] >= 0).sum() == 0


def test_keypoints_at_the_border_get_no_match_instead_of_an_out_of_bounds_read(oracle):
    """Caller-supplied keypoints closer than the SAD window to the image border (the extractor never produces them; the
    reference would throw a cv::Exception from rowRange / colRange, Frame.cc:1054,1069): the kernel must not read outside
    the level planes — such keypoints simply stay unmatched, all others keep their oracle result."""
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.matcher import ORBmatcher
    from visual_sgraphs_b200.synth import synth_stereo_pair
    left, right = synth_stereo_pair(8200)
    exl, exr = ORBextractor(1200, 1.2, 8, 20, 7), ORBextractor(1200, 1.2, 8, 20, 7)
    _, kl, dl = exl(left)
    _, kr, dr = exr(right)
    m = ORBmatcher()
    u0, d0 = m.ComputeStereoMatches(exl, exr, kl, dl, kr, dr, 0.11, 47.9)
    kl2, kr2 = kl.copy(), kr.copy()
    edge = np.arange(0, len(kl2), 7)
    kl2["x"][edge] = np.tile([1.0, 3.5, 750.0, 400.0], len(edge))[: len(edge)]      # left patch off the left / right edge
    kl2["y"][edge] = np.tile([200.0, 1.0, 240.0, 478.5], len(edge))[: len(edge)]    # ... or off the top / bottom
    u1, d1 = m.ComputeStereoMatches(exl, exr, kl2, dl, kr2, dr, 0.11, 47.9)
    assert np.isfinite(u1).all() and ((u1 == -1) | (u1 >= 0)).all()
    assert (u1[edge][(kl2["y"][edge] < 5) | (kl2["y"][edge] > 474)] == -1).all()
