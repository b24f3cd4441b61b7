"""CPU-only: the C++ oracle reproduces the committed golden fixtures (tests/golden/, generated with
real OpenCV for the OpenCV-facing stages — see tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.golden_cases import CASES, frame_of

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INDEX = json.load(open(os.path.join(GOLD, "index.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_golden(oracle, case):
    name, src, wh, nfeat, lap = case
    meta = next(c for c in INDEX["cases"] if c["name"] == name)
    frame = frame_of(src, wh)
    assert sha(frame) == meta["frame_sha"], "synthetic frame generator drifted"
    ex = oracle.OracleExtractor(nfeat)
    mono, kps, desc = ex(frame, lap)
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    assert mono == meta["mono_index"]
    assert len(kps) == meta["n_keypoints"]
    assert kps.tobytes() == gold["keypoints"].tobytes()
    assert np.array_equal(desc, gold["descriptors"])
    for level, lm in enumerate(meta["levels"]):
        assert list(ex.level_size(level)) == lm["size"]
        assert sha(ex.level(level)) == lm["level_sha"]
        assert sha(ex.candidates(level)) == lm["candidates_sha"]
        assert len(ex.level_keypoints(level)) == lm["n_keypoints"]
        b = ex.blurred(level)
        assert (b is None) == ("blurred_sha" not in lm)
        if b is not None:
            assert sha(b) == lm["blurred_sha"]


def test_empty_image_returns_minus_one(oracle):
    ex = oracle.OracleExtractor()
    mono, kps, desc = ex(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(kps) == 0


def test_tables_match_survey_appendix_b(oracle):
    t = oracle.OracleExtractor(1000).tables()
    assert list(t["quota"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(oracle.OracleExtractor(1200).tables()["quota"]) == [261, 217, 181, 151, 126, 105, 87, 72]
    assert list(oracle.OracleExtractor(2000).tables()["quota"]) == [434, 362, 302, 251, 209, 175, 145, 122]


def test_keypoint_invariants(oracle):
    from visual_sgraphs_b200.synth import synth_frame
    ex = oracle.OracleExtractor()
    mono, kps, desc = ex(synth_frame(31))
    assert mono == len(kps)
    assert (kps["class_id"] == -1).all()
    assert ((kps["angle"] >= 0) & (kps["angle"] < 360)).all()
    assert (np.diff(kps["octave"]) >= 0).all()           # level-major output order
    lvl0 = kps[kps["octave"] == 0]
    assert (lvl0["x"] >= 19).all() and (lvl0["x"] < 640 - 19).all()
    assert (kps["size"] == np.floor(31 * ex.tables()["scale"][kps["octave"]])).all()
