#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ (run in the build container, where cv2 exists).

For each case the OpenCV-facing stages come from REAL OpenCV via oracle/cv2_ref.py (pyramid levels,
per-cell FAST candidates, blurred levels) and the remaining stages (oct-tree selection, IC_Angle,
rBRIEF, output ordering) from the C++ oracle, after asserting that the oracle agrees with cv2 on the
shared stages.  Large planes are stored as SHA-256 digests, small results verbatim (npz).

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cv2  # noqa: E402
from oracle import cv2_ref, oracle as orc  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


from tests.golden_cases import CASES, frame_of  # noqa: E402


def main():
    index = {"opencv": cv2.__version__, "cases": []}
    for name, src, wh, nfeat, lap in CASES:
        frame = frame_of(src, wh)
        ex = orc.OracleExtractor(nfeat)
        mono, kps, desc = ex(frame, lap)
        pyr = cv2_ref.pyramid(frame)
        levels = []
        for level in range(8):
            pad = pyr[level]
            assert np.array_equal(ex.level_padded(level), pad), (name, level)
            view = pad[19:-19, 19:-19]
            cands, retries = cv2_ref.fast_candidates(view)
            assert np.array_equal(ex.candidates(level), cands), (name, level)
            lk = ex.level_keypoints(level)
            entry = {
                "size": [int(view.shape[1]), int(view.shape[0])],
                "level_sha": sha(view),
                "n_candidates": int(len(cands)),
                "candidates_sha": sha(cands),
                "min_th_retries": int(retries),
                "n_keypoints": int(len(lk)),
            }
            b = ex.blurred(level)
            if b is not None:
                assert np.array_equal(b, cv2_ref.blur(view)), (name, level)
                entry["blurred_sha"] = sha(b)
            levels.append(entry)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), keypoints=kps, descriptors=desc)
        index["cases"].append({"name": name, "source": list(src), "size": list(wh), "nfeatures": nfeat,
                               "lapping": list(lap), "mono_index": int(mono), "n_keypoints": int(len(kps)),
                               "frame_sha": sha(frame), "levels": levels})
        print(name, "n=%d mono=%d cands=%s" % (len(kps), mono, [l["n_candidates"] for l in levels]))
    with open(os.path.join(HERE, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
