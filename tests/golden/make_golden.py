#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ (run in the build container, where cv2 exists).

For each case the OpenCV-facing stages come from REAL OpenCV via oracle/cv2_ref.py (pyramid levels,
per-cell FAST candidates, blurred levels) and the final keypoints / descriptors / per-level keypoints from the
REFERENCE ITSELF — oracle/_ref/libvsg_ref.so = /root/reference/orb_slam3/src/ORBextractor.cc compiled unmodified
(oracle/ref_build/Makefile) — after asserting that the oracle port agrees with both.  Large planes are stored as
SHA-256 digests, small results verbatim (npz).

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
from oracle import cv2_ref, oracle as orc, ref  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


from tests.golden_cases import CASES, frame_of  # noqa: E402


def main():
    index = {"opencv": cv2.__version__, "final_outputs_from": "oracle/_ref (reference ORBextractor.cc, unmodified)", "cases": []}
    for name, src, wh, nfeat, lap in CASES:
        frame = frame_of(src, wh)
        ex = orc.OracleExtractor(nfeat)
        mono, kps, desc = ex(frame, lap)
        rx = ref.RefExtractor(nfeat)
        rmono, rkps, rdesc = rx(frame, lap)
        assert rmono == mono and rkps.tobytes() == kps.tobytes() and np.array_equal(rdesc, desc), name
        mono, kps, desc = rmono, rkps, rdesc              # what is stored is the reference's output
        pyr = cv2_ref.pyramid(frame)
        levels = []
        for level in range(8):
            pad = pyr[level]
            assert np.array_equal(ex.level_padded(level), pad), (name, level)
            assert np.array_equal(rx.level_padded(level), pad), (name, level)
            view = pad[19:-19, 19:-19]
            cands, retries = cv2_ref.fast_candidates(view)
            assert np.array_equal(ex.candidates(level), cands), (name, level)
            lk = rx.level_keypoints(frame, level)
            assert lk.tobytes() == ex.level_keypoints(level).tobytes(), (name, level)
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
