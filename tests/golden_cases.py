"""The golden-fixture cases shared by tests/golden/make_golden.py and the parity tests.

Adversarial frames follow SURVEY.md §8c: constant image (no keypoints -> descriptors released),
low contrast (minThFAST retry in most cells), checkerboard (massive score ties -> first-max-wins and
sort ties), single bright dot, all corners in one quadrant (oct-tree stops early: size unchanged).
"""
import numpy as np

from visual_sgraphs_b200.synth import synth_frame


def special_frame(kind, w, h):
    if kind == "constant":
        return np.full((h, w), 128, np.uint8)
    if kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        return (((xx // 8 + yy // 8) % 2) * 200 + 20).astype(np.uint8)
    if kind == "dot":
        img = np.full((h, w), 30, np.uint8)
        img[h // 2 - 1:h // 2 + 2, w // 2 - 1:w // 2 + 2] = 250
        return img
    if kind == "quadrant":
        img = np.full((h, w), 90, np.uint8)
        q = synth_frame(99, w // 2, h // 2)
        img[: h // 2, : w // 2] = q
        return img
    if kind == "lowcontrast":
        return (synth_frame(5, w, h).astype(np.float32) * 0.12 + 100).astype(np.uint8)
    raise ValueError(kind)


CASES = [
    # name, frame source, (w, h), nfeatures, lapping
    ("c1_seed1000", ("synth", 1000), (640, 480), 1000, (0, 0)),
    ("c1_seed1001_mono", ("synth", 1001), (640, 480), 1000, (0, 1000)),
    ("c2_seed2000", ("synth", 2000), (752, 480), 1200, (0, 0)),
    ("small_seed7_lap", ("synth", 7), (322, 243), 500, (100, 200)),
    ("constant", ("special", "constant"), (320, 240), 500, (0, 0)),
    ("checker", ("special", "checker"), (320, 240), 500, (0, 0)),
    ("dot", ("special", "dot"), (320, 240), 500, (0, 0)),
    ("quadrant", ("special", "quadrant"), (640, 480), 1000, (0, 0)),
    ("lowcontrast", ("special", "lowcontrast"), (320, 240), 500, (0, 0)),
]


def frame_of(src, wh):
    if src[0] == "synth":
        return synth_frame(src[1], *wh)
    return special_frame(src[1], *wh)


