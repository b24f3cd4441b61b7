import sys, os
sys.path.insert(0, '.')
import numpy as np
from visual_sgraphs_b200._lib import KEYPOINT_DTYPE, TRACK_POINT_DTYPE
from visual_sgraphs_b200.frame import FrameData
from visual_sgraphs_b200.matcher import ORBmatcher
m = ORBmatcher(0.8)
rng = np.random.default_rng(3)
n_kp, n_map = 1000, 200_000
keys = np.zeros(n_kp, KEYPOINT_DTYPE)
keys["x"], keys["y"] = rng.uniform(20, 620, n_kp), rng.uniform(20, 460, n_kp)
keys["octave"] = rng.integers(0, 8, n_kp)
kdesc = rng.integers(0, 256, (n_kp, 32), dtype=np.uint8)
fdata = FrameData(keys, kdesc)
src = rng.integers(0, n_kp, n_map)
pts = np.zeros(n_map, TRACK_POINT_DTYPE)
pts["proj_x"] = keys["x"][src] + rng.normal(0, 2, n_map)
pts["proj_y"] = keys["y"][src] + rng.normal(0, 2, n_map)
pts["view_cos"] = rng.uniform(0.99, 1.0, n_map)
pts["level"] = np.clip(keys["octave"][src] + rng.integers(0, 2, n_map), 0, 7)
pts["in_view"], pts["blocks"] = True, rng.random(n_map) < 0.9
mp_desc = kdesc[src] ^ np.packbits(rng.random((n_map, 32, 8)) < 0.08, axis=2).reshape(n_map, 32)
occ = np.zeros(n_kp, np.uint8)
frame = m.frame(fdata)
for _ in range(3):
    print("---", file=sys.stderr)
    m.SearchByProjectionMap(frame, occ, pts, mp_desc, 3.0)
