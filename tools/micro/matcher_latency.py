"""Per-call latency of the drop-in matcher path: frame upload (vsg_frame_create) + SearchByProjection(Cur, Last)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as orc
from tests import match_scenarios as sc
from visual_sgraphs_b200.matcher import ORBmatcher
ka, da, kb, db = sc.two_frames(orc)
fd = sc.frame_data(ka, da, stereo_seed=9)
pts, desc, occ = sc.proj_points(fd, kb, db, (9, 5), 4)
m = ORBmatcher(0.9, True)
for _ in range(20):
    fr = m.frame(fd); fr.close()
t0 = time.perf_counter()
for _ in range(200):
    fr = m.frame(fd); fr.close()
print("frame upload + destroy: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
fr = m.frame(fd)
for _ in range(20):
    m.SearchByProjectionLast(fr, occ, pts, desc, 15.0, 0)
t0 = time.perf_counter()
for _ in range(200):
    m.SearchByProjectionLast(fr, occ, pts, desc, 15.0, 0)
print("SearchByProjection(Cur, Last) on an uploaded frame: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
