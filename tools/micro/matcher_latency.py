"""Per-call latency of the drop-in matcher entry points at tracking-sized inputs (~1000 features per frame), host
arrays in and out, Python/ctypes overhead included."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from oracle import oracle as orc
from tests import match_scenarios as sc
from visual_sgraphs_b200.matcher import ORBmatcher

ka, da, kb, db = sc.two_frames(orc)
fa, fb = sc.frame_data(ka, da, stereo_seed=9), sc.frame_data(kb, db, stereo_seed=3)
m = ORBmatcher(0.9, True)


def timeit(name, fn, reps=200):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    print("%-58s %7.1f us" % (name, (time.perf_counter() - t0) / reps * 1e6))


def upload():
    fr = m.frame(fa)
    fr.close()


timeit("frame upload + destroy (vsg_frame_create)", upload)
fra, frb = m.frame(fa), m.frame(fb)
pts, desc, occ = sc.proj_points(fa, kb, db, (9, 5), 4)
timeit("SearchByProjection(Cur, Last)", lambda: m.SearchByProjectionLast(fra, occ, pts, desc, 15.0, 0))
tp, td, tocc = sc.track_points(fa, kb, db, (9, 5), 21, True)
timeit("SearchByProjection(Frame, MapPoints) 1000 points", lambda: m.SearchByProjectionMap(fra, tocc, tp, td, 3.0))
sp = sc.search_points(kb, (9, 5), 3)
timeit("SearchByProjection(Frame, KeyFrame) relocalisation", lambda: m.SearchByProjectionReloc(fra, occ, sp, db, 10.0, 100))
timeit("SearchByProjection(KeyFrame, Sim3)", lambda: m.SearchByProjectionSim3(fra, occ, sp, db, 8, 1.0))
_, _, inv_sigma2 = sc.sigma_tables()
timeit("Fuse search (pose variant)", lambda: m.FuseSearch(fra, sp, db, 3.0, inv_sigma2))
sp1 = sc.search_points(ka, (-9, -5), 9)
timeit("SearchBySim3", lambda: m.SearchBySim3(fra, frb, sp1, da, sp, db, 7.5))
prev = np.stack([ka["x"], ka["y"]], 1).astype(np.float32).copy()
timeit("SearchForInitialization (window 100)", lambda: m.SearchForInitialization(fa, frb, prev.copy(), 100))
fva, fvb = sc.feature_vector(da, 64), sc.feature_vector(db, 64)
valid = np.ones(fb.n, np.uint8)
timeit("SearchByBoW(KeyFrame, Frame)", lambda: m.SearchByBoW(fb, valid, fa, fvb, fva))
timeit("SearchByBoW(KeyFrame, KeyFrame)", lambda: m.SearchByBoWKF(fa, np.ones(fa.n, np.uint8), fb, valid, fva, fvb))
_, sigma2, _ = sc.sigma_tables()
timeit("SearchForTriangulation", lambda: m.SearchForTriangulation(fa, np.zeros(fa.n, np.uint8), fb, np.zeros(fb.n, np.uint8), fva, fvb,
                                                                  sc.translation_f12((9, 5)), np.array([300., 200.], np.float32), sigma2))
timeit("knn2 1000 x 1000", lambda: m.knn2(da, db))
