"""C5 (100k x 1M) and C3-shaped (1000 x 200k) brute-force kNN-2 on one GPU: the tensor-core path (VSG_KNN_TC=1/2) against
the POPC kernel (VSG_KNN_TC=0), CUDA events on the matcher stream, results compared."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from visual_sgraphs_b200.matcher import ORBmatcher  # noqa: E402

m = ORBmatcher(device=0)
s = torch.cuda.ExternalStream(m.stream(), device=0)
g = torch.Generator(device="cuda").manual_seed(7)
for nq, nt, reps in ((100_000, 1_000_000, 3), (1000, 200_000, 20), (5000, 50_000, 20)):
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
    t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda", generator=g)
    res = {}
    for mode in ("2", "0"):
        os.environ["VSG_KNN_TC"] = mode
        idx = torch.zeros((nq, 2), dtype=torch.int32, device="cuda")
        dist = torch.zeros((nq, 2), dtype=torch.int32, device="cuda")
        m.knn2_dev(q, t, idx, dist)
        m.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            m.knn2_dev(q, t, idx, dist)
        e1.record(s)
        m.sync()
        ms = e0.elapsed_time(e1) / reps
        res[mode] = (ms, idx.clone(), dist.clone())
        print("%d x %d  VSG_KNN_TC=%s  %.3f ms  %.3e pairs/s" % (nq, nt, mode, ms, nq * nt / ms * 1e3))
    print("   identical:", bool(torch.equal(res["2"][1], res["0"][1]) and torch.equal(res["2"][2], res["0"][2])),
          " speed-up %.2fx" % (res["0"][0] / res["2"][0]))
