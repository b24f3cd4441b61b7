"""One brute-force kNN-2 call at the BASELINE config-5 shape (100k x 1M) — for `ncu -k regex:knn2`."""
import sys
sys.path.insert(0, '.')
import torch
from visual_sgraphs_b200.matcher import ORBmatcher
nq, nt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100_000, 1_000_000)
m = ORBmatcher()
g = torch.Generator(device="cuda").manual_seed(7)
q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device="cuda", generator=g)
t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, device="cuda", generator=g)
idx = torch.zeros((nq, 2), dtype=torch.int32, device="cuda")
dist = torch.zeros((nq, 2), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
m.knn2_dev(q, t, idx, dist)
m.sync()
print("ok", int(dist[:, 0].min()), int(dist[:, 0].max()))
