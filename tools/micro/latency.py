import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
f = synth_frame(1000, 640, 480)
ex = ORBextractor(1000, max_batch=1)
for _ in range(10): ex(f)
ex.profile(True)
for _ in range(100): ex(f)
ms, runs = ex.stage_ms()
print({k: round(v / runs * 1e3, 1) for k, v in ms.items()}, "us per stage, runs", runs)
ex.profile(False)
t0 = time.perf_counter()
for _ in range(200): ex(f)
print("per call us", (time.perf_counter() - t0) / 200 * 1e6)
