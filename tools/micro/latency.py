"""Single-frame latency through vsg_extract (host frame in, host results out): median / p95 of 300 calls, per-stage kernel times."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
frame = synth_frame(1, 640, 480)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(30):
    ex(frame)
ts = []
for _ in range(300):
    t = time.perf_counter(); ex(frame); ts.append(time.perf_counter() - t)
ts = np.array(ts) * 1e6
print("latency us: median %.1f p95 %.1f min %.1f" % (np.median(ts), np.percentile(ts, 95), ts.min()))
try:
    ex.profile(True)
    for _ in range(50):
        ex(frame)
    ms, runs = ex.stage_ms()
    print("stages us:", {k: round(v * 1e3 / max(runs, 1), 1) for k, v in ms.items()})
except Exception as e:  # noqa: BLE001
    print("no stage profile:", e)
