import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
frames = np.stack([synth_frame(100 + i, 640, 480) for i in range(n)])
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=n)
out = ex.extract_batch(frames)
os.environ["VSG_FAST_TMA"] = "0"
ex2 = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=n)
ref = ex2.extract_batch(frames)
ok = all(a[0] == b[0] and a[1].tobytes() == b[1].tobytes() and np.array_equal(a[2], b[2]) for a, b in zip(out, ref))
print("frames", n, "identical to the non-TMA kernel:", ok, "keypoints", [len(o[1]) for o in out[:4]])
