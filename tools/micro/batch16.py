import sys
import numpy as np
sys.path.insert(0, ".")
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
frames = np.stack([synth_frame(i, 640, 480) for i in range(16)])
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=16)
for _ in range(3):
    ex.extract_batch(frames)
