import os, sys
import numpy as np
os.environ["VSG_LIB_PATH"] = os.path.abspath("gpurun_variants/libvsg_ot.so")
sys.path.insert(0, ".")
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
frame = synth_frame(1, 640, 480)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(2):
    ex(frame)
