import ctypes as C, os, sys
import numpy as np
os.environ["VSG_LIB_PATH"] = os.path.abspath("gpurun_variants/libvsg_pt_trace.so")
sys.path.insert(0, ".")
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
from visual_sgraphs_b200 import _lib
frame = synth_frame(1, 640, 480)
ex = ORBextractor(1000, 1.2, 8, 20, 7)
for _ in range(5):
    ex(frame)
out = np.zeros((4, 20), np.int64)
print("rc", _lib.load().vsg_debug_pt_trace(out.ctypes.data_as(C.c_void_p)))
for t in range(4):
    print("tile", 48 * t, [int(x - out[t, 0]) for x in out[t, :10]])
