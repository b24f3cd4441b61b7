"""Timeline of the tensor-core blur's warp roles on CTA 0 (debug build: tools/build_variant.sh bt_trace blur_tc.cu -DBT_TRACE)."""
import ctypes as C, os, sys
import numpy as np
os.environ["VSG_LIB_PATH"] = os.path.abspath("gpurun_variants/libvsg_bt_trace.so")
os.environ["VSG_BLUR_TC"] = "1"
sys.path.insert(0, ".")
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
from visual_sgraphs_b200 import _lib
frames = np.stack([synth_frame(i % 8, 640, 480) for i in range(256)])
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=256)
import torch
d = torch.from_numpy(frames).cuda()
cap = ex.max_keypoints(640, 480)
kd = torch.zeros((256, cap, 28), dtype=torch.uint8, device="cuda"); dd = torch.zeros((256, cap, 32), dtype=torch.uint8, device="cuda")
nd = torch.zeros(256, dtype=torch.int32, device="cuda"); md = torch.zeros(256, dtype=torch.int32, device="cuda")
for _ in range(2):
    ex.extract_batch_dev(d, kd, dd, nd, md); ex.sync()
L = _lib.load()
out = np.zeros((6, 64, 4), np.int64)
print("rc", L.vsg_debug_bt_trace(out.ctypes.data_as(C.c_void_p)))
t0 = out[2, 20, 0]
names = ["tma:   start | in_empty seen | issued", "patch: start | decoded | in_full seen | a_ready", "gemm1: start | a_ready seen | d1_empty seen | committed",
         "gemm2: start | b2_full seen | d2_empty seen | committed", "epi1:  start | d1_full+b2_empty seen | loaded (d1_empty) | b2_full",
         "epi2:  start | d2_full seen | loaded (d2_empty) | stored"]
for it in range(20, 30):
    print("--- tile", it)
    for role in range(6):
        print("  %-70s" % names[role], [int(x - t0) if x else None for x in out[role, it, :4]])
