"""Does a concurrent host-to-device copy stream slow the extraction kernels down?  Device-resident 512-frame extraction
timed alone, with a background H2D loop, and with a background D2H loop."""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame
B, W, H = 512, 640, 480
base = [synth_frame(1000 + i, W, H) for i in range(16)]
host = torch.from_numpy(np.stack([base[i % 16] for i in range(B)])).pin_memory()
frames_dev = host.cuda()
ex = ORBextractor(1000, max_batch=B)
cap = ex.max_keypoints(W, H)
kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda"); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
n = torch.zeros(B, dtype=torch.int32, device="cuda"); mono = torch.zeros(B, dtype=torch.int32, device="cuda")
def rate(reps=8):
    ex.extract_batch_dev(frames_dev, kps, desc, n, mono); ex.sync()
    t = time.perf_counter()
    for _ in range(reps): ex.extract_batch_dev(frames_dev, kps, desc, n, mono)
    ex.sync(); return (time.perf_counter() - t) / reps * 1e3
print("alone: %.3f ms per 512 frames" % rate())
stop = False
side = torch.cuda.Stream(); scratch = torch.empty_like(frames_dev); back = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
def h2d():
    with torch.cuda.stream(side):
        while not stop:
            scratch.copy_(host, non_blocking=True); side.synchronize()
def d2h():
    with torch.cuda.stream(side):
        while not stop:
            back.copy_(scratch, non_blocking=True); side.synchronize()
for name, fn in (("H2D", h2d), ("D2H", d2h)):
    stop = False; th = threading.Thread(target=fn); th.start(); time.sleep(0.1)
    print("with a background %s loop: %.3f ms per 512 frames" % (name, rate()))
    stop = True; th.join()
