"""Where does the end-to-end rate stop?  (a) the host-to-device copies of a step alone, in the bench's chunking, (b) the
same while a long kernel keeps the SMs and HBM busy, (c) the full vsg_extract_batch path with and without the result
copies.  python tools/micro/e2e_probe.py"""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from visual_sgraphs_b200._lib import check, load, ptr
from visual_sgraphs_b200.extractor import ORBextractor
from visual_sgraphs_b200.synth import synth_frame

lib = load()
B, W, H = 512, 640, 480
base = [synth_frame(1000 + i, W, H) for i in range(16)]
host = torch.from_numpy(np.stack([base[i % 16] for i in range(B)])).pin_memory()
dev = torch.empty((B, H, W), dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(4)]

def h2d_only(reps=5, chunk=64):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        for c in range(0, B, chunk):
            with torch.cuda.stream(streams[(c // chunk) % 4]):
                dev[c:c + chunk].copy_(host[c:c + chunk], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps

dt = h2d_only(); print("H2D alone, 64-frame chunks on 4 streams: %.2f ms per step -> %.1f GB/s" % (dt * 1e3, host.numel() / dt / 1e9))
# busy kernel on another stream: a device-resident extraction loop
ex = ORBextractor(1000, max_batch=B)
cap = ex.max_keypoints(W, H)
kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda"); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
n = torch.zeros(B, dtype=torch.int32, device="cuda"); mono = torch.zeros(B, dtype=torch.int32, device="cuda")
frames_dev = host.cuda()
stop = False
def busy():
    while not stop:
        ex.extract_batch_dev(frames_dev, kps, desc, n, mono); ex.sync()
th = threading.Thread(target=busy); th.start(); time.sleep(0.2)
dt = h2d_only(); print("H2D while the extraction kernels run: %.2f ms per step -> %.1f GB/s" % (dt * 1e3, host.numel() / dt / 1e9))
stop = True; th.join(); ex.close()

def e2e(nh, with_results, steps=10):
    cuts = [B * i // nh for i in range(nh + 1)]
    hs = [ORBextractor(1000, max_batch=cuts[i + 1] - cuts[i]) for i in range(nh)]
    outs = [(torch.zeros((cuts[i + 1] - cuts[i], cap, 28), dtype=torch.uint8).pin_memory(), torch.zeros((cuts[i + 1] - cuts[i], cap, 32), dtype=torch.uint8).pin_memory(),
             np.zeros(cuts[i + 1] - cuts[i], np.int32), np.zeros(cuts[i + 1] - cuts[i], np.int32)) for i in range(nh)]
    hn = host.numpy()
    stamps = [[] for _ in range(nh)]
    def work(i, k):
        b, e = cuts[i], cuts[i + 1]
        for _ in range(k):
            stamps[i].append(time.perf_counter())
            check(lib.vsg_extract_batch(hs[i]._h, ptr(hn[b:e]), e - b, W, H, W, W * H, 0, 0, ptr(outs[i][0]) if with_results else None,
                                        ptr(outs[i][1]) if with_results else None, cap, ptr(outs[i][2]), ptr(outs[i][3])))
    def run(k):
        ths = [threading.Thread(target=work, args=(i, k)) for i in range(nh)]
        [t.start() for t in ths]; [t.join() for t in ths]
    run(2); torch.cuda.synchronize(); t = time.perf_counter(); run(steps); torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / steps
    for h in hs: h.close()
    if os.environ.get("PROBE_STAMPS"):
        t0 = min(st[-4] for st in stamps)
        for i, st in enumerate(stamps):
            print("  handle %d call starts (ms): %s" % (i, " ".join("%.2f" % ((x - t0) * 1e3) for x in st[-4:])))
    return dt
for nh in (4, 8):
    for wr in (True,):
        dt = e2e(nh, wr); print("e2e %d handles, results %s: %.2f ms per step -> %.1f k frames/s" % (nh, "copied back" if wr else "left on the device", dt * 1e3, B / dt / 1e3))
