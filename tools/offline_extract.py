#!/usr/bin/env python3
"""Offline sequence extraction (BASELINE config 4): a synthetic 1280x720 RGB sequence, 2000 features per frame, sharded by
frame across the GPUs of one box — one process per GPU, no data-path collective.

    python tools/offline_extract.py [--frames 10000] [--batch 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 \
        tools/offline_extract.py --frames 10000

Every rank owns the contiguous frame range sharded.frame_shard gives it and streams it through
vsg_extract_batch_color (cvtColor RGB->gray on the device, Tracking.cc:1595-1608, then ORBextractor::operator()) from a
ring of pinned host batches with two handles on two host threads, exactly like a ROS-free driver would: pinned RGB
frames in, keypoints + descriptors out.  The frames of the ring are synthetic (32 distinct images, shifted copies) so
that a 10k-frame sequence does not need 27 GB of host memory.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W, H, NFEAT = 1280, 720, 2000


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--handles", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from visual_sgraphs_b200 import sharded
    from visual_sgraphs_b200._lib import check, load, ptr
    from visual_sgraphs_b200.extractor import ORBextractor
    from visual_sgraphs_b200.synth import synth_frame

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load()
    begin, end = sharded.frame_shard(args.frames, rank, world)
    B = args.batch

    # ring of pinned RGB batches: gray synthetic frames with per-channel offsets (the luma keeps its structure)
    base = [synth_frame(31000 + 97 * rank + i, W, H) for i in range(8)]
    rng = np.random.default_rng(rank)
    ring = []
    for h in range(args.handles):
        t = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        a = t.numpy()
        for f in range(B):
            g = np.roll(base[(f + h) % 8], (3 * (f // 8), 5 * (f // 8) + h), (0, 1)).astype(np.int16)
            for c in range(3):
                a[f, :, :, c] = np.clip(g + int(rng.integers(-12, 13)), 0, 255)
        ring.append(t)

    handles = [ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local, max_batch=B) for _ in range(args.handles)]
    cap = handles[0].max_keypoints(W, H)
    outs = [(torch.zeros((B, cap, 28), dtype=torch.uint8).pin_memory(), torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory(),
             np.zeros(B, np.int32), np.zeros(B, np.int32)) for _ in range(args.handles)]
    nbatches = (end - begin + B - 1) // B
    totals = [0] * args.handles

    def worker(h):
        kp, de, n, mono = outs[h]
        frames = ring[h].numpy()
        for b in range(h, nbatches, args.handles):
            nf = min(B, end - begin - b * B)
            check(lib.vsg_extract_batch_color(handles[h]._h, ptr(frames), nf, W, H, W * 3, W * H * 3, 3, 1, 0, 0, ptr(kp), ptr(de),
                                              cap, ptr(n), ptr(mono)))
            totals[h] += int(n[:nf].sum())

    for h in range(args.handles):                     # warm-up: shapes, allocations
        check(lib.vsg_extract_batch_color(handles[h]._h, ptr(ring[h].numpy()), B, W, H, W * 3, W * H * 3, 3, 1, 0, 0,
                                          ptr(outs[h][0]), ptr(outs[h][1]), cap, ptr(outs[h][2]), ptr(outs[h][3])))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker, args=(h,)) for h in range(args.handles)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    kps = torch.tensor([float(sum(totals))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        dist.all_reduce(kps, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"workload": "C4: 1280x720 RGB sequence, 2000 features, cvtColor + extraction, frames sharded by rank",
                          "frames": args.frames, "n_gpus": world, "seconds": float(secs), "frames_per_s": args.frames / float(secs),
                          "keypoints_per_frame": float(kps) / args.frames, "h2d_bytes_per_frame": W * H * 3,
                          "d2h_bytes_per_frame": cap * 60, "batch": B, "handles_per_gpu": args.handles}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
