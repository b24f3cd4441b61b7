#!/usr/bin/env python3
"""Multi-GPU parity + timing of the two sharded matcher paths over NCCL, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--nq 100000 --nt 1000000]

  C5  brute-force kNN-2 with the TRAIN set sharded: local vsg_knn2_dev -> all-gather (NCCL) -> vsg_knn2_merge_dev;
      checked against the single-GPU result on rank 0, timed with CUDA events (max over ranks).
  C3  SearchByProjection(Frame, MapPoints) with the MAP POINTS sharded: vsg_projection_map_candidates per rank ->
      all-gather of the candidate lists -> vsg_projection_map_resolve; checked against the one-call method.
Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nq", type=int, default=100_000)
    ap.add_argument("--nt", type=int, default=1_000_000)
    ap.add_argument("--n-map", type=int, default=200_000)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from visual_sgraphs_b200 import sharded
    from visual_sgraphs_b200._lib import KEYPOINT_DTYPE, TRACK_POINT_DTYPE
    from visual_sgraphs_b200.frame import FrameData
    from visual_sgraphs_b200.matcher import ORBmatcher

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    m = ORBmatcher(0.8, device=local)
    stream = torch.cuda.ExternalStream(m.stream(), device=local)
    out = {"world": world}

    # ---------------- C5: train-sharded kNN-2 ----------------
    g = torch.Generator(device="cuda").manual_seed(7)          # same seed on every rank: identical q / t
    q = torch.randint(0, 256, (args.nq, 32), dtype=torch.uint8, device=dev, generator=g)
    t = torch.randint(0, 256, (args.nt, 32), dtype=torch.uint8, device=dev, generator=g)
    t[7] = t[args.nt - 3]
    q[0] = t[7]                                                # a tie across shards
    b, e = sharded.shard_bounds(args.nt, world)[rank]
    t_shard = t[b:e].contiguous()

    # The matcher's stream is a non-blocking stream: it is not ordered against torch's default stream (allocation
    # fills, NCCL's completion hand-off), so both sides are synchronised explicitly around every hand-over.
    def local_knn2(qq, tt, off):
        idx = torch.zeros((qq.shape[0], 2), dtype=torch.int32, device=dev)
        d = torch.zeros((qq.shape[0], 2), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        m.knn2_dev(qq, tt, idx, d, off)
        m.sync()
        return idx, d

    def merge(ip, dp):
        oi = torch.zeros(ip.shape[1:], dtype=torch.int32, device=dev)
        od = torch.zeros(ip.shape[1:], dtype=torch.int32, device=dev)
        torch.cuda.synchronize()                               # all-gather done, fills done
        m.knn2_merge_dev(ip, dp, oi, od)
        m.sync()
        return oi, od

    def run_c5():
        return sharded.knn2_sharded(dist, q, t_shard, b, local_knn2, merge,
                                    lambda shape: torch.zeros(shape, dtype=torch.int32, device=dev))

    run_c5()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx, d = run_c5()
    torch.cuda.synchronize(); dist.barrier()
    c5_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    dist.all_reduce(c5_ms, op=dist.ReduceOp.MAX)
    ok5 = True
    if rank == 0:
        fi, fd_ = local_knn2(q, t, 0)
        ok5 = bool(torch.equal(fi, idx) and torch.equal(fd_, d))
    out["c5"] = {"nq": args.nq, "nt": args.nt, "ms": float(c5_ms), "pairs_per_s": args.nq * args.nt / (float(c5_ms) * 1e-3),
                 "matches_rank0_single_gpu": ok5}

    # ---------------- C3: map-point-sharded SearchByProjection ----------------
    rng = np.random.default_rng(3)
    n_kp, n_map = 1000, args.n_map
    keys = np.zeros(n_kp, KEYPOINT_DTYPE)
    keys["x"], keys["y"] = rng.uniform(20, 620, n_kp), rng.uniform(20, 460, n_kp)
    keys["octave"] = rng.integers(0, 8, n_kp)
    keys["angle"] = rng.uniform(0, 360, n_kp)
    desc = rng.integers(0, 256, (n_kp, 32), dtype=np.uint8)
    fdata = FrameData(keys, desc)
    src = rng.integers(0, n_kp, n_map)
    pts = np.zeros(n_map, TRACK_POINT_DTYPE)
    pts["proj_x"] = keys["x"][src] + rng.normal(0, 2, n_map)
    pts["proj_y"] = keys["y"][src] + rng.normal(0, 2, n_map)
    pts["view_cos"] = rng.uniform(0.99, 1.0, n_map)
    pts["depth"] = rng.uniform(1, 40, n_map)
    pts["level"] = np.clip(keys["octave"][src] + rng.integers(0, 2, n_map), 0, 7)
    pts["in_view"], pts["blocks"] = rng.random(n_map) < 0.95, rng.random(n_map) < 0.9
    flips = (rng.random((n_map, 32, 8)) < 0.08)
    mp_desc = desc[src] ^ np.packbits(flips, axis=2).reshape(n_map, 32)
    occ = np.zeros(n_kp, np.uint8)
    fr = m.frame(fdata)
    mb, me = sharded.shard_bounds(n_map, world)[rank]

    def run_c3():
        return sharded.search_by_projection_map_sharded(
            dist, n_map, pts[mb:me], lambda: m.ProjectionMapCandidates(fr, pts[mb:me], mp_desc[mb:me], 3.0),
            lambda pa, cp, ci, cd: m.ProjectionMapResolve(fdata, occ, pa, cp, ci, cd), device=dev)

    run_c3()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    nm, assign = run_c3()
    torch.cuda.synchronize(); dist.barrier()
    c3_ms = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev)
    dist.all_reduce(c3_ms, op=dist.ReduceOp.MAX)
    ok3 = True
    if rank == 0:
        t1 = time.perf_counter()
        wnm, wassign = m.SearchByProjectionMap(fr, occ, pts, mp_desc, 3.0)
        single_ms = (time.perf_counter() - t1) * 1e3
        ok3 = bool(nm == wnm and np.array_equal(assign, wassign))
        out["c3"] = {"n_map": n_map, "n_keypoints": n_kp, "ms": float(c3_ms), "single_gpu_one_call_ms": single_ms,
                     "nmatches": int(nm), "matches_single_gpu": ok3}
        print(json.dumps(out), flush=True)
        if not (ok3 and ok5):
            sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
