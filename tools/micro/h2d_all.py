#!/usr/bin/env python3
"""Aggregate host->device copy ceiling of the box: every rank streams the bench's per-step input (512 frames of 640x480 =
157 MB, pinned) to its GPU with plain cudaMemcpyAsync, all ranks at once — no kernels, no library of ours.  This is the
denominator for bench.py's end-to-end scaling: e2e frames/s cannot exceed  aggregate_h2d_GBs / 307200 B per frame.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/micro/h2d_all.py [--wc] [--d2h]
Prints one JSON line on rank 0 (and writes gpurun_out/h2d_ceiling_n<N>[_wc][_d2h].json).
  --wc   stage the frames in write-combined pinned memory (cudaHostAllocWriteCombined)
  --d2h  run the bench's device->host result copy (31.7 MB per step) concurrently on a second stream
"""
import argparse
import ctypes
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wc", action="store_true")
    ap.add_argument("--d2h", action="store_true")
    ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes, out_bytes = 512 * 640 * 480, 31_707_136
    rt = ctypes.CDLL("libcudart.so.12")
    if args.wc:
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04)) == 0   # cudaHostAllocWriteCombined
        ctypes.memset(p, 7, nbytes)
        host_ptr = p.value
        host = None
    else:
        host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        host.fill_(7)
        host_ptr = host.data_ptr()
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res_d = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
    res_h = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    def step():
        rt.cudaMemcpyAsync(dev.data_ptr(), host_ptr, nbytes, 1, s1.cuda_stream)
        if args.d2h:
            rt.cudaMemcpyAsync(res_h.data_ptr(), res_d.data_ptr(), out_bytes, 2, s2.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        secs = float(dt[0])
        h2d = world * nbytes * args.reps / secs / 1e9
        line = {"ranks": world, "write_combined": args.wc, "with_d2h": args.d2h, "h2d_GBs_aggregate": h2d,
                "h2d_GBs_per_gpu": h2d / world, "d2h_GBs_aggregate": world * out_bytes * args.reps / secs / 1e9 if args.d2h else 0.0,
                "frames_per_s_ceiling": h2d * 1e9 / (640 * 480), "bytes_per_step": nbytes, "reps": args.reps,
                "host_cpus": len(os.sched_getaffinity(0))}
        print(json.dumps(line), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        tag = "n%d%s%s" % (world, "_wc" if args.wc else "", "_d2h" if args.d2h else "")
        json.dump(line, open("gpurun_out/h2d_ceiling_%s.json" % tag, "w"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
