"""Multi-GPU plumbing: one process per GPU under torch.distributed.

* Extraction shards by frame — independent units, no data-path collective: `frame_shard`.
* Brute-force kNN-2 shards the TRAIN set: each rank searches its contiguous shard and reports global train
  indices; the per-rank (nq x 2) top-2 lists are all-gathered and merged by (distance, index) lexicographic order,
  which reproduces cv::BFMatcher::knnMatch's tie rule (lower train index first, Frame.cc:1200) exactly.

The compute callables are injected so the same plumbing runs on NCCL + CUDA (bench.py, GPU tests) and on gloo in
the CPU test of the host-side logic.
"""
import numpy as np


def shard_bounds(n, world):
    """Contiguous, balanced [begin, end) per rank (first n % world ranks get one extra row)."""
    base, extra = divmod(n, world)
    bounds, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        bounds.append((start, start + size))
        start += size
    return bounds


def frame_shard(nframes, rank, world):
    return shard_bounds(nframes, world)[rank]


def knn2_sharded(dist, query, train_shard, shard_begin, local_knn2, merge, make_buffer):
    """query: (nq, 32) on every rank; train_shard: this rank's rows; shard_begin: global index of its first row.
    local_knn2(query, train, offset) -> (idx, dist) tensors (nq, 2) int32; merge(idx_parts, dist_parts) ->
    (idx, dist) from (world, nq, 2) stacks; make_buffer(shape) allocates an int32 tensor on the right device."""
    world = dist.get_world_size()
    idx, d = local_knn2(query, train_shard, shard_begin)
    idx_parts = make_buffer((world,) + tuple(idx.shape))
    d_parts = make_buffer((world,) + tuple(d.shape))
    dist.all_gather_into_tensor(idx_parts, idx.contiguous()) if hasattr(dist, "all_gather_into_tensor") and \
        idx.is_cuda else dist.all_gather(list(idx_parts.unbind(0)), idx.contiguous())
    dist.all_gather_into_tensor(d_parts, d.contiguous()) if hasattr(dist, "all_gather_into_tensor") and \
        d.is_cuda else dist.all_gather(list(d_parts.unbind(0)), d.contiguous())
    return merge(idx_parts, d_parts)


def merge_top2_numpy(idx_parts, dist_parts):
    """Host restatement of vsg_knn2_merge_dev for the CPU test: (parts, nq, 2) -> (nq, 2) by (dist, idx)."""
    idx_parts = np.asarray(idx_parts)
    dist_parts = np.asarray(dist_parts)
    parts, nq, _ = idx_parts.shape
    flat_i = idx_parts.transpose(1, 0, 2).reshape(nq, parts * 2).astype(np.int64)
    flat_d = dist_parts.transpose(1, 0, 2).reshape(nq, parts * 2).astype(np.int64)
    key = np.where(flat_i < 0, np.iinfo(np.int64).max, flat_d * (1 << 32) + flat_i)
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    out_i = np.take_along_axis(flat_i, order, 1).astype(np.int32)
    out_d = np.take_along_axis(flat_d, order, 1).astype(np.int32)
    return out_i, out_d


def _all_gather_rows(dist, rows, counts, device):
    """All-gather variable-length 2-D arrays (same trailing shape/dtype on every rank): pad to the longest, gather,
    trim.  `counts` = row count of every rank (already known to all).  Returns the list of per-rank arrays."""
    import torch
    world = dist.get_world_size()
    longest = max(max(counts), 1)
    rows = np.ascontiguousarray(rows)
    pad = np.zeros((longest,) + rows.shape[1:], rows.dtype)
    pad[: rows.shape[0]] = rows
    mine = torch.from_numpy(pad).to(device)
    out = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=device)
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine)
    else:
        dist.all_gather(list(out.unbind(0)), mine)
    out = out.cpu().numpy()
    return [out[r, : counts[r]] for r in range(world)]


def search_by_projection_map_sharded(dist, n_mp_total, pts_shard, local_candidates, resolve, device="cpu"):
    """SearchByProjection(Frame&, vector<MapPoint*>&) (ORBmatcher.cc:42-144) with the MAP POINTS sharded contiguously
    over ranks (BASELINE config 3) and the frame replicated.

    local_candidates() -> (cand_ptr [n_local + 1], cand_idx, cand_dist) int32 arrays for this rank's shard
        (vsg_projection_map_candidates: GPU window query + Hamming distances, reference candidate order);
    pts_shard: this rank's vsg_track_point records (their `blocks` flag is needed by every rank's replay);
    resolve(pts_all, cand_ptr, cand_idx, cand_dist) -> (nmatches, assign) replays ORBmatcher.cc:76-141 over all lists
        (vsg_projection_map_resolve, host code).
    One exchange step: per-rank list sizes (one int each), then the padded lists and point records.  Every rank ends
    with the identical (nmatches, assign) of the single-process call."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(n_mp_total, world)
    n_local = bounds[rank][1] - bounds[rank][0]
    ptr, idx, d = local_candidates()
    ptr = np.asarray(ptr, np.int32)
    assert len(ptr) == n_local + 1 and len(pts_shard) == n_local
    total = torch.tensor([int(ptr[-1])], dtype=torch.int64, device=device)
    totals = torch.zeros(world, dtype=torch.int64, device=device)
    if total.is_cuda:
        dist.all_gather_into_tensor(totals, total)
    else:
        dist.all_gather(list(totals.unbind(0)), total[0])
    totals = [int(t) for t in totals.cpu()]
    n_locals = [e - b for b, e in bounds]
    lens = np.diff(ptr).astype(np.int32).reshape(-1, 1)
    cand = np.stack([np.asarray(idx, np.int32)[: ptr[-1]], np.asarray(d, np.int32)[: ptr[-1]]], 1)
    all_lens = _all_gather_rows(dist, lens, n_locals, device)
    all_cand = _all_gather_rows(dist, cand, totals, device)
    all_pts = _all_gather_rows(dist, np.ascontiguousarray(pts_shard).view(np.uint8).reshape(n_local, -1), n_locals, device)
    lens = np.concatenate([a.reshape(-1) for a in all_lens]) if n_mp_total else np.zeros(0, np.int32)
    gptr = np.zeros(n_mp_total + 1, np.int32)
    np.cumsum(lens, out=gptr[1:])
    cand = np.concatenate(all_cand) if sum(totals) else np.zeros((0, 2), np.int32)
    pts_all = np.concatenate(all_pts).reshape(-1).view(pts_shard.dtype)
    return resolve(pts_all, gptr, np.ascontiguousarray(cand[:, 0]), np.ascontiguousarray(cand[:, 1]))


def search_by_projection_map_token_ring(dist, n_kp, occupied, shard_begin, local_replay):
    """The exchange of vsg_search_by_projection_map_sharded (csrc/match_methods.cu) restated over torch.distributed, so
    that its logic runs on gloo in the CPU tests: the claim state (`blocked`, n_kp bytes, initially `occupied`) travels
    down the ranks as a token; rank r replays its own shard when the token arrives —
    local_replay(blocked, assign) -> nmatches updates both arrays in place (vsg_projection_map_resolve_shard) — and one
    all-gather of the per-rank assignments (n_kp ints + a count) gives every rank the one-call result: shard order = map
    order, so later shards overwrite earlier ones' slots, and nmatches adds up the assignment events."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    blocked = np.ascontiguousarray(occupied, np.uint8).copy()
    if rank > 0:
        tok = torch.zeros(n_kp, dtype=torch.uint8)
        dist.recv(tok, src=rank - 1)
        blocked = tok.numpy().copy()
    assign = np.full(n_kp, -1, np.int32)
    nm = local_replay(blocked, assign)
    if rank + 1 < world:
        dist.send(torch.from_numpy(blocked), dst=rank + 1)
    rec = torch.from_numpy(np.concatenate([assign, np.array([nm], np.int32)]))
    out = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(out, rec)
    final = np.full(n_kp, -1, np.int32)
    total = 0
    for r in range(world):
        a = out[r].numpy()
        final = np.where(a[:n_kp] >= 0, a[:n_kp], final).astype(np.int32)
        total += int(a[n_kp])
    return total, final
