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
