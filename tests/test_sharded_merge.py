"""CPU-only, world_size 2 over gloo: the N>1 host logic of the train-sharded kNN-2 (shard bounds, global index
offsets, all-gather, (distance, index) merge) gives exactly the single-process brute-force answer."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from visual_sgraphs_b200 import sharded  # noqa: E402
from visual_sgraphs_b200.synth import synth_query_train  # noqa: E402

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def knn2_numpy(q, t, offset=0):
    nq = q.shape[0]
    idx = np.full((nq, 2), -1, np.int32)
    d = np.full((nq, 2), np.iinfo(np.int32).max, np.int32)
    if t.shape[0]:
        dm = POP[np.bitwise_xor(q[:, None, :], t[None, :, :])].sum(-1).astype(np.int64)
        key = dm * (1 << 32) + np.arange(t.shape[0])[None, :]
        order = np.argsort(key, axis=1, kind="stable")[:, :2]
        k = order.shape[1]
        idx[:, :k] = order + offset
        d[:, :k] = np.take_along_axis(dm, order, 1)
    return idx, d


def _worker(rank, world, port, nq, nt, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, t = synth_query_train(5, nq, nt)
    if nt > 7:
        t[7] = t[nt - 3]                   # a tie across shards must resolve to the lower global index
        q[0] = t[7]
    b, e = sharded.shard_bounds(nt, world)[rank]

    def local(qq, tt, off):
        i, d = knn2_numpy(qq, tt, off)
        return torch.from_numpy(i), torch.from_numpy(d)

    def merge(ip, dp):
        return sharded.merge_top2_numpy(ip.numpy(), dp.numpy())

    idx, d = sharded.knn2_sharded(dist, q, t[b:e], b, local, merge, lambda shape: torch.zeros(shape, dtype=torch.int32))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), idx=idx, dist=d)
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds():
    assert sharded.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert sharded.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert sharded.frame_shard(1000, 7, 8) == (875, 1000)


@pytest.mark.parametrize("nt", [301, 1])
def test_train_sharded_knn2_over_gloo(tmp_path, nt):
    world, nq = 2, 64
    mp.spawn(_worker, args=(world, _free_port(), nq, nt, str(tmp_path)), nprocs=world, join=True)
    q, t = synth_query_train(5, nq, nt)
    if nt > 7:
        t[7] = t[nt - 3]
        q[0] = t[7]
    want_i, want_d = knn2_numpy(q, t)
    for r in range(world):
        got = np.load(str(tmp_path / ("rank%d.npz" % r)))
        assert np.array_equal(got["idx"], want_i)
        assert np.array_equal(got["dist"], want_d)
    if nt > 7:
        assert want_i[0, 0] == 7 and want_i[0, 1] == nt - 3 and want_d[0, 0] == 0 and want_d[0, 1] == 0
