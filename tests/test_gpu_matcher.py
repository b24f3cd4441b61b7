"""GPU parity tests for the Hamming matching kernels (bit-exact integer work) against the oracle's
DescriptorDistance and numpy restatements of the candidate scans."""
import numpy as np
import pytest

from visual_sgraphs_b200.synth import synth_descriptors, synth_query_train

pytestmark = pytest.mark.gpu

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_matrix(q, t):
    return POP[np.bitwise_xor(q[:, None, :], t[None, :, :])].sum(-1).astype(np.int32)


def knn2_ref(q, t):
    d = hamming_matrix(q, t)
    key = d.astype(np.int64) * (1 << 32) + np.arange(t.shape[0])[None, :]
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    idx = order.astype(np.int32)
    dist = np.take_along_axis(d, order, 1)
    return idx, dist


def _matcher():
    from visual_sgraphs_b200.matcher import ORBmatcher
    return ORBmatcher()


def test_descriptor_distance_matches_oracle(oracle):
    a, b = synth_descriptors(1, 3000), synth_descriptors(2, 3000)
    b[:100] = a[:100]
    a[100] = 0
    b[100] = 255
    got = _matcher().DescriptorDistance(a, b)
    want = np.array([oracle.descriptor_distance(a[i], b[i]) for i in range(len(a))], np.int32)
    assert np.array_equal(got, want)
    assert got[0] == 0 and got[100] == 256


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 2), (7, 5), (100, 1000), (1000, 4097), (1030, 20000), (3, 0)])
def test_knn2_matches_bruteforce(nq, nt):
    q, t = synth_query_train(nq * 31 + nt, nq, max(nt, 1))
    t = t[:nt]
    if nt >= 8:
        t[3] = t[1]          # exact duplicates: ties must resolve to the lower train index
        q[0] = t[1]
    idx, dist = _matcher().knn2(q, t)
    if nt == 0:
        assert (idx == -1).all() and (dist == np.iinfo(np.int32).max).all()
        return
    widx, wdist = knn2_ref(q, t)
    if nt == 1:
        assert np.array_equal(idx[:, 0], widx[:, 0]) and np.array_equal(dist[:, 0], wdist[:, 0])
        assert (idx[:, 1] == -1).all()
        return
    assert np.array_equal(dist, wdist)
    assert np.array_equal(idx, widx)


def test_knn2_index_offset_and_merge_of_shards():
    torch = pytest.importorskip("torch")
    q, t = synth_query_train(77, 500, 6000)
    m = _matcher()
    widx, wdist = knn2_ref(q, t)
    shards = np.array_split(np.arange(6000), 4)
    dq = torch.from_numpy(q).cuda()
    parts_i = torch.zeros((4, 500, 2), dtype=torch.int32, device="cuda")
    parts_d = torch.zeros((4, 500, 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for s, rows in enumerate(shards):
        dt = torch.from_numpy(t[rows]).cuda()
        torch.cuda.synchronize()
        m.knn2_dev(dq, dt, parts_i[s], parts_d[s], train_index_offset=int(rows[0]))
        m.sync()
    oi = torch.zeros((500, 2), dtype=torch.int32, device="cuda")
    od = torch.zeros((500, 2), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    m.knn2_merge_dev(parts_i, parts_d, oi, od)
    m.sync()
    assert np.array_equal(oi.cpu().numpy(), widx)
    assert np.array_equal(od.cpu().numpy(), wdist)


def window_ref(q, t, cand_ptr, cand, skip, level, init):
    nq = q.shape[0]
    out = {k: np.zeros(nq, np.int32) for k in ("best_idx", "best_dist", "second_dist", "best_level", "second_level")}
    for i in range(nq):
        bd = bd2 = init
        bi = bl = bl2 = -1
        for c in range(cand_ptr[i], cand_ptr[i + 1]):
            j = cand[c]
            if skip is not None and skip[j]:
                continue
            d = int(POP[np.bitwise_xor(q[i], t[j])].sum())
            lv = int(level[j]) if level is not None else -1
            if d < bd:
                bd2, bd, bl2, bl, bi = bd, d, bl, lv, j
            elif d < bd2:
                bl2, bd2 = lv, d
        out["best_idx"][i], out["best_dist"][i], out["second_dist"][i] = bi, bd, bd2
        out["best_level"][i], out["second_level"][i] = bl, bl2
    return out


@pytest.mark.parametrize("init", [256, np.iinfo(np.int32).max])
def test_match_window_matches_sequential_scan(init):
    rng = np.random.default_rng(5)
    q, t = synth_query_train(9, 800, 1000, related_frac=0.5)
    counts = rng.integers(0, 25, 800)
    cand_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    cand = rng.integers(0, 1000, cand_ptr[-1]).astype(np.int32)
    skip = (rng.random(1000) < 0.2).astype(np.uint8)
    level = rng.integers(0, 8, 1000).astype(np.int32)
    got = _matcher().match_window(q, t, cand_ptr, cand, skip, level, init)
    want = window_ref(q, t, cand_ptr, cand, skip, level, init)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    got2 = _matcher().match_window(q, t, cand_ptr, cand, None, None, init)
    want2 = window_ref(q, t, cand_ptr, cand, None, None, init)
    for k in want2:
        assert np.array_equal(got2[k], want2[k]), k
