"""CPU-only: pins the oracle's matcher restatements (oracle/match_oracle.cpp) with direct, independent
Python restatements of the reference loops on small inputs (the reference ships no fixtures for them)."""
import numpy as np

from tests import match_scenarios as sc

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming(a, b):
    return int(POP[np.bitwise_xor(a, b)].sum())


def py_features_in_area(fd, x, y, r, min_level, max_level):
    """Frame::GetFeaturesInArea + AssignFeaturesToGrid + PosInGrid (Frame.cc:521-553, 802-880), float32 arithmetic."""
    f32 = np.float32
    v = fd.view
    x, y, r = f32(x), f32(y), f32(r)
    cols, rows = v.grid_cols, v.grid_rows
    minx, miny, iw, ih = f32(v.min_x), f32(v.min_y), f32(v.grid_inv_w), f32(v.grid_inv_h)
    grid = {}
    for i, k in enumerate(fd.keys):
        px = int(np.floor(f32((k["x"] - minx) * iw) + f32(0.5))) if (k["x"] - minx) * iw >= 0 else int(np.ceil(f32((k["x"] - minx) * iw) - f32(0.5)))
        py = int(np.floor(f32((k["y"] - miny) * ih) + f32(0.5))) if (k["y"] - miny) * ih >= 0 else int(np.ceil(f32((k["y"] - miny) * ih) - f32(0.5)))
        if 0 <= px < cols and 0 <= py < rows:
            grid.setdefault((px, py), []).append(i)
    c0 = max(0, int(np.floor(f32(f32(x - minx) - r) * iw)))
    c1 = min(cols - 1, int(np.ceil(f32(f32(x - minx) + r) * iw)))
    r0 = max(0, int(np.floor(f32(f32(y - miny) - r) * ih)))
    r1 = min(rows - 1, int(np.ceil(f32(f32(y - miny) + r) * ih)))
    if c0 >= cols or c1 < 0 or r0 >= rows or r1 < 0:
        return []
    check = (min_level > 0) or (max_level >= 0)
    out = []
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            for i in grid.get((ix, iy), []):
                k = fd.keys[i]
                if check:
                    if k["octave"] < min_level:
                        continue
                    if max_level >= 0 and k["octave"] > max_level:
                        continue
                if abs(f32(k["x"] - x)) < r and abs(f32(k["y"] - y)) < r:
                    out.append(i)
    return out


def test_descriptor_distance_is_popcount(oracle):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for i in range(200):
        assert oracle.descriptor_distance(a[i], b[i]) == hamming(a[i], b[i])
    assert oracle.descriptor_distance(a[0], a[0]) == 0
    assert oracle.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_three_maxima_quirks(oracle):
    assert oracle.three_maxima([0] * 30) == (-1, -1, -1)
    s = [0] * 30
    s[3], s[7], s[9] = 50, 40, 30
    assert oracle.three_maxima(s) == (3, 7, 9)
    s[7], s[9] = 4, 3                      # below 10 % of the maximum: dropped (ORBmatcher.cc:2033-2042)
    assert oracle.three_maxima(s) == (3, -1, -1)
    s[7] = 5
    assert oracle.three_maxima(s) == (3, 7, -1)
    t = [0] * 30
    t[1] = t[2] = t[3] = t[4] = 10          # ties: strict '>' keeps the earliest bins
    assert oracle.three_maxima(t) == (1, 2, 3)


def test_features_in_area_matches_python_restatement(oracle):
    ka, da, kb, db = sc.two_frames(oracle)
    fd = sc.frame_data(ka, da)
    rng = np.random.default_rng(3)
    for _ in range(150):
        x, y = rng.uniform(-20, 660), rng.uniform(-20, 500)
        r = rng.uniform(1, 60)
        lo, hi = [(-1, -1), (0, 0), (2, 3), (1, -1), (0, 4)][rng.integers(0, 5)]
        got = oracle.get_features_in_area(fd.view, x, y, r, lo, hi).tolist()
        assert got == py_features_in_area(fd, x, y, r, lo, hi)


def py_search_by_projection_map(fd, occupied, pts, desc, th, far, th_far, nnratio):
    f32 = np.float32
    blocked = occupied.astype(bool).copy()
    assign = np.full(fd.n, -1, np.int32)
    nm = 0
    for i, mp in enumerate(pts):
        if not mp["in_view"] or (far and mp["depth"] > th_far) or mp["bad"]:
            continue
        lvl = int(mp["level"])
        r = f32(2.5) if float(mp["view_cos"]) > 0.998 else f32(4.0)
        if th != 1.0:
            r = f32(r * f32(th))
        win = f32(r * fd.scale_factors[lvl])
        cand = py_features_in_area(fd, mp["proj_x"], mp["proj_y"], win, lvl - 1, lvl)
        best, bl, best2, bl2, bi = 256, -1, 256, -1, -1
        for idx in cand:
            if blocked[idx]:
                continue
            if fd.u_right is not None and fd.u_right[idx] > 0:
                if abs(f32(mp["proj_xr"] - fd.u_right[idx])) > win:
                    continue
            d = hamming(desc[i], fd.descriptors[idx])
            if d < best:
                best2, best, bl2, bl, bi = best, d, bl, int(fd.keys[idx]["octave"]), idx
            elif d < best2:
                bl2, best2 = int(fd.keys[idx]["octave"]), d
        if best <= 100:
            if bl == bl2 and f32(best) > f32(nnratio) * f32(best2):
                continue
            if bl != bl2 or f32(best) <= f32(nnratio) * f32(best2):
                assign[bi] = i
                blocked[bi] = bool(mp["blocks"])
                nm += 1
    return nm, assign


def test_search_by_projection_map_matches_python_restatement(oracle):
    ka, da, kb, db = sc.two_frames(oracle, size=(322, 243), nfeat=400)
    for stereo in (False, True):
        fd = sc.frame_data(ka, da, size=(322, 243), stereo_seed=5 if stereo else None)
        pts, desc, occ = sc.track_points(fd, kb, db, (9, 5), 21, stereo)
        for th, far in ((3.0, False), (1.0, True)):
            nm, assign = oracle.search_by_projection_map(fd.view, occ, pts, desc, th, far, 40.0, 0.8)
            wnm, wassign = py_search_by_projection_map(fd, occ, pts, desc, th, far, 40.0, 0.8)
            assert nm == wnm and np.array_equal(assign, wassign)
            assert nm > 20


def py_search_by_projection_map_2cam(sc2, th, far, th_far, nnratio):
    """Independent restatement of ORBmatcher.cc:42-216 for F.Nleft != -1."""
    f32 = np.float32
    fl, fr, pl, pr, desc = sc2["fl"], sc2["fr"], sc2["pl"], sc2["pr"], sc2["desc"]
    nl = fl.n
    blocked = sc2["occupied"].astype(bool).copy()
    assign = np.full(fl.n + fr.n, -1, np.int32)
    nm = 0

    def scan(fd, cand, i, offset):
        best, bl, best2, bl2, bi = 256, -1, 256, -1, -1
        for idx in cand:
            if blocked[idx + offset]:
                continue
            d = hamming(desc[i], fd.descriptors[idx])
            if d < best:
                best2, best, bl2, bl, bi = best, d, bl, int(fd.keys[idx]["octave"]), idx
            elif d < best2:
                bl2, best2 = int(fd.keys[idx]["octave"]), d
        return best, bl, best2, bl2, bi

    for i in range(len(pl)):
        l, r = pl[i], pr[i]
        if (not l["in_view"] and not r["in_view"]) or (far and l["depth"] > th_far) or l["bad"]:
            continue
        blocks = bool(l["blocks"])
        if l["in_view"]:
            lvl = int(l["level"])
            rad = f32(2.5) if float(l["view_cos"]) > 0.998 else f32(4.0)
            if th != 1.0:
                rad = f32(rad * f32(th))
            cand = py_features_in_area(fl, l["proj_x"], l["proj_y"], f32(rad * fl.scale_factors[lvl]), lvl - 1, lvl)
            if cand:
                best, bl, best2, bl2, bi = scan(fl, cand, i, 0)
                if best <= 100:
                    if bl == bl2 and f32(best) > f32(nnratio) * f32(best2):
                        continue                                   # the reference's `continue` also skips the right camera
                    assign[bi] = i
                    blocked[bi] = blocks
                    if sc2["l2r"][bi] != -1:
                        assign[sc2["l2r"][bi] + nl] = i
                        blocked[sc2["l2r"][bi] + nl] = blocks
                        nm += 1
                    nm += 1
        if r["in_view"] and int(r["level"]) != -1:
            lvl = int(r["level"])
            rad = f32(2.5) if float(r["view_cos"]) > 0.998 else f32(4.0)
            cand = py_features_in_area(fr, r["proj_x"], r["proj_y"], f32(rad * fr.scale_factors[lvl]), lvl - 1, lvl)
            if not cand:
                continue
            best, bl, best2, bl2, bi = scan(fr, cand, i, nl)
            if best <= 100:
                if bl == bl2 and f32(best) > f32(nnratio) * f32(best2):
                    continue
                if sc2["r2l"][bi] != -1:
                    assign[sc2["r2l"][bi]] = i
                    blocked[sc2["r2l"][bi]] = blocks
                    nm += 1
                assign[bi + nl] = i
                blocked[bi + nl] = blocks
                nm += 1
    return nm, assign


def test_search_by_projection_map_two_cameras_matches_python_restatement(oracle):
    sc2 = sc.two_camera_scene(oracle, size=(322, 243), nfeat=300)
    for th, far in ((3.0, False), (1.0, True)):
        nm, assign = oracle.search_by_projection_map_2cam(sc2["fl"].view, sc2["fr"].view, sc2["occupied"], sc2["l2r"], sc2["r2l"],
                                                          sc2["pl"], sc2["pr"], sc2["desc"], th, far, 40.0, 0.8)
        wnm, wassign = py_search_by_projection_map_2cam(sc2, th, far, 40.0, 0.8)
        assert nm == wnm and np.array_equal(assign, wassign)
        nl = sc2["fl"].n
        assert (assign[:nl] >= 0).sum() > 20 and (assign[nl:] >= 0).sum() > 20


def test_distinctive_descriptor_matches_python_restatement(oracle):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:383-409): least median distance, first row on ties."""
    rng = np.random.default_rng(31)
    assert oracle.distinctive_descriptor(np.zeros((0, 32), np.uint8)) == -1
    for n in (1, 2, 3, 4, 7, 16, 33):
        for trial in range(6):
            base = rng.integers(0, 256, 32, dtype=np.uint8)
            d = np.stack([base ^ np.packbits(rng.random(256) < rng.uniform(0.02, 0.3)) for _ in range(n)])
            if trial == 0 and n > 2:
                d[2] = d[0]                     # identical observations: ties in the medians
            dm = np.array([[hamming(a, b) for b in d] for a in d])
            med = [sorted(row)[int(0.5 * (n - 1))] for row in dm]
            assert oracle.distinctive_descriptor(d) == int(np.argmin(med))


def test_bow_tree_walk_matches_python_restatement(oracle):
    """TemplatedVocabulary::transform(feature, ...) (TemplatedVocabulary.h:1225-1265) on a synthetic unbalanced tree."""
    voc = sc.synthetic_vocabulary(4, k=6, levels=3)
    rng = np.random.default_rng(8)
    nd = voc["node_desc"]
    feats = np.stack([nd[rng.integers(1, len(nd))] ^ np.packbits(rng.random(256) < 0.1) for _ in range(150)])
    feats[0] = nd[voc["child_idx"][voc["child_ptr"][1]]]      # sits exactly on the duplicated child: tie -> first child
    for levelsup in (1, 2, 3, 5):
        leaf, nid = oracle.bow_transform(voc["child_ptr"], voc["child_idx"], nd, voc["levels"], feats, levelsup)
        for f in range(len(feats)):
            node, lvl, want_nid = 0, 0, 0
            while voc["child_ptr"][node] != voc["child_ptr"][node + 1]:
                lvl += 1
                kids = voc["child_idx"][voc["child_ptr"][node]:voc["child_ptr"][node + 1]]
                d = [hamming(feats[f], nd[c]) for c in kids]
                node = int(kids[int(np.argmin(d))])
                if lvl == voc["levels"] - levelsup:
                    want_nid = node
            assert leaf[f] == node and nid[f] == want_nid, (levelsup, f)


def test_cpu_knn2_baseline_matches_numpy(oracle):
    """The CPU baseline of the brute-force matcher (both distance variants, threaded) is the lexicographic (dist, idx)
    top-2 cv::BFMatcher::knnMatch returns."""
    from visual_sgraphs_b200.synth import synth_query_train
    q, t = synth_query_train(9, 60, 900)
    t[100] = t[7]                                           # exact ties between train rows
    d = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(2).astype(np.int64)
    order = np.lexsort((np.broadcast_to(np.arange(t.shape[0]), d.shape), d), axis=1)[:, :2]
    for variant in (0, 1):
        for threads in (1, 3):
            _, idx, dist = oracle.bench_knn2(q, t, threads, variant)
            assert np.array_equal(idx, order)
            assert np.array_equal(dist, np.take_along_axis(d, order, 1))
    _, idx, dist = oracle.bench_knn2(q, t[:1], 2, 0)        # a single train row: no second neighbour
    assert (idx[:, 1] == -1).all() and (dist[:, 1] == -1).all() and (idx[:, 0] == 0).all()
