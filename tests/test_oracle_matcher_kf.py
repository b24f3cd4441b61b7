"""CPU-only: pins the oracle's keyframe-side matcher restatements (oracle/match_oracle.cpp) with independent
Python restatements of the reference loops on small inputs (the reference ships no fixtures for them)."""
import numpy as np

from tests import match_scenarios as sc
from tests.test_oracle_matcher import hamming, py_features_in_area

f32 = np.float32
SIZE = (322, 243)


def _frames(oracle, shift=(9, 5)):
    ka, da, kb, db = sc.two_frames(oracle, shift=shift, size=SIZE, nfeat=400)
    return ka, da, kb, db


def c_round(v):
    return int(np.floor(abs(v) + 0.5) * (1 if v >= 0 else -1))


def py_rot_bin(a1, a2):
    rot = f32(f32(a1) - f32(a2))
    if rot < 0:
        rot = f32(rot + f32(360.0))
    b = c_round(float(f32(rot * f32(f32(1.0) / f32(30)))))
    return 0 if b == 30 else b


def py_three_maxima(sizes):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
        elif s > m2:
            m3, m2, i3, i2 = m2, s, i2, i
        elif s > m3:
            m3, i3 = s, i
    if m2 < f32(0.1) * f32(m1):
        i2 = i3 = -1
    elif m3 < f32(0.1) * f32(m1):
        i3 = -1
    return i1, i2, i3


def rot_filter(hist, drop):
    keep = py_three_maxima([len(h) for h in hist])
    n = 0
    for i, h in enumerate(hist):
        if i in keep:
            continue
        for idx in h:
            drop(idx)
            n += 1
    return n


def test_reloc_projection(oracle):
    ka, da, kb, db = _frames(oracle)
    fd = sc.frame_data(ka, da, size=SIZE)
    pts = sc.search_points(kb, (9, 5), 3, size=SIZE)
    occ = (np.random.default_rng(4).random(fd.n) < 0.1).astype(np.uint8)
    for check_ori in (True, False):
        nm, assign = oracle.search_by_projection_reloc(fd.view, occ, pts, db, 10.0, 90, check_ori)
        taken = occ.astype(bool).copy()
        want = np.full(fd.n, -1, np.int32)
        hist = [[] for _ in range(30)]
        wnm = 0
        for i, p in enumerate(pts):
            if not p["valid"]:
                continue
            lvl = int(p["level"])
            cand = py_features_in_area(fd, p["u"], p["v"], f32(f32(10.0) * fd.scale_factors[lvl]), lvl - 1, lvl + 1)
            best, bi = 256, -1
            for i2 in cand:
                if taken[i2]:
                    continue
                d = hamming(db[i], da[i2])
                if d < best:
                    best, bi = d, i2
            if best <= 90:
                want[bi] = i
                taken[bi] = True
                wnm += 1
                hist[py_rot_bin(p["angle"], ka["angle"][bi])].append(bi)
        if check_ori:
            def drop(idx):
                want[idx] = -2
            wnm -= rot_filter(hist, drop)
        assert nm == wnm and np.array_equal(assign, want)
        assert nm > 30


def test_sim3_projection_and_fuse(oracle):
    ka, da, kb, db = _frames(oracle)
    scale, sigma2, inv_sigma2 = sc.sigma_tables()
    for stereo in (False, True):
        fd = sc.frame_data(ka, da, size=SIZE, stereo_seed=8 if stereo else None)
        pts = sc.search_points(kb, (9, 5), 5, sigma=1.5, size=SIZE)
        if stereo:
            pts["ur"] = pts["u"] - 12.0
        matched = (np.random.default_rng(6).random(fd.n) < 0.2).astype(np.uint8)
        # SearchByProjection(KF, Scw, ...) :430-528
        nm, assign = oracle.search_by_projection_sim3(fd.view, matched, pts, db, 8, float(f32(1.2)))
        taken = matched.astype(bool).copy()
        want = np.full(fd.n, -1, np.int32)
        wnm = 0
        for i, p in enumerate(pts):
            if not p["valid"]:
                continue
            lvl = int(p["level"])
            cand = py_features_in_area(fd, p["u"], p["v"], f32(f32(8) * scale[lvl]), -1, -1)
            best, bi = 256, -1
            for idx in cand:
                if taken[idx] or not (lvl - 1 <= ka["octave"][idx] <= lvl):
                    continue
                d = hamming(db[i], da[idx])
                if d < best:
                    best, bi = d, idx
            if f32(best) <= f32(50) * f32(1.2):
                want[bi] = i
                taken[bi] = True
                wnm += 1
        assert nm == wnm and np.array_equal(assign, want) and nm > 20
        # Fuse :1148-1335 (variant 0) and :1337-1446 (variant 1)
        for variant in (0, 1):
            nf, best_idx = oracle.fuse_search(fd.view, pts, db, 3.0, inv_sigma2, variant)
            want = np.full(len(pts), -1, np.int32)
            for i, p in enumerate(pts):
                if not p["valid"]:
                    continue
                lvl = int(p["level"])
                cand = py_features_in_area(fd, p["u"], p["v"], f32(f32(3.0) * scale[lvl]), -1, -1)
                best, bi = (256 if variant == 0 else 2 ** 31 - 1), -1
                for idx in cand:
                    kl = int(ka["octave"][idx])
                    if kl < lvl - 1 or kl > lvl:
                        continue
                    if variant == 0:
                        ex, ey = f32(p["u"] - ka["x"][idx]), f32(p["v"] - ka["y"][idx])
                        if fd.u_right is not None and fd.u_right[idx] >= 0:
                            er = f32(p["ur"] - fd.u_right[idx])
                            e2 = f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er))
                            if float(f32(e2 * inv_sigma2[kl])) > 7.8:
                                continue
                        else:
                            e2 = f32(f32(ex * ex) + f32(ey * ey))
                            if float(f32(e2 * inv_sigma2[kl])) > 5.99:
                                continue
                    d = hamming(db[i], da[idx])
                    if d < best:
                        best, bi = d, idx
                if best <= 50:
                    want[i] = bi
            assert np.array_equal(best_idx, want) and nf == int((want >= 0).sum())
            assert nf > 5


def test_search_by_sim3(oracle):
    ka, da, kb, db = _frames(oracle)
    f1, f2 = sc.frame_data(ka, da, size=SIZE), sc.frame_data(kb, db, size=SIZE)
    pts1 = sc.search_points(ka, (-9, -5), 9, size=SIZE)
    pts2 = sc.search_points(kb, (9, 5), 10, size=SIZE)
    nf, m12 = oracle.search_by_sim3(f1.view, f2.view, pts1, da, pts2, db, 7.5)

    def one_way(pts, desc, target, tdesc):
        out = np.full(len(pts), -1, np.int32)
        for i, p in enumerate(pts):
            if not p["valid"]:
                continue
            lvl = int(p["level"])
            cand = py_features_in_area(target, p["u"], p["v"], f32(f32(7.5) * target.scale_factors[lvl]), -1, -1)
            best, bi = 2 ** 31 - 1, -1
            for idx in cand:
                if not (lvl - 1 <= target.keys["octave"][idx] <= lvl):
                    continue
                d = hamming(desc[i], tdesc[idx])
                if d < best:
                    best, bi = d, idx
            if best <= 100:
                out[i] = bi
        return out

    a, b = one_way(pts1, da, f2, db), one_way(pts2, db, f1, da)
    want = np.array([a[i] if a[i] >= 0 and b[a[i]] == i else -1 for i in range(len(a))], np.int32)
    assert np.array_equal(m12, want) and nf == int((want >= 0).sum()) and nf > 30


def _walk(fv1, fv2):
    d2 = {int(n): fv2[2][fv2[1][k]:fv2[1][k + 1]] for k, n in enumerate(fv2[0])}
    for k, n in enumerate(fv1[0]):
        if int(n) in d2:
            yield fv1[2][fv1[1][k]:fv1[1][k + 1]], d2[int(n)]


def test_bow_kf_and_triangulation(oracle):
    shift = (9, 1)
    ka, da, kb, db = _frames(oracle, shift)
    k1 = sc.frame_data(ka, da, size=SIZE, stereo_seed=2)
    k2 = sc.frame_data(kb, db, size=SIZE, stereo_seed=3)
    rng = np.random.default_rng(12)
    v1, v2 = (rng.random(k1.n) < 0.85).astype(np.uint8), (rng.random(k2.n) < 0.85).astype(np.uint8)
    fv1, fv2 = sc.feature_vector(da, 12), sc.feature_vector(db, 12)
    # SearchByBoW(KF, KF) :758-900
    for check_ori in (True, False):
        nm, m12 = oracle.search_by_bow_kf(k1.view, v1, k2.view, v2, fv1, fv2, float(f32(0.8)), check_ori)
        want = np.full(k1.n, -1, np.int32)
        matched2 = np.zeros(k2.n, bool)
        hist = [[] for _ in range(30)]
        wnm = 0
        for l1, l2 in _walk(fv1, fv2):
            for i1 in l1:
                if not v1[i1]:
                    continue
                b1, b2, bi = 256, 256, -1
                for i2 in l2:
                    if matched2[i2] or not v2[i2]:
                        continue
                    d = hamming(da[i1], db[i2])
                    if d < b1:
                        b2, b1, bi = b1, d, i2
                    elif d < b2:
                        b2 = d
                if b1 < 50 and f32(b1) < f32(0.8) * f32(b2):
                    want[i1] = bi
                    matched2[bi] = True
                    wnm += 1
                    hist[py_rot_bin(ka["angle"][i1], kb["angle"][bi])].append(i1)
        if check_ori:
            def drop(i1):
                want[i1] = -1
            wnm -= rot_filter(hist, drop)
        assert nm == wnm and np.array_equal(m12, want) and nm > 10
    # SearchForTriangulation :902-1146
    h1, h2 = (rng.random(k1.n) < 0.3).astype(np.uint8), (rng.random(k2.n) < 0.3).astype(np.uint8)
    scale, sigma2, _ = sc.sigma_tables()
    F, ep = sc.translation_f12(shift), np.array([150.0, 100.0], np.float32)
    for only_stereo, coarse in ((False, False), (True, False), (False, True)):
        nm, m12 = oracle.search_for_triangulation(k1.view, h1, k2.view, h2, fv1, fv2, only_stereo, coarse, F, ep, sigma2, True)
        want = np.full(k1.n, -1, np.int32)
        hist = [[] for _ in range(30)]
        wnm = 0
        for l1, l2 in _walk(fv1, fv2):
            for i1 in l1:
                s1 = k1.u_right[i1] >= 0
                if h1[i1] or (only_stereo and not s1):
                    continue
                best, bi = 50, -1
                for i2 in l2:
                    s2 = k2.u_right[i2] >= 0
                    if h2[i2] or (only_stereo and not s2):
                        continue
                    d = hamming(da[i1], db[i2])
                    if d > 50 or d > best:
                        continue
                    x2, y2 = kb["x"][i2], kb["y"][i2]
                    if not s1 and not s2:
                        ex, ey = f32(ep[0] - x2), f32(ep[1] - y2)
                        if f32(f32(ex * ex) + f32(ey * ey)) < f32(f32(100) * scale[kb["octave"][i2]]):
                            continue
                    ok = coarse
                    if not ok:
                        x1, y1 = ka["x"][i1], ka["y"][i1]
                        a = f32(f32(f32(x1 * F[0]) + f32(y1 * F[3])) + F[6])
                        b = f32(f32(f32(x1 * F[1]) + f32(y1 * F[4])) + F[7])
                        c = f32(f32(f32(x1 * F[2]) + f32(y1 * F[5])) + F[8])
                        num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
                        den = f32(f32(a * a) + f32(b * b))
                        if den != 0:
                            ok = float(f32(f32(num * num) / den)) < 3.84 * float(sigma2[kb["octave"][i2]])
                    if ok:
                        best, bi = d, i2
                if bi >= 0:
                    want[i1] = bi
                    wnm += 1
                    hist[py_rot_bin(ka["angle"][i1], kb["angle"][bi])].append(i1)

        def drop(i1):
            want[i1] = -1
        wnm -= rot_filter(hist, drop)
        assert nm == wnm and np.array_equal(m12, want), (only_stereo, coarse)
        assert nm > 3


def test_search_by_projection_last_two_cameras(oracle):
    """ORBmatcher.cc:1667-1878 with CurrentFrame.Nleft != -1 against an independent restatement."""
    from tests.test_oracle_matcher import py_features_in_area
    sc2 = sc.two_camera_last_scene(oracle, size=(322, 243), nfeat=300)
    fl, fr, pl, pr, desc = sc2["fl"], sc2["fr"], sc2["pl"], sc2["pr"], sc2["desc"]
    nl = fl.n
    for mode in (0, 1, 2):
        for check_ori in (True, False):
            blocked = sc2["occupied"].astype(bool).copy()
            assign = np.full(fl.n + fr.n, -1, np.int32)
            hist = [[] for _ in range(30)]
            nm = 0
            for i in range(len(pl)):
                p = pl[i]
                if not p["valid"]:
                    continue
                oct_ = int(p["octave"])
                radius = f32(f32(15.0) * fl.scale_factors[oct_])
                lo, hi = {0: (oct_ - 1, oct_ + 1), 1: (oct_, -1), 2: (0, oct_)}[mode]
                cand = py_features_in_area(fl, p["u"], p["v"], radius, lo, hi)
                if not cand:
                    continue                                     # skips the right camera as well
                for fd, cands, off in ((fl, cand, 0), (fr, None, nl)):
                    if cands is None:
                        cands = py_features_in_area(fr, pr[i]["u"], pr[i]["v"], radius, lo, hi)
                    best, bi = 256, -1
                    for i2 in cands:
                        if blocked[i2 + off]:
                            continue
                        d = hamming(desc[i], fd.descriptors[i2])
                        if d < best:
                            best, bi = d, i2
                    if best <= 100:
                        assign[bi + off] = i
                        blocked[bi + off] = bool(p["blocks"])
                        nm += 1
                        if check_ori:
                            hist[py_rot_bin(p["angle"], fd.keys[bi]["angle"])].append(bi + off)
            if check_ori:
                def drop(idx):
                    assign[idx] = -2
                nm -= rot_filter(hist, drop)
            got_nm, got = oracle.search_by_projection_last_2cam(fl.view, fr.view, sc2["occupied"], pl, pr, desc, 15.0, mode, check_ori)
            assert got_nm == nm and np.array_equal(got, assign)
            assert (assign[:nl] >= 0).sum() > 20 and (assign[nl:] >= 0).sum() > 20


def py_search_by_bow(kf, valid, f, kfv, ffv, nnratio, check_ori, f_nleft):
    """Independent restatement of ORBmatcher.cc:226-428 incl. the two-camera branch (f_nleft != -1)."""
    kn, kp, ki = kfv
    fn, fp, fi = ffv
    fbucket = {int(n): fi[fp[j]:fp[j + 1]] for j, n in enumerate(fn)}
    matches = np.full(f.n, -1, np.int32)
    hist = [[] for _ in range(30)]
    nm = 0
    for a, node in enumerate(kn):
        if int(node) not in fbucket:
            continue
        for real_kf in ki[kp[a]:kp[a + 1]]:
            if not valid[real_kf]:
                continue
            b = {False: [256, -1, 256], True: [256, -1, 256]}       # per camera: best, idx, second
            for real_f in fbucket[int(node)]:
                if matches[real_f] >= 0:
                    continue
                d = hamming(kf.descriptors[real_kf], f.descriptors[real_f])
                cam = f_nleft != -1 and real_f >= f_nleft
                if d < b[cam][0]:
                    b[cam] = [d, int(real_f), b[cam][0]]
                elif d < b[cam][2]:
                    b[cam][2] = d
            if b[False][0] <= 50:
                if f32(b[False][0]) < f32(nnratio) * f32(b[False][2]):
                    matches[b[False][1]] = real_kf
                    nm += 1
                    if check_ori:
                        hist[py_rot_bin(kf.keys[real_kf]["angle"], f.keys[b[False][1]]["angle"])].append(b[False][1])
                if b[True][0] <= 50:                                  # nested in the left block, ratio test `|| true`
                    matches[b[True][1]] = real_kf
                    nm += 1
                    if check_ori:
                        hist[py_rot_bin(kf.keys[real_kf]["angle"], f.keys[b[True][1]]["angle"])].append(b[True][1])
    if check_ori:
        def drop(idx):
            matches[idx] = -1
        nm -= rot_filter(hist, drop)
    return nm, matches


def test_search_by_bow_single_and_two_cameras(oracle):
    ka, da, kb, db = _frames(oracle, shift=(3, 2))
    # F: a two-camera frame whose left camera sees frame A and whose right camera sees frame B
    keys = np.concatenate([ka, kb])
    desc = np.concatenate([da, db])
    f = sc.frame_data(keys, desc, size=(322, 243))
    kf = sc.frame_data(kb, db, size=(322, 243))
    rng = np.random.default_rng(12)
    valid = (rng.random(kf.n) < 0.85).astype(np.uint8)
    kfv, ffv = sc.feature_vector(db, 12), sc.feature_vector(desc, 12)
    for f_nleft in (-1, len(ka)):
        for check_ori in (True, False):
            nm, mf = oracle.search_by_bow(kf.view, valid, f.view, kfv, ffv, float(f32(0.7)), check_ori, f_nleft)
            wnm, wmf = py_search_by_bow(kf, valid, f, kfv, ffv, 0.7, check_ori, f_nleft)
            assert nm == wnm and np.array_equal(mf, wmf)
            assert nm > 10
            if f_nleft != -1:
                assert (mf[f_nleft:] >= 0).sum() > 5 and (mf[:f_nleft] >= 0).sum() > 5
