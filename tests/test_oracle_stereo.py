"""CPU-only: the oracle's ComputeStereoMatches restatement (Frame.cc:957-1127) against an independent numpy
restatement on a synthetic stereo pair (the reference ships no fixture for it)."""
import numpy as np

from visual_sgraphs_b200.synth import synth_stereo_pair

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def c_round(v):
    return np.float32(np.floor(abs(v) + np.float32(0.5)) * (1 if v >= 0 else -1))


def py_stereo(exl, exr, kl, dl, kr, dr, mb, mbf):
    f32 = np.float32
    scale, inv = exl.tables()["scale"], exl.tables()["inv_scale"]
    n_rows = exl.level_size(0)[1]
    rows = [[] for _ in range(n_rows)]
    for ir in range(len(kr)):
        r = f32(2.0) * scale[kr["octave"][ir]]
        lo, hi = int(np.floor(f32(kr["y"][ir] - r))), int(np.ceil(f32(kr["y"][ir] + r)))
        for y in range(lo, hi + 1):
            if 0 <= y < n_rows:
                rows[y].append(ir)
    max_d = f32(f32(mbf) / f32(mb))
    u_right = np.full(len(kl), -1, np.float32)
    depth = np.full(len(kl), -1, np.float32)
    pairs = []
    levels_l = [exl.level(l) for l in range(8)]
    levels_r = [exr.level(l) for l in range(8)]
    for il in range(len(kl)):
        ul, vl, lvl = kl["x"][il], kl["y"][il], int(kl["octave"][il])
        cands = rows[int(vl)]
        if not cands or f32(ul) < 0:
            continue
        min_u, max_u = f32(ul - max_d), ul
        best, best_ir = 100, 0
        for ir in cands:
            if abs(int(kr["octave"][ir]) - lvl) > 1:
                continue
            if min_u <= kr["x"][ir] <= max_u:
                d = int(POP[np.bitwise_xor(dl[il], dr[ir])].sum())
                if d < best:
                    best, best_ir = d, ir
        if best >= 75:
            continue
        sf = inv[lvl]
        sul, svl, sur0 = c_round(f32(ul * sf)), c_round(f32(vl * sf)), c_round(f32(kr["x"][best_ir] * sf))
        PL, PR = levels_l[lvl], levels_r[lvl]
        if sur0 + 5 - 5 < 0 or sur0 + 5 + 5 + 1 >= PR.shape[1]:
            continue
        y0, xl0 = int(svl) - 5, int(sul) - 5
        IL = PL[y0:y0 + 11, xl0:xl0 + 11].astype(np.int32)
        sads = []
        for inc in range(-5, 6):
            xr0 = int(sur0) + inc - 5
            sads.append(int(np.abs(IL - PR[y0:y0 + 11, xr0:xr0 + 11].astype(np.int32)).sum()))
        b = int(np.argmin(sads))
        if b == 0 or b == 10:
            continue
        d1, d2, d3 = f32(sads[b - 1]), f32(sads[b]), f32(sads[b + 1])
        with np.errstate(divide="ignore", invalid="ignore"):
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
        if delta < -1 or delta > 1:
            continue
        best_ur = f32(scale[lvl] * f32(f32(sur0 + f32(b - 5)) + delta))
        disp = f32(ul - best_ur)
        if disp >= 0 and disp < max_d:
            if disp <= 0:
                disp = f32(0.01)
                best_ur = f32(np.float64(ul) - 0.01)
            depth[il] = f32(f32(mbf) / disp)
            u_right[il] = best_ur
            pairs.append((sads[b], il))
    if pairs:
        pairs.sort()
        th = f32(f32(1.5) * f32(1.4)) * f32(pairs[len(pairs) // 2][0])
        for sad, il in reversed(pairs):
            if f32(sad) < th:
                break
            u_right[il] = -1
            depth[il] = -1
    return u_right, depth


def test_stereo_matches_against_numpy_restatement(oracle):
    left, right = synth_stereo_pair(42, 376, 240)
    exl, exr = oracle.OracleExtractor(500), oracle.OracleExtractor(500)
    _, kl, dl = exl(left)
    _, kr, dr = exr(right)
    mb, mbf = 0.11, 47.9
    got_u, got_d = oracle.stereo_matches(exl, exr, kl, dl, kr, dr, mb, mbf)
    want_u, want_d = py_stereo(exl, exr, kl, dl, kr, dr, mb, mbf)
    assert np.array_equal(got_u, want_u)
    assert np.array_equal(got_d, want_d)
    assert (got_u >= 0).sum() > 50
