"""Seeded synthetic inputs for tests and benchmarks.

Frames: uint8 uniform noise, Gaussian-smoothed (sigma 1.5), min-max normalised to [0,255], plus 60
random filled rectangles (side 8-80 px, random gray).  Frame i of a sequence uses seed 1000+i.
Descriptors: uniform random 256-bit rows; a fraction of the queries are train rows with a few bits
flipped so ratio tests and TH gates fire.
Pure numpy; no OpenCV, no torch.
"""
import numpy as np


def _gauss_kernel(sigma):
    r = int(np.ceil(3 * sigma))
    x = np.arange(-r, r + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _smooth(img, sigma):
    k = _gauss_kernel(sigma)
    r = len(k) // 2
    p = np.pad(img, ((0, 0), (r, r)), mode="reflect")
    out = np.zeros_like(img)
    for i, kv in enumerate(k):
        out += kv * p[:, i:i + img.shape[1]]
    p = np.pad(out, ((r, r), (0, 0)), mode="reflect")
    out2 = np.zeros_like(img)
    for i, kv in enumerate(k):
        out2 += kv * p[i:i + img.shape[0], :]
    return out2


def synth_frame(seed, width=640, height=480, nrect=60):
    """One corner-rich grayscale frame, uint8 (height, width)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (height, width), dtype=np.uint8).astype(np.float64)
    base = _smooth(base, 1.5)
    lo, hi = base.min(), base.max()
    img = np.round((base - lo) * (255.0 / max(hi - lo, 1e-9))).astype(np.uint8)
    for _ in range(nrect):
        rw, rh = rng.integers(8, 81, 2)
        x0 = int(rng.integers(0, max(1, width - 8)))
        y0 = int(rng.integers(0, max(1, height - 8)))
        img[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 256)
    return img


def synth_sequence(n, width=640, height=480, first_seed=1000):
    """(n, height, width) uint8; frame i uses seed first_seed + i."""
    return np.stack([synth_frame(first_seed + i, width, height) for i in range(n)])


def synth_stereo_pair(seed, width=752, height=480):
    """Left frame and a right frame = left shifted by a per-row-band disparity (5-40 px) plus +-2 noise."""
    left = synth_frame(seed, width, height)
    rng = np.random.default_rng(seed + 7_000_000)
    right = np.empty_like(left)
    y = 0
    while y < height:
        bh = int(rng.integers(16, 64))
        d = int(rng.integers(5, 41))
        rows = slice(y, min(height, y + bh))
        right[rows, : width - d] = left[rows, d:]
        right[rows, width - d:] = left[rows, width - d - 1: width - d]
        y += bh
    noise = rng.integers(-2, 3, right.shape)
    right = np.clip(right.astype(np.int16) + noise, 0, 255).astype(np.uint8)
    return left, right


def synth_descriptors(seed, n):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def synth_query_train(seed, nq, nt, related_frac=0.1, max_flips=40):
    """Train set + query set where `related_frac` of the queries are train rows with 0..max_flips bit flips."""
    rng = np.random.default_rng(seed)
    train = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    query = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    nrel = int(nq * related_frac)
    if nrel and nt:
        qi = rng.choice(nq, nrel, replace=False)
        ti = rng.integers(0, nt, nrel)
        bits = np.unpackbits(train[ti], axis=1)
        for r in range(nrel):
            k = int(rng.integers(0, max_flips + 1))
            if k:
                flip = rng.choice(256, k, replace=False)
                bits[r, flip] ^= 1
        query[qi] = np.packbits(bits, axis=1)
    return query, train
