"""The algebra behind csrc/blur_tc.cu, checked on the CPU against the oracle's GaussianBlur restatement (cv2-pinned):
OpenCV's 8-bit 7x7 sigma-2 blur (ORBextractor.cc:1129-1130) is exactly linear, so it equals two banded (Toeplitz) matrix products
with the 16-bit intermediate split into its byte planes, and the final rounding  (256 Vh + Vl + 32768) >> 16  equals
(Vh + (Vl >> 8) + 128) >> 8  on values that stay below 2^16 — what the kernel computes with u8 x u8 -> s32 tensor-core GEMMs."""
import numpy as np

TAPS = np.array([18, 34, 48, 56, 48, 34, 18], np.int64)


def _band(n_out, n_in, offset):
    """B[o][k] = TAPS[k - offset - o]: output o reads inputs offset + o .. offset + o + 6."""
    b = np.zeros((n_out, n_in), np.int64)
    for o in range(n_out):
        for j in range(7):
            if 0 <= offset + o + j < n_in:
                b[o, offset + o + j] = TAPS[j]
    return b


def test_blur_equals_two_banded_gemms_with_byte_planes(oracle):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (150, 131), dtype=np.uint8)
    img[:4] = 255                                           # saturated borders exercise the 16-bit bound
    img[:, -3:] = 255
    want = oracle.gaussian_blur7(img)
    h, w = img.shape
    # the kernel's tile: 122 x 96 outputs from a 128 x 128 window that starts 3 rows above and 16 columns left of the tile;
    # out-of-plane window bytes are the REFLECT_101 mirror images (the patch warps write them into shared memory)
    pad = np.pad(img, ((3, 128), (16, 128)), mode="reflect").astype(np.int64)
    bh, bv = _band(96, 128, 13), _band(122, 128, 0)         # Bh[n][k] = taps[k - 13 - n], Bv[r][k] = taps[k - r]
    assert bh.max() < 256 and bv.max() < 256                # u8 operands
    got = np.zeros_like(img)
    for y0 in range(0, h, 122):
        for x0 in range(0, w, 96):
            win = pad[y0:y0 + 128, x0:x0 + 128]             # rows y0-3.., columns x0-16..
            ht = bh @ win.T                                  # GEMM 1, transposed: Ht[n][i]
            assert ht.max() <= 65280
            hh, hl = ht >> 8, ht & 255                       # byte planes = u8 operands of GEMM 2
            vh, vl = bv @ hh.T, bv @ hl.T                    # V[r][n], one accumulator per plane
            assert vh.max() < 65536 and vl.max() < 65536     # tcgen05.ld .pack::16b keeps the low halves
            exact = (256 * vh + vl + 32768) >> 16
            packed = (vh + (vl >> 8) + 128) >> 8             # the epilogue's form on 16-bit lanes
            assert np.array_equal(exact, packed) and (vh + (vl >> 8) + 128).max() < 65536
            rows, cols = min(122, h - y0), min(96, w - x0)
            got[y0:y0 + rows, x0:x0 + cols] = packed[:rows, :cols]
    assert np.array_equal(got, want)
