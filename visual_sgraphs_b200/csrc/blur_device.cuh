// blur_device.cuh — device code of the 7x7 Gaussian blur (GaussianBlur 7x7 sigma 2 BORDER_REFLECT_101,
// reference orb_slam3/src/ORBextractor.cc:1129-1130; OpenCV's 8-bit fixed-point path).
// Shared by the standalone blur kernel (pyramid.cu) and the fused FAST + blur kernel (fast.cu).
#pragma once
#include "vsg_internal.cuh"

namespace vsg {

// ------------------------------------------------------------------------------------------------
// blur: no shared memory.  One thread owns a strip of 4 columns x 32 rows and walks down it with a
// 7-row register window of horizontal sums: per source row three aligned 32-bit loads, the four
// horizontal sums by funnel-shift + two IDP4A each (taps packed as bytes), then four vertical sums
// and one 32-bit store.  REFLECT_101 reflects the level itself, not a padded buffer:
// rows by index arithmetic, columns by giving the first and the last one or two 4-column groups of a
// row to separate "edge" work items that assemble their 12 source bytes one by one — they are queued
// after all interior items so that interior warps never diverge.  One launch covers every level.
// ------------------------------------------------------------------------------------------------
constexpr int kBlurRows = 32, kBlurRowsSmallBatch = 8, kBlurThreads = 128;   // strip height: large batches / few frames

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

struct BlurLevels {
    int nlevels;
    int rows;                         // rows per strip (one thread walks a strip top to bottom)
    int block_begin[kMaxLevels + 1];  // prefix sums of thread blocks per level
    int n_int_cg[kMaxLevels];         // interior 4-column groups per row: cg = 1 .. n_int_cg
    int n_edge_cg[kMaxLevels];        // edge groups per row: cg = 0 and cg > n_int_cg
    int n_strips[kMaxLevels];
};

template <bool kEdge>
__device__ __forceinline__ void blur_load_row(const uint8_t *__restrict__ row, int x, int w, uint32_t h[4]) {
    uint32_t w0, w1, w2;
    if (!kEdge) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(row + x);
        w0 = __ldg(p - 1); w1 = __ldg(p); w2 = __ldg(p + 1);
    } else {
        uint32_t b[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) b[k] = __ldg(row + reflect101(min(x - 4 + k, w + 2), w));
        w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
        w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
        w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
    }
    // output k: bytes x+k-3 .. x+k+3; taps {18,34,48,56 | 48,34,18,-} packed little-endian
    const uint32_t kLo = 0x38302212u, kHi = 0x00122230u;
    h[0] = __dp4a(__funnelshift_r(w0, w1, 8), kLo, __dp4a(__funnelshift_r(w1, w2, 8), kHi, 0u));
    h[1] = __dp4a(__funnelshift_r(w0, w1, 16), kLo, __dp4a(__funnelshift_r(w1, w2, 16), kHi, 0u));
    h[2] = __dp4a(__funnelshift_r(w0, w1, 24), kLo, __dp4a(__funnelshift_r(w1, w2, 24), kHi, 0u));
    h[3] = __dp4a(w1, kLo, __dp4a(w2, kHi, 0u));
}

template <bool kEdge>
__device__ __forceinline__ void blur_strip(const uint8_t *__restrict__ src, int spitch, uint8_t *__restrict__ dst,
                                           int dpitch, int w, int h, int x, int y0, int rows) {
    uint32_t hw[7][4];
#pragma unroll
    for (int r = 0; r < 6; ++r)
        blur_load_row<kEdge>(src + (int64_t)reflect101(min(y0 - 3 + r, h + 2), h) * spitch, x, w, hw[r]);
    const int yend = min(y0 + rows, h);
    for (int base = 0; base < rows; base += 7) {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int y = y0 + base + k;
            if (y < yend) {
                blur_load_row<kEdge>(src + (int64_t)reflect101(min(y + 3, h + 2), h) * spitch, x, w, hw[(k + 6) % 7]);
                uint32_t v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    v[c] = 18u * (hw[k % 7][c] + hw[(k + 6) % 7][c]) + 34u * (hw[(k + 1) % 7][c] + hw[(k + 5) % 7][c]) +
                           48u * (hw[(k + 2) % 7][c] + hw[(k + 4) % 7][c]) + (56u * hw[(k + 3) % 7][c] + 32768u);
                // (v + 32768) >> 16 is byte 2 of each sum
                const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0062), __byte_perm(v[2], v[3], 0x0062), 0x5410);
                *reinterpret_cast<uint32_t *>(dst + (int64_t)y * dpitch + x) = out;
            }
        }
    }
}

// One blur work block (blockDim.x threads) of frame `frame`: block `blur_block` of the per-level block table.
__device__ __forceinline__ void blur_block_body(const FrameGeom &g, const BlurLevels &bl, const uint8_t *__restrict__ lvl0_base,
                                                int lvl0_pitch, int64_t lvl0_stride, const uint8_t *__restrict__ pyr,
                                                uint8_t *__restrict__ blur, int blur_block, int frame) {
    int level = 0;
    while (level + 1 < bl.nlevels && blur_block >= bl.block_begin[level + 1]) ++level;
    const LevelGeom &L = g.lv[level];
        const uint8_t *src;
    int spitch;
    if (level == 0) { src = lvl0_base + (int64_t)frame * lvl0_stride; spitch = lvl0_pitch; }
    else { src = pyr + L.plane_offset + (int64_t)frame * L.plane_stride; spitch = L.pitch; }
    uint8_t *dst = blur + L.plane_offset + (int64_t)frame * L.plane_stride;

    const int item = (blur_block - bl.block_begin[level]) * (int)blockDim.x + threadIdx.x;   // bl was built for this block size
    const int n_int = bl.n_int_cg[level], n_edge = bl.n_edge_cg[level];
    const int items_int = bl.n_strips[level] * n_int;
    if (item < items_int) {
        const int strip = item / n_int, cg = 1 + (item - strip * n_int);
        blur_strip<false>(src, spitch, dst, L.pitch, L.w, L.h, 4 * cg, strip * bl.rows, bl.rows);
    } else {
        const int e = item - items_int;
        if (e >= bl.n_strips[level] * n_edge) return;
        const int strip = e / n_edge, k = e - strip * n_edge;
        const int cg = k == 0 ? 0 : n_int + k;
        blur_strip<true>(src, spitch, dst, L.pitch, L.w, L.h, 4 * cg, strip * bl.rows, bl.rows);
    }
}

// Block table of one launch (host side).
// Few frames in flight: the GPU is not full and a strip is a serial chain of row loads, so strips are kept short (more,
// shorter threads); large batches amortise the 6-row prologue of a strip over 32 rows.
inline BlurLevels make_blur_levels(const FrameGeom &g, int nframes, int threads = kBlurThreads) {
    BlurLevels bl;
    bl.nlevels = g.nlevels;
    bl.rows = nframes <= 8 ? kBlurRowsSmallBatch : kBlurRows;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const int w = g.lv[l].w, h = g.lv[l].h;
        const int ncg = (w + 3) / 4;
        // interior groups: x >= 4 and x + 7 <= w - 1 (the three aligned words lie inside the row)
        int n_int = 0;
        for (int cg = 1; cg < ncg; ++cg)
            if (4 * cg + 7 <= w - 1) n_int = cg;
        bl.block_begin[l] = total;
        bl.n_int_cg[l] = n_int;
        bl.n_edge_cg[l] = ncg - n_int;
        bl.n_strips[l] = (h + bl.rows - 1) / bl.rows;
        total += (bl.n_strips[l] * ncg + threads - 1) / threads;
    }
    bl.block_begin[g.nlevels] = total;
    return bl;
}

}  // namespace vsg
