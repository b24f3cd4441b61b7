// fast.cu — per-cell FAST-9/16 with cell-local non-max suppression and the iniThFAST/minThFAST retry.
//
// Reference (snt-arg/visual_sgraphs):
//   ORBextractor::ComputeKeyPointsOctTree, cell loop      orb_slam3/src/ORBextractor.cc:811-876
//     cv::FAST(cell, kps, iniThFAST, true); if empty -> cv::FAST(cell, kps, minThFAST, true)   (:832-851)
//   cv::FAST = FAST-9/16 + 3x3 NMS on the cell sub-image:
//     corner at threshold t  <=>  strength K > t, where K = max over the sixteen 9-arcs of the
//     smallest same-signed difference on the arc; response = K - 1; a corner survives NMS iff its
//     response is strictly greater than the responses of its 8 neighbours, where non-corners and
//     pixels outside the cell's interior ([3,w-3) x [3,h-3)) count as 0.
//
// One CTA per (cell, frame).  The cell window is staged in shared memory with aligned 32-bit loads,
// the strength K of every interior pixel is computed once (it does not depend on the threshold), NMS
// runs at iniThFAST and — only if the cell produced nothing — again at minThFAST, exactly the
// reference's retry rule.  Survivors are appended to the (frame, level) candidate list with one
// global atomic per CTA.  The list order is not the reference's cell-row-major order; the oct-tree
// only depends on order through the first-max-wins tie break, which octree.cu reproduces from the
// coordinates (see order_key there).
#include "vsg_internal.cuh"

namespace vsg {

// The 16-pixel Bresenham ring is read in OpenCV's order: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)
// (1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3).
// Strength K of the pixel at c (tile pitch tp): max over 16 circular 9-windows of the window minimum
// of the ring (bright arcs) and of the negated window maximum (dark arcs), relative to the centre.
__device__ __forceinline__ int fast_strength(const uint8_t *c, int tp) {
    int p[16];
    p[0] = c[3 * tp];      p[1] = c[3 * tp + 1];   p[2] = c[2 * tp + 2];   p[3] = c[tp + 3];
    p[4] = c[3];           p[5] = c[-tp + 3];      p[6] = c[-2 * tp + 2];  p[7] = c[-3 * tp + 1];
    p[8] = c[-3 * tp];     p[9] = c[-3 * tp - 1];  p[10] = c[-2 * tp - 2]; p[11] = c[-tp - 3];
    p[12] = c[-3];         p[13] = c[tp - 3];      p[14] = c[2 * tp - 2];  p[15] = c[3 * tp - 1];
    const int v = c[0];
    // sliding-window min / max of width 9 by doubling: 2, 4, 8, then +1
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        lo2[i] = min(p[i], p[(i + 1) & 15]);
        hi2[i] = max(p[i], p[(i + 1) & 15]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        lo4[i] = min(lo2[i], lo2[(i + 2) & 15]);
        hi4[i] = max(hi2[i], hi2[(i + 2) & 15]);
    }
    int best_lo = 0, best_hi = 255;  // max over arcs of window-min, min over arcs of window-max
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int lo9 = min(min(lo4[i], lo4[(i + 4) & 15]), p[(i + 8) & 15]);
        const int hi9 = max(max(hi4[i], hi4[(i + 4) & 15]), p[(i + 8) & 15]);
        best_lo = max(best_lo, lo9);
        best_hi = min(best_hi, hi9);
    }
    return max(max(best_lo - v, v - best_hi), 0);
}

__global__ void __launch_bounds__(256) fast_kernel(FrameGeom g, const Cell *__restrict__ cells,
                                                   const uint8_t *__restrict__ lvl0_base, int lvl0_pitch,
                                                   int64_t lvl0_stride, const uint8_t *__restrict__ pyr,
                                                   Cand *__restrict__ cand, int *__restrict__ cand_count, int ini_th,
                                                   int min_th, int tile_pitch, int tile_rows, int score_pitch,
                                                   int list_cap) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *tile = smem;                                        // tile_rows x tile_pitch
    uint8_t *score = tile + tile_rows * tile_pitch;              // (ih + 2) x score_pitch, zero apron
    uint32_t *list = reinterpret_cast<uint32_t *>(score + (tile_rows - 4) * score_pitch);
    __shared__ int s_count, s_base;

    const Cell cell = cells[blockIdx.x];
    const int frame = blockIdx.y;
    const LevelGeom &L = g.lv[cell.level];
    const uint8_t *src;
    int spitch;
    if (cell.level == 0) { src = lvl0_base + (int64_t)frame * lvl0_stride; spitch = lvl0_pitch; }
    else { src = pyr + L.plane_offset + (int64_t)frame * L.plane_stride; spitch = L.pitch; }

    const int tid = threadIdx.x;
    const int ax0 = cell.x0 & ~3;                 // 4-byte aligned tile origin
    const int xoff = cell.x0 - ax0;
    const int nwords = (xoff + cell.cw + 3) >> 2;
    for (int i = tid; i < cell.ch * nwords; i += 256) {
        const int r = i / nwords, wi = i - r * nwords;
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(src + (int64_t)(cell.y0 + r) * spitch + ax0) + wi);
        *reinterpret_cast<uint32_t *>(tile + r * tile_pitch + 4 * wi) = w;
    }
    const int iw = cell.cw - 6, ih = cell.ch - 6;  // interior = FAST's [3,w-3) x [3,h-3)
    for (int i = tid; i < (ih + 2) * score_pitch; i += 256) score[i] = 0;
    if (tid == 0) s_count = 0;
    __syncthreads();

    for (int i = tid; i < iw * ih; i += 256) {
        const int iy = i / iw, ix = i - iy * iw;
        const int K = fast_strength(tile + (iy + 3) * tile_pitch + xoff + ix + 3, tile_pitch);
        score[(iy + 1) * score_pitch + ix + 1] = (uint8_t)K;
    }
    __syncthreads();

    for (int pass = 0; pass < 2; ++pass) {
        const int t = pass == 0 ? ini_th : min_th;
        for (int i = tid; i < iw * ih; i += 256) {
            const int iy = i / iw, ix = i - iy * iw;
            const uint8_t *s = score + (iy + 1) * score_pitch + ix + 1;
            const int K = s[0];
            if (K <= t) continue;
            const int resp = K - 1;
            bool is_max = true;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    if (dx == 0 && dy == 0) continue;
                    const int kn = s[dy * score_pitch + dx];
                    const int rn = kn > t ? kn - 1 : 0;
                    is_max = is_max && (resp > rn);
                }
            if (is_max) {
                const int slot = atomicAdd(&s_count, 1);
                if (slot < list_cap) list[slot] = (uint32_t)ix | ((uint32_t)iy << 8) | ((uint32_t)resp << 16);
            }
        }
        __syncthreads();
        if (s_count > 0) break;   // uniform: the retry happens only when the cell is empty (:842)
    }
    const int n = min(s_count, list_cap);
    if (n == 0) return;
    const int slot_idx = frame * g.nlevels + cell.level;
    if (tid == 0) s_base = atomicAdd(&cand_count[slot_idx], n);
    __syncthreads();
    Cand *out = cand + L.cand_offset + (int64_t)frame * g.cand_total;
    for (int i = tid; i < n; i += 256) {
        const int dst = s_base + i;
        if (dst >= L.cand_cap) break;
        const uint32_t e = list[i];
        Cand c;
        c.x = (unsigned short)(cell.x0 + 3 + (e & 0xff));
        c.y = (unsigned short)(cell.y0 + 3 + ((e >> 8) & 0xff));
        c.score = (unsigned short)(e >> 16);
        c.pad = 0;
        out[dst] = c;
    }
}

void launch_fast(const FrameGeom &g, const Cell *cells, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride,
                 const uint8_t *pyr, Cand *cand, int *cand_count, int ini_th, int min_th, int max_cw, int max_ch,
                 int nframes, cudaStream_t s) {
    if (g.ncells == 0) return;
    const int tile_pitch = ((max_cw + 3 + 3) & ~3) + 4;           // room for the alignment shift
    const int tile_rows = max_ch;
    const int score_pitch = ((max_cw - 6 + 2) + 3) & ~3;
    const int list_cap = ((max_cw - 6 + 1) / 2) * ((max_ch - 6 + 1) / 2) + 1;
    const size_t smem = (size_t)tile_rows * tile_pitch + (size_t)(tile_rows - 4) * score_pitch + (size_t)list_cap * 4 + 16;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    fast_kernel<<<dim3(g.ncells, nframes), 256, smem, s>>>(g, cells, lvl0_base, lvl0_pitch, lvl0_stride, pyr, cand,
                                                          cand_count, ini_th, min_th, tile_pitch, tile_rows,
                                                          score_pitch, list_cap);
    count_launch();
}

}  // namespace vsg
