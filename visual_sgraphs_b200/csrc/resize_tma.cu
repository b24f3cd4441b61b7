// resize_tma.cu — one pyramid level for a large batch, source tiles staged through shared memory by the TMA unit.
//
// Reference (snt-arg/visual_sgraphs): level l = cv::resize(level l-1, sz_l, INTER_LINEAR), ORBextractor::ComputePyramid,
// orb_slam3/src/ORBextractor.cc:1171-1195; OpenCV's 8-bit fixed-point path:
//   h = S[sx0]*a0 + S[sx1]*a1 (11-bit coefficients);  out = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2.
// resize_kernel (pyramid.cu) interpolates every source row once per OUTPUT row that reads it (2 per output row, 1.2 new ones)
// behind its own global loads.  Here a CTA owns a 128 x 32 output tile:
//   1. one thread asks the TMA unit for the tile's source window (cp.async.bulk.tensor.3d: 192 bytes x 48 rows of the
//      level below, coordinates (16-byte aligned column, first source row, frame), out-of-plane bytes zero-filled);
//   2. horizontal pass: every source row of the window is interpolated ONCE for the tile's 128 columns
//      (warp = row, lane = group of 4 columns: three shared-memory words, byte permutes, IDP2A) into a 32-bit plane;
//   3. vertical pass: warp = output row, lane = 4 columns: two 16-byte reads of that plane, the two multiply-highs per pixel,
//      one 32-bit store.
// Same tables, same arithmetic, bit-identical output; chosen by launch_resize_level for batches the tensor maps can address.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "tma_util.cuh"
#include "vsg_internal.cuh"

namespace vsg {

constexpr int kRtW = 128, kRtH = 32;             // output tile
constexpr int kRtBoxW = 192, kRtBoxH = 48;       // source window (bytes x rows)
constexpr int kRtWarps = 8;                      // consumer warps; one more warp produces
constexpr int kRtThreads = 32 * (kRtWarps + 1);
constexpr int kRtHPitch = kRtW + 4;              // words per row of the horizontal plane (+4: rows start 16 bytes apart in the banks)
constexpr int kRtWin = kRtBoxW * kRtBoxH;
constexpr int kRtSmem = 2 * kRtWin + kRtBoxH * kRtHPitch * 4 + 128;

struct RtTileInfo {
    int x0, y0, frame, xs, ys, nrows, ylast, pad;
    short4 yt[kRtH];                             // table entries of the tile's output rows
};

// Persistent CTAs: CTA b owns the contiguous tile range [b * per, (b + 1) * per) of the (frame, tile column, tile row) order —
// row fastest, so consecutive tiles share their column setup.  Warp 8 is the producer: per tile it looks up the window origin
// and the row table, waits for a free window buffer and issues the TMA load; warps 0-7 consume.
__global__ void __launch_bounds__(kRtThreads, 4) resize_tma_kernel(const __grid_constant__ CUtensorMap src_map, uint8_t *__restrict__ dst,
                                                                   int dst_pitch, int64_t dst_stride, int dw, int dh, int tiles_x,
                                                                   int tiles_y, int total_tiles, int per_cta,
                                                                   const short4 *__restrict__ xtab, const short4 *__restrict__ ytab) {
    extern __shared__ uint8_t rt_smem_raw[];
    uint8_t *win = reinterpret_cast<uint8_t *>(((uintptr_t)rt_smem_raw + 127) & ~(uintptr_t)127);
    uint32_t *hp = reinterpret_cast<uint32_t *>(win + 2 * kRtWin);
    __shared__ uint64_t full[2], empty[2];
    __shared__ RtTileInfo info[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t_begin = blockIdx.x * per_cta, t_end = min(t_begin + per_cta, total_tiles);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kRtWarps); }
        mbar_init_fence();
    }
    __syncthreads();
    const int tiles_per_frame = tiles_x * tiles_y;

    if (warp == kRtWarps) {
        // ===== producer =====
        for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
            const int st = it & 1;
            const int frame = t / tiles_per_frame, rem = t - frame * tiles_per_frame;
            const int tx = rem / tiles_y, ty = rem - tx * tiles_y;
            const int x0 = tx * kRtW, y0 = ty * kRtH, ylast = min(y0 + kRtH, dh) - 1;
            const short4 ytl = __ldg(&ytab[min(y0 + lane, dh - 1)]);           // kRtH == 32 rows, one per lane
            const int xs = __ldg(&xtab[x0]).x & ~15;
            const int ys = __shfl_sync(0xffffffffu, (int)ytl.x, 0);
            const int nrows = __shfl_sync(0xffffffffu, (int)ytl.y, ylast - y0) - ys + 1;
            mbar_wait(&empty[st], ((it >> 1) & 1) ^ 1);                        // the consumers are done with this buffer
            info[st].yt[lane] = ytl;
            if (lane == 0) {
                info[st].x0 = x0; info[st].y0 = y0; info[st].frame = frame; info[st].xs = xs; info[st].ys = ys;
                info[st].nrows = nrows; info[st].ylast = ylast;
            }
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(&full[st], kRtWin);
                tma_load_3d(win + st * kRtWin, &src_map, xs, ys, frame, &full[st]);
            }
        }
        return;
    }

    // ===== consumers =====
    int cur_x0 = -1;
    bool col_ok = false;
    uint32_t sel[4], coef[4];
    bool hi[4];
    int base = 0;
    for (int t = t_begin, it = 0; t < t_end; ++t, ++it) {
        const int st = it & 1;
        mbar_wait(&full[st], (it >> 1) & 1);
        const int x0 = info[st].x0, y0 = info[st].y0, frame = info[st].frame, xs = info[st].xs, ys = info[st].ys, nrows = info[st].nrows,
                  ylast = info[st].ylast;
        const int x4 = x0 + 4 * lane;
        if (x0 != cur_x0) {      // column setup (as in resize_kernel): aligned base word, selectors, packed coefficients
            cur_x0 = x0;
            col_ok = x4 < dw;
            short4 xt[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) xt[k] = col_ok ? __ldg(&xtab[x4 + k]) : make_short4(0, 0, 0, 0);   // table padded to a multiple of 4
            base = xt[0].x & ~3;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o0 = xt[k].x - base, o1 = xt[k].y - base;      // 0..11
                hi[k] = o1 >= 8;
                const int f = hi[k] ? 4 : 0;
                sel[k] = (uint32_t)(o0 - f) | ((uint32_t)(o1 - f) << 4);
                coef[k] = (uint32_t)(uint16_t)xt[k].z | ((uint32_t)(uint16_t)xt[k].w << 16);
            }
        }
        const uint8_t *wcol = win + st * kRtWin + (col_ok ? base - xs : 0);
        // horizontal pass: warp = source row, lane = 4 columns
        for (int r = warp; r < nrows; r += kRtWarps) {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(wcol + r * kRtBoxW);
            const uint32_t w0 = p[0], w1 = p[1], w2 = p[2];
            uint4 g;
            uint32_t *gp = &g.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t bytes = hi[k] ? __byte_perm(w1, w2, sel[k]) : __byte_perm(w0, w1, sel[k]);
                gp[k] = __dp2a_lo(coef[k], bytes, 0u) >> 4;
            }
            *reinterpret_cast<uint4 *>(hp + r * kRtHPitch + 4 * lane) = g;
        }
        short4 yt[kRtH / kRtWarps];
#pragma unroll
        for (int i = 0; i < kRtH / kRtWarps; ++i) yt[i] = info[st].yt[warp + kRtWarps * i];
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);                  // window and tile record are free for the producer
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kRtWarps) : "memory");
        // vertical pass: warp = output row, lane = 4 columns
        uint8_t *d = dst + (int64_t)frame * dst_stride + x4;
#pragma unroll
        for (int i = 0; i < kRtH / kRtWarps; ++i) {
            const int y = y0 + warp + kRtWarps * i;
            if (y > ylast || !col_ok) continue;
            const uint4 ga = *reinterpret_cast<const uint4 *>(hp + (yt[i].x - ys) * kRtHPitch + 4 * lane);
            const uint4 gb = *reinterpret_cast<const uint4 *>(hp + (yt[i].y - ys) * kRtHPitch + 4 * lane);
            const uint32_t b0 = (uint32_t)yt[i].z << 16, b1 = (uint32_t)yt[i].w << 16;
            const uint32_t v0 = (__umulhi(b0, ga.x) + __umulhi(b1, gb.x) + 2u) >> 2, v1 = (__umulhi(b0, ga.y) + __umulhi(b1, gb.y) + 2u) >> 2;
            const uint32_t v2 = (__umulhi(b0, ga.z) + __umulhi(b1, gb.z) + 2u) >> 2, v3 = (__umulhi(b0, ga.w) + __umulhi(b1, gb.w) + 2u) >> 2;
            *reinterpret_cast<uint32_t *>(d + (int64_t)y * dst_pitch) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kRtWarps) : "memory");   // the horizontal plane is free for the next tile
    }
}

// VSG_RESIZE_TMA = n: batches of at least n frames resize through the TMA-staged kernel (0 = never)
static int resize_tma_min_frames() {
    const char *e = getenv("VSG_RESIZE_TMA");
    const int v = e ? atoi(e) : 16;
    return v <= 0 ? INT32_MAX : v;
}

// true: launched.  false: the level does not fit this kernel (small batch, unaligned caller-owned source, a scale factor whose
// tile window exceeds the box) — the caller launches resize_kernel.
bool launch_resize_level_tma(const FrameGeom &g, int level, const uint8_t *src_base, int src_pitch, int64_t src_stride, uint8_t *pyr,
                             int nframes, cudaStream_t s) {
    if (nframes < resize_tma_min_frames()) return false;
    const LevelGeom &L = g.lv[level], &P = g.lv[level - 1];
    if (!L.resize_tma_ok || ((uintptr_t)src_base & 15) || (src_pitch & 15) || (src_stride & 15) || src_pitch < kRtBoxW) return false;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)src_pitch, (cuuint64_t)P.h, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)src_pitch, (cuuint64_t)src_stride};
    const cuuint32_t box[3] = {kRtBoxW, kRtBoxH, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)src_base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cudaFuncSetAttribute(resize_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRtSmem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    const int tiles_x = (L.w + kRtW - 1) / kRtW, tiles_y = (L.h + kRtH - 1) / kRtH;
    const int64_t total = (int64_t)tiles_x * tiles_y * nframes;
    if (total > INT32_MAX) return false;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ctas = (int)std::min<int64_t>(total, (int64_t)sms * 4);
    const int per_cta = (int)((total + ctas - 1) / ctas);
    resize_tma_kernel<<<(unsigned)((total + per_cta - 1) / per_cta), kRtThreads, kRtSmem, s>>>(
        map, pyr + L.plane_offset, L.pitch, L.plane_stride, L.w, L.h, tiles_x, tiles_y, (int)total, per_cta, L.xtab, L.ytab);
    count_launch();
    return true;
}

// host side of LevelGeom::resize_tma_ok: every tile's source window fits the box
bool resize_tma_fits(const std::vector<short4> &xt, const std::vector<short4> &yt, int dw, int dh) {
    for (int x0 = 0; x0 < dw; x0 += kRtW) {
        const int xl = std::min(x0 + kRtW, dw) - 1;
        if ((xt[xl].y & ~3) + 12 - (xt[x0].x & ~15) > kRtBoxW) return false;   // the last column group reads three whole words
    }
    for (int y0 = 0; y0 < dh; y0 += kRtH) {
        const int yl = std::min(y0 + kRtH, dh) - 1;
        if (yt[yl].y - yt[y0].x + 1 > kRtBoxH) return false;
    }
    return true;
}

}  // namespace vsg
