// pyramid.cu — ComputePyramid and the per-level Gaussian blur on the device.
//
// Reference (snt-arg/visual_sgraphs):
//   ORBextractor::ComputePyramid            orb_slam3/src/ORBextractor.cc:1171-1195
//     level l = cv::resize(level l-1, sz_l, INTER_LINEAR)   (:1184)   -> resize_kernel
//   GaussianBlur(clone, 7x7, sigma 2, BORDER_REFLECT_101)   (:1129-1130) -> blur_kernel
// Both are OpenCV's 8-bit fixed-point paths, reproduced bit-exactly:
//   resize:  h = S[sx0]*a0 + S[sx1]*a1 (11-bit coefs); out = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2
//   blur:    taps {18,34,48,56,48,34,18}/256 per pass, no intermediate rounding, out = (v + 32768) >> 16
// The coefficient tables (float/double arithmetic of cv::resize) are built on the host (vsg_api.cu).
#include "vsg_internal.cuh"

namespace vsg {

// ------------------------------------------------------------------------------------------------
// resize: one thread owns 4 output columns x kResizeRows output rows.  Its four
// (sx0, sx1, a0, a1) column entries are loop invariant: the source bytes of a row come from three
// aligned 32-bit loads and one byte-permute per column, the horizontal interpolation is one IDP2A
// (a0*s0 + a1*s1).  The rows of a strip are independent, so all their loads are in flight together
// (the source row shared by two consecutive output rows is simply re-read from L1).
// Vertical: ((b*(h>>4))>>16) is a 32x32 high multiply by b<<16.
// grid (ceil(ncg * nstrips / 128), 1, nframes), block 128.
// ------------------------------------------------------------------------------------------------
#ifndef VSG_RESIZE_ROWS
#define VSG_RESIZE_ROWS 4
#endif
#ifndef VSG_RESIZE_MINB
#define VSG_RESIZE_MINB 12
#endif
#ifndef VSG_BLUR_MINB
#define VSG_BLUR_MINB 10
#endif
constexpr int kResizeRows = VSG_RESIZE_ROWS, kResizeThreads = 128;

__global__ void __launch_bounds__(kResizeThreads, VSG_RESIZE_MINB) resize_kernel(const uint8_t *__restrict__ src, int src_pitch,
                                                                int64_t src_stride, uint8_t *__restrict__ dst,
                                                                int dst_pitch, int64_t dst_stride, int dw, int dh,
                                                                int sw, const uint8_t *__restrict__ src_end,
                                                                const short4 *__restrict__ xtab,
                                                                const short4 *__restrict__ ytab) {
    pdl_launch_dependents();
    pdl_wait();
    // (row strip, column group) flattened over the grid's x dimension: no idle lanes at the right edge of narrow levels
    const int ncg = (dw + 3) >> 2;
    const int idx = blockIdx.x * kResizeThreads + threadIdx.x;
    const int strip = idx / ncg;
    const int x4 = (idx - strip * ncg) * 4;
    const int y0 = strip * kResizeRows, y1 = min(y0 + kResizeRows, dh);
    if (y0 >= dh) return;
    const uint8_t *s = src + (int64_t)blockIdx.z * src_stride;
    uint8_t *d = dst + (int64_t)blockIdx.z * dst_stride + x4;

    // column setup: aligned base word, byte selectors relative to it, packed coefficients
    short4 xt[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xt[k] = __ldg(&xtab[x4 + k]);   // table is padded to a multiple of 4 entries
    const int base = xt[0].x & ~3;                               // all 8 source columns lie in [base, base + 12)
    uint32_t sel[4], coef[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int o0 = xt[k].x - base, o1 = xt[k].y - base;      // 0..11
        sel[k] = (uint32_t)o0 | ((uint32_t)o1 << 4);             // resolved against (w0,w1) or (w1,w2) below
        coef[k] = (uint32_t)(uint16_t)xt[k].z | ((uint32_t)(uint16_t)xt[k].w << 16);
    }
    // bytes 0..7 come from (w0, w1), 4..11 from (w1, w2): pick per column which pair holds both taps
    bool hi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int o0 = sel[k] & 15, o1 = sel[k] >> 4;
        hi[k] = o1 >= 8;
        const int f = hi[k] ? 4 : 0;
        sel[k] = (uint32_t)(o0 - f) | ((uint32_t)(o1 - f) << 4);
    }
    const bool need_w2 = hi[0] || hi[1] || hi[2] || hi[3];       // false only when all taps sit in the first 8 bytes

    // Words past the end of a source row only ever hold unused bytes (taps are clamped to sw-1), and reading
    // them is harmless except at the very end of the source buffer (level 0 may be a caller-owned buffer with
    // nothing behind its last row): only there the 12 bytes are assembled one by one.
    auto hrow = [&](int sy, uint32_t g[4]) {                    // g = (S[sx0]*a0 + S[sx1]*a1) >> 4 for the 4 columns
        const uint8_t *row = s + (int64_t)sy * src_pitch;
        uint32_t w0, w1, w2 = 0u;
        if (row + base + 12 <= src_end) {
            const uint32_t *p = reinterpret_cast<const uint32_t *>(row + base);
            w0 = __ldg(p); w1 = __ldg(p + 1);
            if (need_w2) w2 = __ldg(p + 2);
        } else {
            uint32_t b[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) b[k] = __ldg(row + min(base + k, sw - 1));
            w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
            w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
            w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t bytes = hi[k] ? __byte_perm(w1, w2, sel[k]) : __byte_perm(w0, w1, sel[k]);
            g[k] = __dp2a_lo(coef[k], bytes, 0u) >> 4;
        }
    };

    // kResizeRows independent output rows: all table and source loads are issued before any arithmetic
    short4 yt[kResizeRows];
#pragma unroll
    for (int r = 0; r < kResizeRows; ++r) yt[r] = __ldg(&ytab[min(y0 + r, dh - 1)]);
    uint32_t ga[kResizeRows][4], gb[kResizeRows][4];
#pragma unroll
    for (int r = 0; r < kResizeRows; ++r) {
        hrow(yt[r].x, ga[r]);
        hrow(yt[r].y, gb[r]);
    }
#pragma unroll
    for (int r = 0; r < kResizeRows; ++r) {
        if (y0 + r >= y1) break;
        const uint32_t b0 = (uint32_t)yt[r].z << 16, b1 = (uint32_t)yt[r].w << 16;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (__umulhi(b0, ga[r][k]) + __umulhi(b1, gb[r][k]) + 2u) >> 2;
        *reinterpret_cast<uint32_t *>(d + (int64_t)(y0 + r) * dst_pitch) = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
    }
}

// ------------------------------------------------------------------------------------------------
// The same pyramid for a few frames in one launch (single-frame latency: seven dependent launches cost 33 us, of which the
// arithmetic is a few).  CTA = (tile, frame).  Level l of a tile is computed from level l-1 of the SAME tile in shared
// memory, so no CTA ever waits for another: the region a tile computes at level l-1 is the bounding box of the taps its
// level-l region reads and of the box it owns there; it stores only what it owns.  Halo pixels are computed by two or
// more tiles from the same inputs with the same formula, i.e. identically.  Arithmetic as in resize_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kTileThreads = 256;
#ifdef PT_TRACE
__device__ long long g_pt_trace[4][20];
#define PT(i) do { if (tid == 0 && blockIdx.x % 48 == 0 && blockIdx.x / 48 < 4) g_pt_trace[blockIdx.x / 48][i] = clock64(); } while (0)
extern "C" int vsg_debug_pt_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_pt_trace, sizeof(g_pt_trace)); }
#else
#define PT(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kTileThreads) pyramid_tile_kernel(FrameGeom g, const PyrTile *__restrict__ tiles,
                                                                    const uint8_t *__restrict__ lvl0_base, int lvl0_pitch,
                                                                    int64_t lvl0_stride, uint8_t *__restrict__ pyr, int buf_bytes) {
    extern __shared__ __align__(16) uint8_t tile_smem[];
    __shared__ PyrTile T;
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, frame = blockIdx.y;
    PT(0);
    if (tid < (int)(sizeof(PyrTile) / 4))
        reinterpret_cast<uint32_t *>(&T)[tid] = __ldg(reinterpret_cast<const uint32_t *>(tiles + blockIdx.x) + tid);
    __syncthreads();
    PT(1);
    short4 *tab = reinterpret_cast<short4 *>(tile_smem + 2 * buf_bytes);
    // every table entry the tile will use (x entries of a level, then its y entries) and the level-0 region in ONE round
    // trip: all loads are issued before the first store (a region is at most kTileThreads wide and high — checked on the host)
    {
        short4 vx[kMaxLevels - 1], vy[kMaxLevels - 1];
#pragma unroll
        for (int l = 1; l < kMaxLevels; ++l) {
            if (l < g.nlevels) {
                const PyrTileBox R = T.region[l];
                if (tid <= R.x1 - R.x0) vx[l - 1] = __ldg(&g.lv[l].xtab[R.x0 + tid]);
                if (tid <= R.y1 - R.y0) vy[l - 1] = __ldg(&g.lv[l].ytab[R.y0 + tid]);
            }
        }
        const PyrTileBox R0 = T.region[0];
        const int rw0 = R0.x1 - R0.x0 + 1, area0 = rw0 * (R0.y1 - R0.y0 + 1);
        const uint32_t rcp0 = T.rcp_w[0];
        const uint8_t *src = lvl0_base + (int64_t)frame * lvl0_stride + (int64_t)R0.y0 * lvl0_pitch + R0.x0;
        for (int base = 0; base < area0; base += 16 * kTileThreads) {       // 16 loads in flight per thread
            uint8_t px[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int i = base + k * kTileThreads + tid;
                const int r = (int)__umulhi((uint32_t)i, rcp0);
                px[k] = i < area0 ? __ldg(src + (int64_t)r * lvl0_pitch + (i - r * rw0)) : 0;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int i = base + k * kTileThreads + tid;
                if (i < area0) tile_smem[i] = px[k];
            }
        }
        int off = 0;
#pragma unroll
        for (int l = 1; l < kMaxLevels; ++l) {
            if (l < g.nlevels) {
                const PyrTileBox R = T.region[l];
                const int rw = R.x1 - R.x0 + 1, rh = R.y1 - R.y0 + 1;
                if (tid < rw) tab[off + tid] = vx[l - 1];
                if (tid < rh) tab[off + rw + tid] = vy[l - 1];
                off += rw + rh;
            }
        }
    }
    __syncthreads();
    PT(2);
    int off = 0;
    for (int l = 1; l < g.nlevels; ++l) {
        const PyrTileBox S = T.region[l - 1], R = T.region[l], O = T.owned[l];
        const int sw = S.x1 - S.x0 + 1, rw = R.x1 - R.x0 + 1, rh = R.y1 - R.y0 + 1;
        const uint8_t *__restrict__ src = tile_smem + ((l - 1) & 1) * buf_bytes - (S.y0 * sw + S.x0);   // indexable by plane coordinates
        uint8_t *__restrict__ dst = tile_smem + (l & 1) * buf_bytes;
        const short4 *__restrict__ tx = tab + off, *__restrict__ ty = tx + rw;
        const LevelGeom &L = g.lv[l];
        uint8_t *out = pyr + L.plane_offset + (int64_t)frame * L.plane_stride;
        const uint32_t rcp = T.rcp_w[l];
        const int area = rw * rh;
        for (int base = 0; base < area; base += 4 * kTileThreads) {         // four independent pixels per thread and pass
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = base + k * kTileThreads + tid;
                if (i < area) {
                    const int r = (int)__umulhi((uint32_t)i, rcp), c = i - r * rw;
                    const short4 xt = tx[c], yt = ty[r];
                    const uint8_t *s0 = src + yt.x * sw, *s1 = src + yt.y * sw;
                    const uint32_t ga = (uint32_t)(s0[xt.x] * xt.z + s0[xt.y] * xt.w) >> 4;
                    const uint32_t gb = (uint32_t)(s1[xt.x] * xt.z + s1[xt.y] * xt.w) >> 4;
                    const uint32_t v = (__umulhi((uint32_t)yt.z << 16, ga) + __umulhi((uint32_t)yt.w << 16, gb) + 2u) >> 2;
                    dst[i] = (uint8_t)v;
                    const int x = R.x0 + c, y = R.y0 + r;
                    if (x >= O.x0 && x <= O.x1 && y >= O.y0 && y <= O.y1) out[(int64_t)y * L.pitch + x] = (uint8_t)v;
                }
            }
        }
        off += rw + rh;
        __syncthreads();
        PT(2 + l);
    }
}

void launch_pyramid_tiles(const FrameGeom &g, const PyrTile *tiles, int ntiles, int buf_bytes, size_t smem, const uint8_t *lvl0_base,
                          int lvl0_pitch, int64_t lvl0_stride, uint8_t *pyr, int nframes, cudaStream_t s) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(pyramid_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_kernel(pyramid_tile_kernel, dim3(ntiles, nframes), dim3(kTileThreads), smem, s, true, g, tiles, lvl0_base, lvl0_pitch,
                  lvl0_stride, pyr, buf_bytes);
    count_launch();
}

void launch_resize_level(const FrameGeom &g, int level, const uint8_t *src_base, int src_pitch, int64_t src_stride,
                         uint8_t *pyr, int nframes, cudaStream_t s) {
    if (launch_resize_level_tma(g, level, src_base, src_pitch, src_stride, pyr, nframes, s)) return;
    const LevelGeom &P = g.lv[level - 1];
    // one past the last byte of the source batch (our own planes have slack behind them, but stay exact)
    const uint8_t *src_end = src_base + (int64_t)(nframes - 1) * src_stride + (int64_t)(P.h - 1) * src_pitch + P.w;
    const LevelGeom &L = g.lv[level];
    const int ncg = (L.w + 3) / 4;
    const int nstrips = (L.h + kResizeRows - 1) / kResizeRows;
    dim3 grid((ncg * nstrips + kResizeThreads - 1) / kResizeThreads, 1, nframes);
    launch_kernel(resize_kernel, grid, dim3(kResizeThreads), 0, s, true, src_base, src_pitch, src_stride, pyr + L.plane_offset,
                  L.pitch, L.plane_stride, L.w, L.h, P.w, src_end, L.xtab, L.ytab);
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// colour ingest: cv::cvtColor(RGB|BGR|RGBA|BGRA -> GRAY) as Tracking::GrabImageRGBD / GrabImageMonocular /
// GrabImageStereo run it ahead of the extractor (orb_slam3/src/Tracking.cc:1526-1551, 1595-1608, 1646-1660).
// OpenCV's 8-bit path: gray = (R*9798 + G*19235 + B*3735 + (1 << 14)) >> 15 (verified bit-exact against cv2 4.13).
// One thread converts 4 pixels: 3 (or 4) aligned 32-bit loads, one 32-bit store into the level-0 plane.
// ------------------------------------------------------------------------------------------------
template <int kChannels>
__global__ void __launch_bounds__(128) cvt_gray_kernel(const uint8_t *__restrict__ src, int src_pitch, int64_t src_stride,
                                                       uint8_t *__restrict__ dst, int dst_pitch, int64_t dst_stride, int w,
                                                       int h, int r_first) {
    const int x4 = (blockIdx.x * 128 + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= w) return;
    const uint8_t *row = src + (int64_t)blockIdx.z * src_stride + (int64_t)y * src_pitch + (int64_t)x4 * kChannels;
    uint8_t px[4][4];
    if (x4 + 4 <= w) {
        uint32_t words[kChannels];
#pragma unroll
        for (int k = 0; k < kChannels; ++k) words[k] = __ldg(reinterpret_cast<const uint32_t *>(row) + k);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int byte = i * kChannels + c;
                px[i][c] = (uint8_t)(words[byte >> 2] >> (8 * (byte & 3)));
            }
    } else {
        for (int i = 0; i < 4; ++i)
            for (int c = 0; c < 3; ++c) px[i][c] = x4 + i < w ? __ldg(row + i * kChannels + c) : 0;
    }
    const uint32_t c0 = r_first ? 9798u : 3735u, c2 = r_first ? 3735u : 9798u;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) out |= ((px[i][0] * c0 + px[i][1] * 19235u + px[i][2] * c2 + (1u << 14)) >> 15) << (8 * i);
    uint8_t *d = dst + (int64_t)blockIdx.z * dst_stride + (int64_t)y * dst_pitch + x4;
    if (x4 + 4 <= w) *reinterpret_cast<uint32_t *>(d) = out;   // level-0 planes have a pitch that is a multiple of 64
    else for (int i = 0; x4 + i < w; ++i) d[i] = (uint8_t)(out >> (8 * i));
}

// ------------------------------------------------------------------------------------------------
// rectification ingest: cv::remap(src, dst, M1, M2, INTER_LINEAR) as System::TrackStereo / TrackMonocular run it on
// every incoming image when the settings ask for it (orb_slam3/src/System.cc:284-292; float maps from
// cv::initUndistortRectifyMap, Settings.cc:571-574).  OpenCV's 8-bit path: map quantised to 1/32 pixel (done once on
// the host when the map is set: integer part as two shorts, fraction as fy * 32 + fx), weights (32-fy)(32-fx)*32 ...
// summing to 2^15, dst = (sum + 2^14) >> 15, taps outside the source are 0 (BORDER_CONSTANT).  One thread = 4 output
// pixels, one 32-bit store into the level-0 plane; the map of a camera is shared by every frame of the batch (L2).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) remap_kernel(const uint8_t *__restrict__ src, int src_pitch, int64_t src_stride, int sw,
                                                    int sh, RectifyMaps maps, int first_frame, uint8_t *__restrict__ dst,
                                                    int dst_pitch, int64_t dst_stride, int w, int h) {
    const int x4 = (blockIdx.x * 128 + threadIdx.x) * 4, y = blockIdx.y;
    if (x4 >= w) return;
    const int slot = (first_frame + (int)blockIdx.z) % maps.nslots;
    const uint32_t *mxy = maps.xy[slot] + (size_t)y * w + x4;
    const uint16_t *mf = maps.frac[slot] + (size_t)y * w + x4;
    const uint8_t *s = src + (int64_t)blockIdx.z * src_stride;
    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (x4 + i >= w) break;
        const uint32_t xy = __ldg(mxy + i);
        const int f = __ldg(mf + i);
        const int ix = (short)(xy & 0xFFFF), iy = (short)(xy >> 16), fx = f & 31, fy = f >> 5;
        const uint8_t *p = s + (int64_t)iy * src_pitch + ix;
        const bool x0 = (unsigned)ix < (unsigned)sw, x1 = (unsigned)(ix + 1) < (unsigned)sw;
        const bool y0 = (unsigned)iy < (unsigned)sh, y1 = (unsigned)(iy + 1) < (unsigned)sh;
        const int t00 = (x0 && y0) ? __ldg(p) : 0, t01 = (x1 && y0) ? __ldg(p + 1) : 0;
        const int t10 = (x0 && y1) ? __ldg(p + src_pitch) : 0, t11 = (x1 && y1) ? __ldg(p + src_pitch + 1) : 0;
        const int v = t00 * ((32 - fy) * (32 - fx) * 32) + t01 * ((32 - fy) * fx * 32) + t10 * (fy * (32 - fx) * 32) + t11 * (fy * fx * 32);
        out |= (uint32_t)((v + (1 << 14)) >> 15) << (8 * i);
    }
    uint8_t *d = dst + (int64_t)blockIdx.z * dst_stride + (int64_t)y * dst_pitch + x4;
    if (x4 + 4 <= w) *reinterpret_cast<uint32_t *>(d) = out;   // level-0 planes have a pitch that is a multiple of 64
    else for (int i = 0; x4 + i < w; ++i) d[i] = (uint8_t)(out >> (8 * i));
}

void launch_remap(const uint8_t *src, int src_pitch, int64_t src_stride, int sw, int sh, const RectifyMaps &maps, int first_frame,
                  uint8_t *dst, int dst_pitch, int64_t dst_stride, int w, int h, int nframes, cudaStream_t s) {
    dim3 grid(((w + 3) / 4 + 127) / 128, h, nframes);
    remap_kernel<<<grid, 128, 0, s>>>(src, src_pitch, src_stride, sw, sh, maps, first_frame, dst, dst_pitch, dst_stride, w, h);
    count_launch();
}

void launch_cvt_gray(const uint8_t *src, int src_pitch, int64_t src_stride, int channels, int r_first, uint8_t *dst,
                     int dst_pitch, int64_t dst_stride, int w, int h, int nframes, cudaStream_t s) {
    dim3 grid(((w + 3) / 4 + 127) / 128, h, nframes);
    if (channels == 3) cvt_gray_kernel<3><<<grid, 128, 0, s>>>(src, src_pitch, src_stride, dst, dst_pitch, dst_stride, w, h, r_first);
    else cvt_gray_kernel<4><<<grid, 128, 0, s>>>(src, src_pitch, src_stride, dst, dst_pitch, dst_stride, w, h, r_first);
    count_launch();
}

}  // namespace vsg

#include "blur_device.cuh"

namespace vsg {

__global__ void __launch_bounds__(kBlurThreads, VSG_BLUR_MINB) blur_kernel(FrameGeom g, BlurLevels bl, const uint8_t *__restrict__ lvl0_base,
                                                            int lvl0_pitch, int64_t lvl0_stride,
                                                            const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur) {
    pdl_launch_dependents();
    pdl_wait();
    blur_block_body(g, bl, lvl0_base, lvl0_pitch, lvl0_stride, pyr, blur, blockIdx.x, blockIdx.y);
}

void launch_blur(const FrameGeom &g, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride, const uint8_t *pyr,
                 uint8_t *blur, int nframes, cudaStream_t s) {
    const BlurLevels bl = make_blur_levels(g, nframes);
    launch_kernel(blur_kernel, dim3(bl.block_begin[g.nlevels], nframes), dim3(kBlurThreads), 0, s, true, g, bl, lvl0_base,
                  lvl0_pitch, lvl0_stride, pyr, blur);
    count_launch();
}

}  // namespace vsg
