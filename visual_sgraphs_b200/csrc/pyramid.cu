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
// resize: one thread = 4 horizontally adjacent output pixels of one row (one 32-bit store).
// grid (ceil(dw/128), ceil(dh/8), nframes), block (32, 8).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_kernel(const uint8_t *__restrict__ src, int src_pitch, int64_t src_stride,
                                                     uint8_t *__restrict__ dst, int dst_pitch, int64_t dst_stride,
                                                     int dw, int dh, const short4 *__restrict__ xtab,
                                                     const short4 *__restrict__ ytab) {
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x4 >= dw || y >= dh) return;
    const uint8_t *s = src + (int64_t)blockIdx.z * src_stride;
    uint8_t *d = dst + (int64_t)blockIdx.z * dst_stride;
    const short4 yt = __ldg(&ytab[y]);
    const uint8_t *r0 = s + (int64_t)yt.x * src_pitch;
    const uint8_t *r1 = s + (int64_t)yt.y * src_pitch;
    const int b0 = yt.z, b1 = yt.w;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const short4 xt = __ldg(&xtab[x4 + k]);  // table is padded to a multiple of 4 entries
        const int h0 = (int)__ldg(r0 + xt.x) * xt.z + (int)__ldg(r0 + xt.y) * xt.w;
        const int h1 = (int)__ldg(r1 + xt.x) * xt.z + (int)__ldg(r1 + xt.y) * xt.w;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        packed |= (uint32_t)min(max(v, 0), 255) << (8 * k);
    }
    *reinterpret_cast<uint32_t *>(d + (int64_t)y * dst_pitch + x4) = packed;  // pitch % 16 == 0: pad bytes are scratch
}

void launch_resize_level(const FrameGeom &g, int level, const uint8_t *src_base, int src_pitch, int64_t src_stride,
                         uint8_t *pyr, int nframes, cudaStream_t s) {
    const LevelGeom &L = g.lv[level];
    dim3 block(32, 8);
    dim3 grid((L.w + 127) / 128, (L.h + 7) / 8, nframes);
    resize_kernel<<<grid, block, 0, s>>>(src_base, src_pitch, src_stride, pyr + L.plane_offset, L.pitch, L.plane_stride,
                                         L.w, L.h, L.xtab, L.ytab);
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// blur: CTA = 64 x 32 output tile; source tile (70 x 38) staged in shared memory with the
// REFLECT_101 index map applied at load time (the blur reflects the level itself, not a padded
// buffer — SURVEY App. A2), horizontal pass into a u16 plane, vertical pass from it.
// One launch covers every level: blockIdx.x walks a flat tile table.
// ------------------------------------------------------------------------------------------------
constexpr int kBlurTW = 64, kBlurTH = 32;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

struct BlurLevels {
    int nlevels;
    int tile_begin[kMaxLevels + 1];  // prefix sums of tiles per level
    int tiles_x[kMaxLevels];
};

__global__ void __launch_bounds__(256) blur_kernel(FrameGeom g, BlurLevels bl, const uint8_t *__restrict__ lvl0_base,
                                                   int lvl0_pitch, int64_t lvl0_stride, const uint8_t *__restrict__ pyr,
                                                   uint8_t *__restrict__ blur) {
    __shared__ uint8_t tile[kBlurTH + 6][kBlurTW + 8];
    __shared__ uint16_t hbuf[kBlurTH + 6][kBlurTW];
    int level = 0;
    while (level + 1 < bl.nlevels && (int)blockIdx.x >= bl.tile_begin[level + 1]) ++level;
    const LevelGeom &L = g.lv[level];
    const int t = blockIdx.x - bl.tile_begin[level];
    const int tx0 = (t % bl.tiles_x[level]) * kBlurTW, ty0 = (t / bl.tiles_x[level]) * kBlurTH;
    const int frame = blockIdx.y;
    const uint8_t *src;
    int spitch;
    if (level == 0) { src = lvl0_base + (int64_t)frame * lvl0_stride; spitch = lvl0_pitch; }
    else { src = pyr + L.plane_offset + (int64_t)frame * L.plane_stride; spitch = L.pitch; }
    uint8_t *dst = blur + L.plane_offset + (int64_t)frame * L.plane_stride;

    const int tid = threadIdx.x;
    for (int i = tid; i < (kBlurTH + 6) * (kBlurTW + 6); i += 256) {
        const int ly = i / (kBlurTW + 6), lx = i - ly * (kBlurTW + 6);
        const int sy = reflect101(min(ty0 + ly - 3, L.h + 2), L.h);  // rows/cols past the image are never used
        const int sx = reflect101(min(tx0 + lx - 3, L.w + 2), L.w);
        tile[ly][lx] = __ldg(src + (int64_t)sy * spitch + sx);
    }
    __syncthreads();
    for (int i = tid; i < (kBlurTH + 6) * kBlurTW; i += 256) {
        const int ly = i / kBlurTW, lx = i - ly * kBlurTW;
        const uint8_t *p = &tile[ly][lx];
        const int h = 18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3];
        hbuf[ly][lx] = (uint16_t)h;
    }
    __syncthreads();
    // each thread: 4 adjacent columns x 2 rows -> two 32-bit stores
    for (int i = tid; i < kBlurTH * (kBlurTW / 4); i += 256) {
        const int ly = i / (kBlurTW / 4), lx = (i - ly * (kBlurTW / 4)) * 4;
        const int y = ty0 + ly, x = tx0 + lx;
        if (y >= L.h || x >= L.w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = 18u * (hbuf[ly][lx + k] + hbuf[ly + 6][lx + k]) +
                               34u * (hbuf[ly + 1][lx + k] + hbuf[ly + 5][lx + k]) +
                               48u * (hbuf[ly + 2][lx + k] + hbuf[ly + 4][lx + k]) + 56u * hbuf[ly + 3][lx + k];
            packed |= ((v + 32768u) >> 16) << (8 * k);
        }
        *reinterpret_cast<uint32_t *>(dst + (int64_t)y * L.pitch + x) = packed;
    }
}

void launch_blur(const FrameGeom &g, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride, const uint8_t *pyr,
                 uint8_t *blur, int nframes, cudaStream_t s) {
    BlurLevels bl;
    bl.nlevels = g.nlevels;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        bl.tile_begin[l] = total;
        bl.tiles_x[l] = (g.lv[l].w + kBlurTW - 1) / kBlurTW;
        total += bl.tiles_x[l] * ((g.lv[l].h + kBlurTH - 1) / kBlurTH);
    }
    bl.tile_begin[g.nlevels] = total;
    blur_kernel<<<dim3(total, nframes), 256, 0, s>>>(g, bl, lvl0_base, lvl0_pitch, lvl0_stride, pyr, blur);
    count_launch();
}

}  // namespace vsg
