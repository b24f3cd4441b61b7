// blur_tc.cu — the 7x7 Gaussian blur of every pyramid level as two banded int8 GEMMs on the tensor cores.
//
// Reference (snt-arg/visual_sgraphs): GaussianBlur(workingMat, workingMat, Size(7, 7), 2, 2, BORDER_REFLECT_101) on a clone of
// every level, orb_slam3/src/ORBextractor.cc:1129-1130.  OpenCV's 8-bit path is exactly linear:
//   h = sum_k taps[k] * src(x + k - 3)   (<= 65280),   v = sum_j taps[j] * h(y + j - 3),   out = (v + 32768) >> 16,
// taps = {18, 34, 48, 56, 48, 34, 18}.  A linear stencil is a product with a banded (Toeplitz) matrix, and u8 x u8 -> s32 is
// exact on tcgen05.mma.kind::i8, so per tile of 122 output rows x 96 output columns:
//   GEMM 1 (horizontal):  H[128 rows][96] = In[128 rows][128 cols] * Bh^T,  Bh[n][k] = taps[k - 13 - n]
//   GEMM 2 (vertical)  :  V[128][96]      = Bv[128][128] * H,               Bv[r][k] = taps[k - r]
// H has 16-bit entries, so it is split into its high and low bytes (two u8 operands, two accumulators, v = (Vh << 8) + Vl).
// The stencil's 98 multiply-adds per pixel become a few hundred on the tensor pipe — 90 % of them with zero taps — and
// still cost a fraction of what the CUDA cores need, because they leave the issue slots and the ALU pipe to the FAST kernel.
//
// One persistent warp-specialised CTA per SM:
//   warp 0      TMA: the tile's 128 x 128-byte source window (cp.async.bulk.tensor.3d, SWIZZLE_128B, starts 16 bytes left of
//               the tile so that the box is 16-byte aligned, 3 rows above it), two stages
//   warps 2-5   (a) REFLECT_101: tiles that touch the plane's border get the three out-of-plane columns / rows patched in
//               shared memory from their mirror images (TMA fills out-of-bounds bytes with zeros);
//               (b) after GEMM 1: tcgen05.ld of H, split into bytes, written as the K-major operand of GEMM 2
//   warp 1      one lane issues GEMM 1 (4 x M128 N96 K32) and GEMM 2 (2 x 4), tcgen05.commit
//   warps 6-9   tcgen05.ld of Vh / Vl, rounding, 16-byte stores of the blurred rows
// Results are bit-identical to blur_block_body (blur_device.cuh), which stays for small batches and unaligned inputs.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "tma_util.cuh"
#include "vsg_internal.cuh"

namespace vsg {

constexpr int kBtW = 96, kBtH = 122;            // output tile
constexpr int kBtXoff = 16;                     // the source box starts this many columns left of the tile (16-byte alignment)
constexpr int kBtThreads = 32 * 10;
constexpr int kBtTileA = 128 * 128;             // source window / Bv: 128 rows of 128 bytes
constexpr int kBtTileB = kBtW * 128;            // Bh / H operands: 96 rows of 128 bytes
constexpr int kBtSmem = 2 * kBtTileA + kBtTileB + kBtTileA + 2 * kBtTileB + 256 + 1024;

struct BlurTcParams {
    CUtensorMap map[8];       // per level: (x bytes, y rows, frame), box 128 x 128 x 1, SWIZZLE_128B
    int nlevels;
    int tile_begin[9];        // prefix sums of tiles per frame
    int ntx[8], w[8], h[8], dst_pitch[8];
    int64_t dst_offset[8], dst_stride[8];
    int tiles_per_frame;
};

// byte (row, col) of a 128-byte-row tile in the SWIZZLE_128B layout (tile base 1024-byte aligned)
__device__ __forceinline__ int swz(int row, int col) { return row * 128 + ((((col >> 4) ^ (row & 7)) << 4) | (col & 15)); }

struct BtTile {
    int level, frame, x0, y0;
};
__device__ __forceinline__ BtTile bt_tile(const BlurTcParams &p, int t) {
    BtTile r;
    r.frame = t / p.tiles_per_frame;
    const int rem = t - r.frame * p.tiles_per_frame;
    int level = 0;
#pragma unroll
    for (int l = 1; l < 8; ++l) level += (l < p.nlevels && rem >= p.tile_begin[l]) ? 1 : 0;
    r.level = level;
    const int local = rem - p.tile_begin[level];
    const int ty = local / p.ntx[level], tx = local - ty * p.ntx[level];
    r.x0 = tx * kBtW;
    r.y0 = ty * kBtH;
    return r;
}

__device__ __forceinline__ void bar_sync_128(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// u8 x u8 -> s32, both operands K-major, M = 128, N = 96
constexpr uint32_t kBtIdesc = (2u << 4) | ((uint32_t)(kBtW >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(kBtThreads, 1) blur_tc_kernel(const __grid_constant__ BlurTcParams p, uint8_t *__restrict__ blur,
                                                                int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a1 = smem;                               // 2 stages of the source window
    uint8_t *bh = a1 + 2 * kBtTileA;                  // Bh[n][k]
    uint8_t *bv = bh + kBtTileB;                      // Bv[r][k]
    uint8_t *b2h = bv + kBtTileA, *b2l = b2h + kBtTileB;   // H high / low bytes as [n][k]
    uint64_t *bars = reinterpret_cast<uint64_t *>(b2l + kBtTileB);
    uint64_t *in_full = bars, *in_empty = bars + 2, *a_ready = bars + 4;
    uint64_t *d1_full = bars + 6, *d1_empty = bars + 7, *b2_full = bars + 8, *b2_empty = bars + 9, *d2_full = bars + 10,
             *d2_empty = bars + 11;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the two constant band matrices
    for (int i = threadIdx.x; i < (kBtTileB + kBtTileA) / 16; i += kBtThreads) reinterpret_cast<uint4 *>(bh)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    {
        const uint64_t taps = 0x12223038302212ull;          // {18, 34, 48, 56, 48, 34, 18}, one byte each
        for (int i = threadIdx.x; i < kBtW * 7; i += kBtThreads) {
            const int n = i / 7, j = i - n * 7;
            bh[swz(n, kBtXoff - 3 + n + j)] = (uint8_t)(taps >> (8 * j));
        }
        for (int i = threadIdx.x; i < kBtH * 7; i += kBtThreads) {
            const int r = i / 7, j = i - r * 7;
            bv[swz(r, r + j)] = (uint8_t)(taps >> (8 * j));
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); mbar_init(&a_ready[i], 4); }
        mbar_init(d1_full, 1); mbar_init(d1_empty, 4);
        mbar_init(b2_full, 4); mbar_init(b2_empty, 1);
        mbar_init(d2_full, 1); mbar_init(d2_empty, 4);
        mbar_init_fence();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();                               // the band matrices were written by ordinary stores
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tD1 = tmem, tD2h = tmem + 128, tD2l = tmem + 256;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int stage = it & 1;
                const BtTile tl = bt_tile(p, t);
                mbar_wait(&in_empty[stage], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&in_full[stage], kBtTileA);
                tma_load_3d(a1 + stage * kBtTileA, &p.map[tl.level], tl.x0 - kBtXoff, tl.y0 - 3, tl.frame, &in_full[stage]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t it = 0;
            const uint64_t d_bh = tc_smem_desc(bh), d_bv = tc_smem_desc(bv), d_b2h = tc_smem_desc(b2h), d_b2l = tc_smem_desc(b2l);
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int stage = it & 1;
                mbar_wait(&in_full[stage], (it >> 1) & 1);       // window landed ...
                mbar_wait(&a_ready[stage], (it >> 1) & 1);       // ... and, on border tiles, patched
                mbar_wait(d1_empty, (it & 1) ^ 1);               // H of the previous tile has been read
                tc_fence_after();
                const uint64_t d_a1 = tc_smem_desc(a1 + stage * kBtTileA);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_i8(tD1, d_a1 + 2 * k, d_bh + 2 * k, kBtIdesc, k ? 1u : 0u);
                tc_commit(&in_empty[stage]);
                tc_commit(d1_full);
                mbar_wait(b2_full, it & 1);                      // H bytes are in shared memory
                mbar_wait(d2_empty, (it & 1) ^ 1);               // V of the previous tile has been read
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_i8(tD2h, d_bv + 2 * k, d_b2h + 2 * k, kBtIdesc, k ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_i8(tD2l, d_bv + 2 * k, d_b2l + 2 * k, kBtIdesc, k ? 1u : 0u);
                tc_commit(b2_empty);
                tc_commit(d2_full);
            }
        }
    } else if (warp < 6) {
        // ===== border patch, then H -> byte operands =====
        const int q = warp & 3, row = q * 32 + lane;             // TMEM lane = window row = K index of GEMM 2
        const int tid = (warp - 2) * 32 + lane;                  // 0..127 within this group
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int stage = it & 1;
            const BtTile tl = bt_tile(p, t);
            uint8_t *A = a1 + stage * kBtTileA;
            mbar_wait(&in_full[stage], (it >> 1) & 1);
            const int w = p.w[tl.level], h = p.h[tl.level];
            const int kw = kBtXoff + (w - tl.x0);                // window column of x = w
            const int rb = h - tl.y0 + 3;                        // window row of y = h
            const bool left = tl.x0 == 0, right = kw < 128, top = tl.y0 == 0, bottom = rb < 128;
            if (left || right || top || bottom) {                // REFLECT_101 of the plane itself
                if (left) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) A[swz(tid, kBtXoff - 3 + j)] = A[swz(tid, kBtXoff + 3 - j)];
                }
                if (right) {
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (kw + j < 128 && kw - 2 - j >= 0) A[swz(tid, kw + j)] = A[swz(tid, kw - 2 - j)];
                }
                bar_sync_128(1);
                if (top && tid < 96) {                           // rows y = -3..-1 <- y = 3, 2, 1
                    const int j = tid >> 5, c = (tid & 31) * 4;
                    *reinterpret_cast<uint32_t *>(A + swz(j, c)) = *reinterpret_cast<const uint32_t *>(A + swz(6 - j, c));
                }
                if (bottom && tid < 96) {                        // rows y = h..h+2 <- y = h-2, h-3, h-4
                    const int j = tid >> 5, c = (tid & 31) * 4;
                    if (rb + j < 128 && rb - 2 - j >= 0)
                        *reinterpret_cast<uint32_t *>(A + swz(rb + j, c)) = *reinterpret_cast<const uint32_t *>(A + swz(rb - 2 - j, c));
                }
                fence_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_ready[stage]);

            mbar_wait(d1_full, it & 1);
            tc_fence_after();
            mbar_wait(b2_empty, (it & 1) ^ 1);                   // GEMM 2 of the previous tile no longer reads the byte planes
            const uint32_t taddr = tD1 + ((uint32_t)(q * 32) << 16);
            const int kc = row >> 4, kl = row & 15;
            uint8_t *ph[8], *pl[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int off = m * 128 + (((kc ^ m) << 4) | kl);
                ph[m] = b2h + off;
                pl[m] = b2l + off;
            }
#pragma unroll
            for (int c = 0; c < kBtW / 32; ++c) {
                int v[32];
                tc_ld32(taddr + c * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = c * 32 + j;                    // operand row n, byte `row`: offset n * 128 + swizzled(row)
                    ph[n & 7][(n >> 3) * 1024] = (uint8_t)(v[j] >> 8);
                    pl[n & 7][(n >> 3) * 1024] = (uint8_t)v[j];
                }
            }
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(d1_empty); mbar_arrive(b2_full); }
        }
    } else {
        // ===== V -> blurred rows =====
        const int q = warp & 3, r = q * 32 + lane;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const BtTile tl = bt_tile(p, t);
            const int w = p.w[tl.level], h = p.h[tl.level];
            mbar_wait(d2_full, it & 1);
            tc_fence_after();
            const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
            const int y = tl.y0 + r;
            const bool row_ok = r < kBtH && y < h;
            uint8_t *dst = blur + p.dst_offset[tl.level] + (int64_t)tl.frame * p.dst_stride[tl.level] + (int64_t)y * p.dst_pitch[tl.level] + tl.x0;
#pragma unroll
            for (int c = 0; c < kBtW / 32; ++c) {
                int vh[32], vl[32];
                tc_ld32(tD2h + lane_bits + c * 32, vh);
                tc_ld32(tD2l + lane_bits + c * 32, vl);
                tc_wait_ld();
                uint32_t o[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint32_t wv = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const uint32_t v = ((uint32_t)vh[4 * g + b] << 8) + (uint32_t)vl[4 * g + b] + 32768u;
                        wv |= (v >> 16) << (8 * b);
                    }
                    o[g] = wv;
                }
                if (row_ok) {
                    const int x = tl.x0 + c * 32;
                    if (x + 32 <= w) {
                        reinterpret_cast<uint4 *>(dst + c * 32)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        reinterpret_cast<uint4 *>(dst + c * 32)[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {
#pragma unroll
                        for (int b = 0; b < 32; ++b)
                            if (x + b < w) dst[c * 32 + b] = (uint8_t)(o[b >> 2] >> (8 * (b & 3)));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// VSG_BLUR_TC = n: batches of at least n frames blur on the tensor cores (0 = never); default 16
static int blur_tc_min_frames() {
    const char *e = getenv("VSG_BLUR_TC");
    const int v = e ? atoi(e) : 16;
    return v <= 0 ? INT32_MAX : v;
}

// Launches the tensor-core blur of `nframes` frames if the planes can be addressed by TMA; returns false (nothing launched)
// otherwise — the caller then uses the CUDA-core blur.
bool launch_blur_tc(const FrameGeom &g, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride, const uint8_t *pyr, uint8_t *blur,
                    int nframes, cudaStream_t s) {
    if (nframes < blur_tc_min_frames() || g.nlevels > 8) return false;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    BlurTcParams p;
    p.nlevels = g.nlevels;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom &L = g.lv[l];
        const uint8_t *base = l == 0 ? lvl0_base : pyr + L.plane_offset;
        const int64_t pitch = l == 0 ? lvl0_pitch : L.pitch, stride = l == 0 ? lvl0_stride : L.plane_stride;
        if (((uintptr_t)base & 15) || (pitch & 15) || (stride & 15) || (((uintptr_t)(blur + L.plane_offset)) & 15)) return false;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)L.h, (cuuint64_t)nframes};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)stride};
        const cuuint32_t box[3] = {128, 128, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (fn(&p.map[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
        p.tile_begin[l] = total;
        p.ntx[l] = (L.w + kBtW - 1) / kBtW;
        total += p.ntx[l] * ((L.h + kBtH - 1) / kBtH);
        p.w[l] = L.w; p.h[l] = L.h; p.dst_pitch[l] = L.pitch;
        p.dst_offset[l] = L.plane_offset; p.dst_stride[l] = L.plane_stride;
    }
    for (int l = g.nlevels; l < 8; ++l) { p.map[l] = p.map[0]; p.tile_begin[l] = total; p.ntx[l] = 1; p.w[l] = p.h[l] = 0; p.dst_pitch[l] = 0; p.dst_offset[l] = p.dst_stride[l] = 0; }
    p.tile_begin[8] = total;
    p.tiles_per_frame = total;
    const int64_t tiles = (int64_t)total * nframes;
    if (tiles <= 0 || tiles > INT32_MAX) return false;
    if (cudaFuncSetAttribute(blur_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtSmem) != cudaSuccess) { cudaGetLastError(); return false; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    blur_tc_kernel<<<(unsigned)std::min<int64_t>(tiles, sms), kBtThreads, kBtSmem, s>>>(p, blur, (int)tiles);
    count_launch();
    return true;
}

}  // namespace vsg
