// blur_tc.cu — the 7x7 Gaussian blur of every pyramid level as two banded u8 GEMMs on the tensor cores.
//
// Reference (snt-arg/visual_sgraphs): GaussianBlur(workingMat, workingMat, Size(7, 7), 2, 2, BORDER_REFLECT_101) on a clone of
// every level, orb_slam3/src/ORBextractor.cc:1129-1130.  OpenCV's 8-bit path is exactly linear:
//   h = sum_k taps[k] * src(x + k - 3)   (<= 65280),   v = sum_j taps[j] * h(y + j - 3),   out = (v + 32768) >> 16,
// taps = {18, 34, 48, 56, 48, 34, 18}.  A linear stencil is a product with a banded (Toeplitz) matrix, and u8 x u8 -> s32 is
// exact on tcgen05.mma.kind::i8, so per tile of 122 output rows x 96 output columns, with W the 128 x 128-byte source window
// (3 rows above the tile, 16 columns left of it so that the TMA box starts on a 16-byte boundary):
//   GEMM 1 (horizontal, transposed):  Ht[n][i] = sum_k Bh[n][k] * W[i][k],   Bh[n][k] = taps[k - 13 - n]     M128 N128 K128
//   GEMM 2 (vertical):                V[r][n]  = sum_k Bv[r][k] * Ht[n][k],  Bv[r][k] = taps[k - r]          M128 N96  K128
// Ht has 16-bit entries, so it is split into its high and low bytes (two u8 operands, two accumulators, v = 256 Vh + Vl).
// GEMM 1 is computed transposed so that one thread reads one row of Ht from tensor memory and writes it as one contiguous
// (swizzled) 128-byte row of GEMM 2's K-major operand.  Nine in ten multiplications are by a zero tap; they still cost a
// fraction of the issue slots the CUDA-core stencil needs, which is what the FAST kernel is short of.
//
// One persistent warp-specialised CTA per SM (640 threads); every hand-off is an mbarrier:
//   warp 3          TMA: source windows (cp.async.bulk.tensor.3d, SWIZZLE_128B, out-of-plane bytes zero-filled), 4 stages
//   warps 16, 19    REFLECT_101 (even / odd tiles): windows that touch the plane's border get the three out-of-plane columns /
//                   rows patched in shared memory from their mirror images
//   warp 7          one lane issues GEMM 1 (tcgen05.mma, A = Bh in tensor memory, B = the window), tcgen05.commit
//   warps 0-2, 4-6  tcgen05.ld of Ht (even / odd tiles, two buffers), byte split, operand rows of GEMM 2
//   warps 17, 18    one lane each issues GEMM 2 for the high / low byte plane (A = Bv in tensor memory)
//   warps 8-15      tcgen05.ld of Vh / Vl, rounding, 16-byte stores of the blurred rows
// Tensor memory: Ht x 2 buffers (256 columns), Vh, Vl (96 each), Bh, Bv (32 each).  Results are bit-identical to
// blur_block_body (blur_device.cuh), which stays for small batches and planes TMA cannot address.  profiles/r02_blur_tc.md
// has the measurements behind this layout.
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
constexpr int kBtStages = 4;
constexpr int kBtThreads = 32 * 20;
// warps 0-2 / 4-6: Ht epilogue of the even / odd tiles (TMEM lane quarters 0-2; quarter 3 of Ht is padding, so the two warps
// that could only read it, 3 and 7, do other work); warps 8-15: V epilogue; 16, 19: patch (even / odd tiles); 17, 18: GEMM 2
// issue for the high / low byte plane
constexpr int kWarpTma = 3, kWarpMma1 = 7, kWarpPatch0 = 16, kWarpMma2h = 17, kWarpMma2l = 18, kWarpPatch1 = 19;
constexpr int kBtTileA = 128 * 128;             // source window / Bh / Bv: 128 rows of 128 bytes
constexpr int kBtTileB = kBtW * 128;            // byte planes of Ht: 96 rows of 128 bytes
constexpr int kBtSmem = kBtStages * kBtTileA + 4 * kBtTileB + 256 + 1024;

struct BlurTcParams {
    CUtensorMap map[8];       // per level: (x bytes, y rows, frame), box 128 x 128 x 1, SWIZZLE_128B
    int nlevels;
    int tile_begin[9];        // prefix sums of tiles per frame
    int ntx[8], w[8], h[8], dst_pitch[8];
    int64_t dst_offset[8], dst_stride[8];
    int tiles_per_frame;
    uint32_t tpf_rcp, ntx_rcp[8];   // ceil(2^32 / d): x / d == umulhi(x, rcp) for x * d < 2^32; 0 stands for d == 1
};

// byte (row, col) of a 128-byte-row tile in the SWIZZLE_128B layout (tile base 1024-byte aligned)
__device__ __forceinline__ int swz(int row, int col) { return row * 128 + ((((col >> 4) ^ (row & 7)) << 4) | (col & 15)); }

struct BtTile {
    int level, frame, x0, y0;
};
__device__ __forceinline__ BtTile bt_tile(const BlurTcParams &p, int t) {
    BtTile r;
    r.frame = p.tpf_rcp ? (int)__umulhi((uint32_t)t, p.tpf_rcp) : t;
    const int rem = t - r.frame * p.tiles_per_frame;
    int level = 0;
#pragma unroll
    for (int l = 1; l < 8; ++l) level += (l < p.nlevels && rem >= p.tile_begin[l]) ? 1 : 0;
    r.level = level;
    const int local = rem - p.tile_begin[level];
    const uint32_t nr = p.ntx_rcp[level];
    const int ty = nr ? (int)__umulhi((uint32_t)local, nr) : local, tx = local - ty * p.ntx[level];
    r.x0 = tx * kBtW;
    r.y0 = ty * kBtH;
    return r;
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// tcgen05.ld .pack::16b: register j = low 16 bits of column 2j | low 16 bits of column 2j + 1 << 16.  Every accumulator of
// this kernel is below 2^16, and tensor-memory reads are the kernel's bound (64 B / clk / SM), so half the registers per
// column is half the time.
#define R4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
__device__ __forceinline__ void tc_ld_pack8(uint32_t taddr, uint32_t (&v)[8]) {        // 16 columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : R4(v, 0), R4(v, 4) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_pack16(uint32_t taddr, uint32_t (&v)[16]) {      // 32 columns
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_pack32(uint32_t taddr, uint32_t (&v)[32]) {      // 64 columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : R4(v, 0), R4(v, 4), R4(v, 8), R4(v, 12), R4(v, 16), R4(v, 20), R4(v, 24), R4(v, 28) : "r"(taddr) : "memory");
}
#undef R4
// u8 x u8 -> s32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t bt_idesc(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// ring position of the k-th use of an n-deep ring: slot and the parity its `full` barrier completes with
struct Ring {
    int slot = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void next(int n) { if (++slot == n) { slot = 0; phase ^= 1u; } }
};

// -DBT_TRACE: CTA 0 records clock64() at the hand-offs of its first 64 tiles (tools/micro/bt_trace.py prints the timeline)
#ifdef BT_TRACE
__device__ long long g_bt_trace[6][64][4];
#define TR(role, it, ev) do { if (blockIdx.x == 0 && lane == 0 && (it) < 64) g_bt_trace[role][it][ev] = clock64(); } while (0)
#else
#define TR(role, it, ev) do { } while (0)
#endif

__global__ void __launch_bounds__(kBtThreads, 1) blur_tc_kernel(const __grid_constant__ BlurTcParams p, uint8_t *__restrict__ blur,
                                                                int total_tiles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *a1 = smem;                                   // kBtStages source windows
    uint8_t *b2 = a1 + kBtStages * kBtTileA;              // 2 buffers x {high, low} byte plane of Ht as [n][k]
    uint64_t *bars = reinterpret_cast<uint64_t *>(b2 + 4 * kBtTileB);
    uint64_t *in_full = bars, *in_empty = bars + kBtStages, *a_ready = bars + 2 * kBtStages;
    uint64_t *d1_full = bars + 3 * kBtStages, *d1_empty = d1_full + 2, *b2_full = d1_full + 4, *b2_empty = d1_full + 6;
    uint64_t *d2_full = d1_full + 8, *d2_empty = d1_full + 9;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d1_full + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kBtStages; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], 1); mbar_init(&a_ready[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 3); mbar_init(&b2_full[i], 3); mbar_init(&b2_empty[i], 2); }
        mbar_init(d2_full, 2);
        mbar_init(d2_empty, 8);
        mbar_init_fence();
    }
    if (warp == kWarpMma1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // tensor-memory columns: Ht buffers at 0 and 128, Vh, Vl, then the two constant band matrices as A operands
    const uint32_t tD2h = tmem + 256, tD2l = tmem + 256 + kBtW, tBh = tmem + 448, tBv = tmem + 480;
    if (warp < 4) {
        // A operands live in tensor memory (they never change, and operand reads are what saturates shared memory): lane m holds
        // row m of the M x K matrix, column j its bytes k = 4j .. 4j + 3.  Bh[n][k] = taps[k - 13 - n] (n < 96), Bv[r][k] = taps[k - r] (r < 122)
        const uint64_t taps = 0x12223038302212ull;          // {18, 34, 48, 56, 48, 34, 18}, one byte each
        const int m = warp * 32 + lane;
        uint32_t ah[32], av[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            uint32_t wh = 0, wv = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int k = 4 * j + b, th = k - (kBtXoff - 3) - m, tv = k - m;
                if (m < kBtW && th >= 0 && th < 7) wh |= (uint32_t)((taps >> (8 * th)) & 255) << (8 * b);
                if (m < kBtH && tv >= 0 && tv < 7) wv |= (uint32_t)((taps >> (8 * tv)) & 255) << (8 * b);
            }
            ah[j] = wh;
            av[j] = wv;
        }
        tc_st32(tBh + ((uint32_t)(warp * 32) << 16), ah);
        tc_st32(tBv + ((uint32_t)(warp * 32) << 16), av);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == kWarpTma) {
        // ===== TMA producer =====
        if (lane == 0) {
            Ring ld;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                TR(0, it, 0);
                const BtTile tl = bt_tile(p, t);
                mbar_wait(&in_empty[ld.slot], ld.phase ^ 1);
                TR(0, it, 1);
                mbar_expect_tx(&in_full[ld.slot], kBtTileA);
                tma_load_3d(a1 + ld.slot * kBtTileA, &p.map[tl.level], tl.x0 - kBtXoff, tl.y0 - 3, tl.frame, &in_full[ld.slot]);
                TR(0, it, 2);
                ld.next(kBtStages);
            }
        }
    } else if (warp == kWarpPatch0 || warp == kWarpPatch1) {
        // ===== REFLECT_101 of the plane itself: windows that touch its border get the three out-of-plane columns /
        // rows from their mirror images.  Loads are staged in registers so that they do not queue behind the stores. =====
        static_assert(kBtStages == 4, "slot / phase arithmetic below");
        struct { int slot; uint32_t phase; } pt;
        int it = warp == kWarpPatch1 ? 1 : 0;
        for (int64_t t = (int64_t)blockIdx.x + (int64_t)it * gridDim.x; t < total_tiles; t += 2 * (int64_t)gridDim.x, it += 2) {
            TR(1, it, 0);
            pt.slot = it & 3;
            pt.phase = (it >> 2) & 1;
            const BtTile tl = bt_tile(p, (int)t);
            const uint32_t A = smem_u32(a1) + pt.slot * kBtTileA;
            const int w = p.w[tl.level], h = p.h[tl.level];
            const int kw = kBtXoff + (w - tl.x0);                // window column of x = w   (>= 17)
            const int rb = h - tl.y0 + 3;                        // window row of y = h      (>= 4)
            const bool left = tl.x0 == 0, right = kw < 128, top = tl.y0 == 0, bottom = rb < 128;
            TR(1, it, 1);
            mbar_wait(&in_full[pt.slot], pt.phase);
            TR(1, it, 2);
            if (left || right || top || bottom) {
                if (left || right) {                             // columns first (every row of the window) ...
                    uint32_t cl[4][3], cr[4][3];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = rr * 32 + lane;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            cl[rr][j] = left ? lds_u8(A + swz(row, kBtXoff + 3 - j)) : 0;      // x = -3 + j  <-  x = 3 - j
                            cr[rr][j] = right ? lds_u8(A + swz(row, kw - 2 - j)) : 0;          // x = w + j   <-  x = w - 2 - j
                        }
                    }
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int row = rr * 32 + lane;
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            if (left) sts_u8(A + swz(row, kBtXoff - 3 + j), cl[rr][j]);
                            if (right && kw + j < 128) sts_u8(A + swz(row, kw + j), cr[rr][j]);
                        }
                    }
                    __syncwarp();
                }
                if (top || bottom) {                             // ... then whole rows, corners included
                    const int c = lane * 4;
                    uint32_t rt[3], rbm[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        rt[j] = top ? lds_u32(A + swz(6 - j, c)) : 0;                                 // y = -3 + j <- 3 - j
                        rbm[j] = bottom && rb - 2 - j >= 0 ? lds_u32(A + swz(rb - 2 - j, c)) : 0;     // y = h + j <- h - 2 - j
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        if (top) sts_u32(A + swz(j, c), rt[j]);
                        if (bottom && rb + j < 128 && rb - 2 - j >= 0) sts_u32(A + swz(rb + j, c), rbm[j]);
                    }
                }
                fence_async_smem();
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&a_ready[pt.slot]);
            TR(1, it, 3);
        }
    } else if (warp == kWarpMma1) {
        // ===== GEMM 1 issuer: Ht[n][i] = sum_k Bh[n][k] W[i][k] into Ht buffer (tile & 1) =====
        if (lane == 0) {
            Ring in;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                TR(2, it, 0);
                mbar_wait(&a_ready[in.slot], in.phase);              // window landed (the patch warp saw in_full) and patched
                TR(2, it, 1);
                mbar_wait(&d1_empty[buf], ((it >> 1) & 1) ^ 1);      // this Ht buffer has been read
                TR(2, it, 2);
                tc_fence_after();
                const uint64_t d_w = tc_smem_desc(a1 + in.slot * kBtTileA);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_i8_ts(tmem + buf * 128, tBh + 8 * k, d_w + 2 * k, bt_idesc(128), k ? 1u : 0u);
                tc_commit(&in_empty[in.slot]);
                tc_commit(&d1_full[buf]);
                TR(2, it, 3);
                in.next(kBtStages);
            }
        }
    } else if (warp == kWarpMma2h || warp == kWarpMma2l) {
        // ===== GEMM 2 issuers: V[r][n] = sum_k Bv[r][k] Ht[n][k], one thread per byte plane =====
        if (lane == 0) {
            const int plane = warp == kWarpMma2l ? 1 : 0;
            const uint32_t tD2 = plane ? tD2l : tD2h;
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int buf = it & 1;
                if (plane == 0) TR(3, it, 0);
                mbar_wait(&b2_full[buf], (it >> 1) & 1);             // byte planes of Ht are in shared memory
                if (plane == 0) TR(3, it, 1);
                mbar_wait(d2_empty, (it & 1) ^ 1);                   // V of the previous tile has been read
                if (plane == 0) TR(3, it, 2);
                tc_fence_after();
                const uint64_t d_b = tc_smem_desc(b2 + (buf * 2 + plane) * kBtTileB);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_i8_ts(tD2, tBv + 8 * k, d_b + 2 * k, bt_idesc(kBtW), k ? 1u : 0u);
                tc_commit(&b2_empty[buf]);
                tc_commit(d2_full);
                if (plane == 0) TR(3, it, 3);
            }
        }
    } else if (warp < 8) {
        // ===== Ht -> byte operands of GEMM 2: warps 0-2 take the even tiles (buffer 0), warps 4-6 the odd ones (buffer 1) =====
        const int buf = warp >> 2, q = warp & 3, n = q * 32 + lane;    // TMEM lane = output column of the tile
        const uint32_t taddr = tmem + buf * 128 + ((uint32_t)(q * 32) << 16);
        const uint32_t row_h = smem_u32(b2) + buf * 2 * kBtTileB + n * 128, row_l = row_h + kBtTileB;
        const int x7 = (n & 7) << 4;
        uint32_t phase = 0;
        int it = buf;                                            // tile index within this CTA (trace only)
        for (int64_t t = (int64_t)blockIdx.x + (int64_t)buf * gridDim.x; t < total_tiles; t += 2 * (int64_t)gridDim.x, phase ^= 1, it += 2) {
            if (q == 0) TR(4, it, 0);
            mbar_wait(&d1_full[buf], phase);
            tc_fence_after();
            mbar_wait(&b2_empty[buf], phase ^ 1);                // GEMM 2 of two tiles ago no longer reads these byte planes
            if (q == 0) TR(4, it, 1);
#pragma unroll
            for (int c = 0; c < 2; ++c) {                        // window rows 64c .. 64c + 63, two per register
                uint32_t v[32];
                tc_ld_pack32(taddr + c * 64, v);
                tc_wait_ld();
                if (c == 1) {                                    // everything is in registers: GEMM 1 may overwrite this buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d1_empty[buf]);
                    if (q == 0) TR(4, it, 2);
                }
#pragma unroll
                for (int m = 0; m < 4; ++m) {                    // 16-byte chunk 4c + m of the row, swizzled
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {                // bytes of a register: lo(H0) hi(H0) lo(H1) hi(H1)
                        lo[g] = __byte_perm(v[8 * m + 2 * g], v[8 * m + 2 * g + 1], 0x6420);
                        hi[g] = __byte_perm(v[8 * m + 2 * g], v[8 * m + 2 * g + 1], 0x7531);
                    }
                    const int off = ((4 * c + m) << 4) ^ x7;
                    sts_v4(row_h + off, hi[0], hi[1], hi[2], hi[3]);
                    sts_v4(row_l + off, lo[0], lo[1], lo[2], lo[3]);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&b2_full[buf]);
            if (q == 0) TR(4, it, 3);
        }
    } else if (warp < 16) {
        // ===== V -> blurred rows: warp (q, half) owns rows 32q.. and columns 48 half.. =====
        const int q = warp & 3, half = (warp - 8) >> 2, r = q * 32 + lane;
        uint32_t v_phase = 0;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const BtTile tl = bt_tile(p, t);
            const int w = p.w[tl.level], h = p.h[tl.level];
            const int y = tl.y0 + r;
            const bool row_ok = r < kBtH && y < h;
            const int xh = tl.x0 + half * 48;
            uint8_t *dst = blur + p.dst_offset[tl.level] + (int64_t)tl.frame * p.dst_stride[tl.level] + (int64_t)y * p.dst_pitch[tl.level] + xh;
            const uint32_t col = ((uint32_t)(q * 32) << 16) + half * 48;
            mbar_wait(d2_full, v_phase);
            tc_fence_after();
            if (warp == 8) TR(5, it, 1);
            // 256 Vh + Vl + 32768 >> 16  ==  Vh + (Vl >> 8) + 128 >> 8, and Vh + (Vl >> 8) <= 65280: it stays inside 16-bit lanes
            uint32_t vh0[16], vl0[16], vh1[8], vl1[8];       // all loads first: V is released as early as possible
            tc_ld_pack16(tD2h + col, vh0);
            tc_ld_pack16(tD2l + col, vl0);
            tc_ld_pack8(tD2h + col + 32, vh1);
            tc_ld_pack8(tD2l + col + 32, vl1);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d2_empty);
            if (warp == 8) TR(5, it, 2);
            auto round2 = [](uint32_t h, uint32_t l) { return h + __byte_perm(l, 0, 0x4341) + 0x00800080u; };   // results in bytes 1 and 3
            uint4 o[3];
            {
                uint32_t wv[12];
#pragma unroll
                for (int g = 0; g < 8; ++g) wv[g] = __byte_perm(round2(vh0[2 * g], vl0[2 * g]), round2(vh0[2 * g + 1], vl0[2 * g + 1]), 0x7531);
#pragma unroll
                for (int g = 0; g < 4; ++g) wv[8 + g] = __byte_perm(round2(vh1[2 * g], vl1[2 * g]), round2(vh1[2 * g + 1], vl1[2 * g + 1]), 0x7531);
#pragma unroll
                for (int c = 0; c < 3; ++c) o[c] = make_uint4(wv[4 * c], wv[4 * c + 1], wv[4 * c + 2], wv[4 * c + 3]);
            }
            if (row_ok) {
                // whole 16-byte chunks only: x and the pitch are multiples of 16, so a chunk that starts inside the row's pitch ends
                // inside it, and the bytes between w and the pitch are padding nobody reads
                const int pitch = p.dst_pitch[tl.level];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if (xh + c * 16 < w && xh + c * 16 + 16 <= pitch) *reinterpret_cast<uint4 *>(dst + c * 16) = o[c];
            }
            if (warp == 8) TR(5, it, 3);
            v_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

#ifdef BT_TRACE
extern "C" int vsg_debug_bt_trace(long long *out) { return (int)cudaMemcpyFromSymbol(out, g_bt_trace, sizeof(g_bt_trace)); }
#endif

// VSG_BLUR_TC = n: batches of at least n frames blur on the tensor cores (0 = never); default 16
static int blur_tc_min_frames() {
    const char *e = getenv("VSG_BLUR_TC");
    const int v = e ? atoi(e) : 16;
    return v <= 0 ? INT32_MAX : v;
}

struct BlurTcPlan {
    BlurTcParams p;
    int64_t tiles;
};

// Plans the tensor-core blur of `nframes` frames: tensor maps and the tile table.  nullptr if it does not apply (small batch,
// planes TMA cannot address) — the caller then uses the CUDA-core blur.  The plan lives in thread-local storage until the next call.
const BlurTcPlan *plan_blur_tc(const FrameGeom &g, const uint8_t *lvl0_base, int lvl0_pitch, int64_t lvl0_stride, const uint8_t *pyr,
                               const uint8_t *blur, int nframes) {
    if (nframes < blur_tc_min_frames() || g.nlevels > 8) return nullptr;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return nullptr;
    static thread_local BlurTcPlan plan;
    BlurTcParams &p = plan.p;
    p.nlevels = g.nlevels;
    int total = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom &L = g.lv[l];
        const uint8_t *base = l == 0 ? lvl0_base : pyr + L.plane_offset;
        const int64_t pitch = l == 0 ? lvl0_pitch : L.pitch, stride = l == 0 ? lvl0_stride : L.plane_stride;
        if (((uintptr_t)base & 15) || (pitch & 15) || (stride & 15) || (((uintptr_t)(blur + L.plane_offset)) & 15) ||
            (L.pitch & 15) || (L.plane_stride & 15))
            return nullptr;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)L.h, (cuuint64_t)nframes};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)stride};
        const cuuint32_t box[3] = {128, 128, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        if (fn(&p.map[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return nullptr;
        p.tile_begin[l] = total;
        p.ntx[l] = (L.w + kBtW - 1) / kBtW;
        p.ntx_rcp[l] = p.ntx[l] == 1 ? 0u : (uint32_t)((0x100000000ull + p.ntx[l] - 1) / p.ntx[l]);
        total += p.ntx[l] * ((L.h + kBtH - 1) / kBtH);
        p.w[l] = L.w; p.h[l] = L.h; p.dst_pitch[l] = L.pitch;
        p.dst_offset[l] = L.plane_offset; p.dst_stride[l] = L.plane_stride;
    }
    for (int l = g.nlevels; l < 8; ++l) { p.map[l] = p.map[0]; p.tile_begin[l] = total; p.ntx[l] = 1; p.ntx_rcp[l] = 0; p.w[l] = p.h[l] = 0; p.dst_pitch[l] = 0; p.dst_offset[l] = p.dst_stride[l] = 0; }
    p.tile_begin[8] = total;
    p.tiles_per_frame = total;
    const int64_t tiles = (int64_t)total * nframes;
    if (tiles <= 0 || tiles * total >= (1ll << 32)) return nullptr;
    p.tpf_rcp = total == 1 ? 0u : (uint32_t)((0x100000000ull + total - 1) / total);
    if (cudaFuncSetAttribute(blur_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBtSmem) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    plan.tiles = tiles;
    return &plan;
}

void launch_blur_tc(const BlurTcPlan *plan, uint8_t *blur, cudaStream_t s) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    blur_tc_kernel<<<(unsigned)std::min<int64_t>(plan->tiles, sms), kBtThreads, kBtSmem, s>>>(plan->p, blur, (int)plan->tiles);
    count_launch();
}

}  // namespace vsg
