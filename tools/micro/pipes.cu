// Micro-benchmark: which pipe do packed min/max variants issue on (sm_100a)?  Times N dependent-chain-free
// min/max instructions per thread for (a) VIMNMX.U16x2, (b) HMNMX2 (half2), (c) an interleaved mix.
// If (c) ~ max(a, b) / 1 the two use different pipes; if (c) ~ a + b they share one.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t imin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t imax2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
template <int MODE>
__global__ void k(uint32_t *out, int iters) {
    uint32_t a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0x64016402u + threadIdx.x * 3 + i; b[i] = 0x64036401u + threadIdx.x * 5 + i * 7; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = imin2(a[i], b[i]); b[i] = imax2(b[i], a[(i + 1) & 7]); }
            if (MODE == 1) { a[i] = hmin2(a[i], b[i]); b[i] = hmax2(b[i], a[(i + 1) & 7]); }
            if (MODE == 2) { a[i] = imin2(a[i], b[i]); b[i] = hmax2(b[i], a[(i + 1) & 7]); }
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    uint32_t *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters);
            if (mode == 1) k<1><<<148 * 8, 256>>>(out, iters);
            if (mode == 2) k<2><<<148 * 8, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("mode %d: %.3f ms  -> %.1f warp-instr/clk/SM at 1.9 GHz\n", mode, ms,
                            (double)iters * 16 * 148 * 8 * 8 / (ms * 1e-3) / 148 / 1.9e9);
        }
    }
    return 0;
}
