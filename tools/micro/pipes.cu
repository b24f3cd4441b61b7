// Micro-benchmark (sm_100a): which pipes do the packed min/max variants issue on, and at what rate?
// Each mode runs, per inner iteration and per thread, NA three-input VIMNMX3.U16x2 and NB instructions of a
// second kind on independent registers.  If the second kind lives on another pipe the mixed modes take
// max(time_A, time_B); if it shares the ALU pipe they take time_A + time_B.
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu     run: ./pipes
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t imin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t imin3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("{\n\t.reg .b32 t;\n\tmin.u16x2 t, %1, %2;\n\tmin.u16x2 %0, t, %3;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t hfma2relu(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
__device__ __forceinline__ uint32_t vmin4(uint32_t a, uint32_t b) { return __vminu4(a, b); }
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

// KIND: 0 none, 1 VIMNMX (2-input), 2 HMNMX2, 3 HADD2, 4 HFMA2.RELU, 5 IMAD, 6 __vminu4, 7 LOP3
template <int NA, int NB, int KIND>
__global__ void k(uint32_t *out, int iters) {
    uint32_t a[8], b[8], c[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = 0x64016402u + threadIdx.x * 3 + i; b[i] = 0x64036401u + threadIdx.x * 5 + i * 7;
        c[i] = 0x64056407u + threadIdx.x + i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NA; ++i) a[i & 7] = imin3(a[i & 7], b[(i + 1) & 7], a[(i + 3) & 7]);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            uint32_t &x = c[i & 7];
            const uint32_t y = c[(i + 1) & 7], z = c[(i + 5) & 7];
            if (KIND == 1) x = imin2(x, y);
            if (KIND == 2) x = hmin2(x, y);
            if (KIND == 3) x = hadd2(x, y);
            if (KIND == 4) x = hfma2relu(x, y, z);
            if (KIND == 5) x = imad(x, y, z);
            if (KIND == 6) x = vmin4(x, y);
            if (KIND == 7) x = lop3(x, y, z);
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] ^ b[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NA, int NB, int KIND>
void run(const char *name, uint32_t *out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<NA, NB, KIND><<<148 * 8, 256>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    // cycles per SMSP per inner iteration per warp: each SMSP holds 8*256/32/4 = 16 warps
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 16.0;
    printf("%-34s NA=%2d NB=%2d : %.3f ms, %.2f cycles per (warp, iteration) per SMSP\n", name, NA, NB, ms, cyc);
}

int main() {
    uint32_t *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<16, 0, 0>("VIMNMX3 only", out);
    run<0, 16, 1>("VIMNMX (2-input) only", out);
    run<0, 16, 2>("HMNMX2 only", out);
    run<0, 16, 3>("HADD2 only", out);
    run<0, 16, 4>("HFMA2.RELU only", out);
    run<0, 16, 5>("IMAD only", out);
    run<0, 16, 6>("__vminu4 only", out);
    run<0, 16, 7>("LOP3 only", out);
    run<16, 16, 1>("VIMNMX3 + VIMNMX", out);
    run<16, 16, 2>("VIMNMX3 + HMNMX2", out);
    run<16, 8, 2>("VIMNMX3 + HMNMX2", out);
    run<16, 16, 3>("VIMNMX3 + HADD2", out);
    run<16, 16, 4>("VIMNMX3 + HFMA2.RELU", out);
    run<16, 16, 5>("VIMNMX3 + IMAD", out);
    run<16, 16, 7>("VIMNMX3 + LOP3", out);
    return 0;
}
