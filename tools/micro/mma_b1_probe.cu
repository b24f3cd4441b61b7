#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// throughput probe: legacy mma.sync b1 (and.popc / xor.popc) and s8 on sm_100a
template <int MODE>
__global__ void probe(int iters, int *out) {
    unsigned a[4] = {threadIdx.x * 2654435761u, 0x9e3779b9u, 0x7f4a7c15u, threadIdx.x ^ 0x85ebca6bu};
    unsigned b[2] = {0xc2b2ae35u ^ threadIdx.x, 0x27d4eb2fu};
    int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
    for (int i = 0; i < iters; ++i) {
#define MMA(C)                                                                                                            \
    if (MODE == 0)                                                                                                        \
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                     : "+r"(C[0]), "+r"(C[1]), "+r"(C[2]), "+r"(C[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1])); \
    else if (MODE == 1)                                                                                                   \
        asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.xor.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                     : "+r"(C[0]), "+r"(C[1]), "+r"(C[2]), "+r"(C[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1])); \
    else                                                                                                                  \
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};" \
                     : "+r"(C[0]), "+r"(C[1]), "+r"(C[2]), "+r"(C[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        MMA(c0) MMA(c1) MMA(c2) MMA(c3)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[2] + c3[3];
}
template <int MODE>
void run(const char *name, double macs_per_mma) {
    int *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000, blocks = 148 * 4, threads = 256;
    probe<MODE><<<blocks, threads>>>(100, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(iters, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)blocks * (threads / 32) * iters * 4;
    printf("%-10s %8.3f ms  %.3e MAC/s  (%.0f MAC/clk/SM at 1.965 GHz)  err=%s\n", name, ms, mmas * macs_per_mma / (ms * 1e-3),
           mmas * macs_per_mma / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    run<0>("b1.and", 16.0 * 8 * 256);
    run<1>("b1.xor", 16.0 * 8 * 256);
    run<2>("s8", 16.0 * 8 * 32);
    return 0;
}
