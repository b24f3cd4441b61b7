// popc_peak.cu — measured POPC throughput of this GPU: the denominator of the POPC-kernel roofline in bench.py
// ("confirm with a popc micro-benchmark on the box and use the measured peak as denominator").
// Every thread runs 8 independent dependent-chains of POPC (+ one XOR per POPC on the ALU pipe, as the Hamming kernel has).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o popc_peak popc_peak.cu && ./popc_peak   -> one JSON line
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) popc_loop(int iters, unsigned seed, unsigned *out) {
    unsigned a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed * (threadIdx.x + 1) + k * 0x9e3779b9u;
    unsigned acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] = __popc(a[k] ^ seed) + a[k];     // POPC (XU pipe) + XOR / ADD (ALU pipe)
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    unsigned *out; cudaMalloc(&out, (size_t)blocks * threads * 4);
    popc_loop<<<blocks, threads>>>(100, 12345u, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        popc_loop<<<blocks, threads>>>(iters, 12345u + r, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double popc = (double)blocks * threads * iters * 8;
    const double rate = popc / (best * 1e-3);
    printf("{\"popc_per_s\": %.4e, \"popc_per_clk_per_sm\": %.2f, \"sms\": %d, \"sm_clock_mhz\": %.0f, \"ms\": %.3f}\n", rate,
           rate / p.multiProcessorCount / (clk_khz * 1e3), p.multiProcessorCount, clk_khz / 1e3, best);
    return 0;
}
