// Pipe-rate probe (exploration, not product): DFMA vs IMAD.WIDE throughput on this GPU and whether
// the two pipes overlap when interleaved in one warp.  nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE> __global__ void probe(double *out, int iters, double seed) {
    double a[8], m = seed + threadIdx.x * 1e-9;
    uint64_t x[8];
    uint32_t b = (uint32_t)(seed * 7) + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed + k; x[k] = threadIdx.x + k; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0 || MODE == 2) a[k] = fma(a[k], m, a[k]);
            if (MODE == 1 || MODE == 2) x[k] = (uint64_t)(uint32_t)x[k] * b + x[k];
        }
    }
    double s = 0; uint64_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { s += a[k]; t ^= x[k]; }
    if (s == 12345.678 || t == 0x123456789abcdefull) out[0] = s + (double)t;
}

template <int MODE> double run(int sms) {
    double *d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float ms = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        probe<MODE><<<blocks, threads>>>(d, iters, 1.000001 + rep);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaFree(d);
    return (double)blocks * threads * iters * 8 / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double clk = p.clockRate * 1e3;
    double d = run<0>(sms), i = run<1>(sms), m = run<2>(sms);
    printf("%s: %d SMs, %.0f MHz\n", p.name, sms, clk / 1e6);
    printf("DFMA only      : %.3e /s  = %.1f per clk per SM\n", d, d / clk / sms);
    printf("IMAD.WIDE only : %.3e /s  = %.1f per clk per SM\n", i, i / clk / sms);
    printf("interleaved 1:1: %.3e pairs/s = %.1f DFMA + %.1f IMAD.WIDE per clk per SM\n", m, m / clk / sms, m / clk / sms);
    return 0;
}
