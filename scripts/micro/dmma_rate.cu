// Microbenchmark: mma.sync.m8n8k4.f64 (DMMA) issue rate on sm_100a vs DFMA.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = threadIdx.x * 2e-3 + i; }
    double a = threadIdx.x * 1e-4 + 1.0, b = 0.999 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
            else { asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[i][0]) : "d"(a), "d"(b)); asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[i][1]) : "d"(a), "d"(b)); }
        }
    }
    double r = 0; for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind) for (int warps : {4, 8, 16}) {
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); if (kind == 0) k<0><<<148, warps * 32>>>(out, iters); else k<1><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        const double n = 148.0 * warps * iters * (kind == 0 ? 8 : 16);
        const double flop = n * (kind == 0 ? 2.0 * 8 * 8 * 4 : 2.0 * 32);
        printf("%s warps/SM %2d: %.3f ms  %.3f inst/clk/SM @1.965GHz  %.1f TFLOP/s\n", kind == 0 ? "DMMA m8n8k4" : "DFMA       ", warps, best, n / 148 / (best * 1e-3 * 1.965e9), flop / (best * 1e-3) / 1e12);
    }
    return 0;
}
