// Microbenchmark 2: DMMA (mma.sync.m8n8k4.f64) rate when every product reads DIFFERENT operand registers (10 A x 2 B values per
// lane, as in the reconstruction kernel: 4 chains x 10 k-steps) instead of one constant pair.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, int iters) {
    double c[CH][2], a[10], b[CH][10];
    for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = threadIdx.x * 2e-3 + i; }
    for (int s = 0; s < 10; ++s) { a[s] = threadIdx.x * 1e-4 + 1.0 + s; for (int i = 0; i < CH; ++i) b[i][s] = 0.999 - threadIdx.x * 1e-6 * (s + i + 1); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < 10; ++s)
#pragma unroll
            for (int i = 0; i < CH; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[s]), "d"(b[i][s]));
    }
    double r = 0; for (int i = 0; i < CH; ++i) r += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int CH> void run(double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 400;
    for (int warps : {4, 8, 14, 16}) {
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<CH><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        const double n = 148.0 * warps * iters * 10 * CH;
        printf("chains %d warps/SM %2d: %.3f ms  %.3f DMMA/clk/SM @1.965GHz  %.1f TFLOP/s\n", CH, warps, best, n / 148 / (best * 1e-3 * 1.965e9), n * 512 / (best * 1e-3) / 1e12);
    }
}
int main() { double* out; cudaMalloc(&out, 148 * 1024 * 8); run<2>(out); run<4>(out); return 0; }
