// Microbenchmark: issue rate of legacy mma.sync on sm_100a: m16n8k8 tf32 and m16n8k16 f16/bf16, fp32 accumulate.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int CHAINS>
__global__ void k(float* out, int iters) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) for (int q = 0; q < 4; ++q) c[i][q] = threadIdx.x * 1e-3f + i;
    unsigned a0 = threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, b0 = a0 * 11 + 5, b1 = a0 * 13 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    float r = 0; for (int i = 0; i < CHAINS; ++i) for (int q = 0; q < 4; ++q) r += c[i][q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int KIND> void run(const char* name, double flop_per_mma, float* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    for (int warps : {4, 8, 16}) {
        float best = 1e9f;
        for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<KIND, 6><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
        const double n = 148.0 * warps * iters * 6;
        printf("%s warps/SM %2d: %.3f ms  %.3f mma/clk/SM @1.965GHz  %.1f TFLOP/s\n", name, warps, best, n / 148 / (best * 1e-3 * 1.965e9), n * flop_per_mma / (best * 1e-3) / 1e12);
    }
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    run<0>("m16n8k8  tf32", 2.0 * 16 * 8 * 8, out);
    run<1>("m16n8k16 f16 ", 2.0 * 16 * 8 * 16, out);
    run<2>("m16n8k16 bf16", 2.0 * 16 * 8 * 16, out);
    return 0;
}
