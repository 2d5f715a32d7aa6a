// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) vs FMUL2/FADD2 on sm_100a.  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0) {  // 8 independent FFMA chains
        float a[8];
        for (int i = 0; i < 8; ++i) a[i] = tid * 1e-3f + i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], s, 1.0f);
        float r = 0; for (int i = 0; i < 8; ++i) r += a[i];
        out[tid] = r;
    } else {          // 8 independent FFMA2 chains (16 fp32 lanes of work)
        u64 a[8];
        for (int i = 0; i < 8; ++i) { float2 v = make_float2(tid * 1e-3f + i, tid * 2e-3f + i); a[i] = *reinterpret_cast<u64*>(&v); }
        float2 sv = make_float2(s, s), ov = make_float2(1.f, 1.f);
        const u64 s2 = *reinterpret_cast<u64*>(&sv), o2 = *reinterpret_cast<u64*>(&ov);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], s2, o2);
        float r = 0; for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&a[i]); r += v.x + v.y; }
        out[tid] = r;
    }
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int mode = 0; mode < 2; ++mode) {
            float ms = 0;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 0.999f); else k<1><<<148, warps * 32>>>(out, iters, 0.999f);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            const double inst = 148.0 * warps * iters * 8;          // warp-instructions
            printf("warps/SM %2d %s: %.3f ms  %.2f warp-inst/clk/SM @1.965GHz  (%.1f TFLOP/s)\n", warps, mode ? "FFMA2" : "FFMA ", ms,
                   inst / 148.0 / (ms * 1e-3 * 1.965e9), inst * 32 * 2 * (mode ? 2 : 1) / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
