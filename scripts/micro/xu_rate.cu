// Microbenchmark: throughput of the XU-pipe instructions the Gram kernels use (MUFU.EX2 / LG2 / RSQ / SQRT / RCP,
// F2F.F32.F64, F2F.F64.F32) and of their ALU-pipe replacements, in warp-instructions per clock per SM (sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float s) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float a[8];
    double d[8];
    for (int i = 0; i < 8; ++i) { a[i] = 1.0f + tid * 1e-6f + i * 0.01f; d[i] = a[i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 3) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 5) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a[i]) : "d"(d[i])); d[i] = __hiloint2double(__double2hiint(d[i]) ^ __float_as_int(a[i]) & 1, __double2loint(d[i])); }
            if (MODE == 6) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(a[i])); a[i] = __int_as_float(__float_as_int(a[i]) ^ (__double2hiint(d[i]) & 1)); }
            if (MODE == 7) { const int hi = max(__double2hiint(d[i]), 0x3CD20000);      // ALU replacement of mode 5
                             a[i] = __uint_as_float(__funnelshift_l((unsigned)__double2loint(d[i]), (unsigned)hi, 3) + 0x40000000u);
                             d[i] = __hiloint2double(__double2hiint(d[i]) ^ __float_as_int(a[i]) & 1, __double2loint(d[i])); }
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) r += a[i] + (float)d[i];
    out[tid] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    const char* names[] = {"MUFU.EX2", "MUFU.LG2", "MUFU.RSQ", "MUFU.SQRT", "MUFU.RCP", "F2F.F32.F64 (+2 ALU)", "F2F.F64.F32 (+2 ALU)", "ALU cvt f64->f32 (+2 ALU)"};
    for (int mode = 0; mode < 8; ++mode) {
        float ms = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            switch (mode) {
                case 0: k<0><<<148, 1024>>>(out, iters, 1.f); break;
                case 1: k<1><<<148, 1024>>>(out, iters, 1.f); break;
                case 2: k<2><<<148, 1024>>>(out, iters, 1.f); break;
                case 3: k<3><<<148, 1024>>>(out, iters, 1.f); break;
                case 4: k<4><<<148, 1024>>>(out, iters, 1.f); break;
                case 5: k<5><<<148, 1024>>>(out, iters, 1.f); break;
                case 6: k<6><<<148, 1024>>>(out, iters, 1.f); break;
                case 7: k<7><<<148, 1024>>>(out, iters, 1.f); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double inst = 148.0 * 32 * iters * 8;   // warp-level operations of the measured kind
        printf("%-28s %.3f ms  %.3f warp-ops/clk/SM @1.965GHz (= %.1f lanes/clk/SM)\n", names[mode], ms,
               inst / 148.0 / (ms * 1e-3 * 1.965e9), 32 * inst / 148.0 / (ms * 1e-3 * 1.965e9));
    }
    return 0;
}
