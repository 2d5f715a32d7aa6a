// Microbenchmark: HBM WRITE-only and READ-only bandwidth vs the read+write copy figure of MEASURED_PEAKS.json.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void __launch_bounds__(256) wr(float4* p, int64_t n4) {
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        if (MODE == 0) p[i] = v; else if (MODE == 1) __stcs(p + i, v); else __stwt(p + i, v);
    }
}
__global__ void __launch_bounds__(256) rd(const float4* p, int64_t n4, float* out) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void __launch_bounds__(256) cp(const float4* a, float4* b, int64_t n4) {
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) __stcs(b + i, __ldcs(a + i));
}
int main() {
    const int64_t bytes = 4LL << 30; float4 *a, *b; float* out;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&out, 4); cudaMemset(a, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto t = [&](auto f) { float best = 1e9f; for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; } return best; };
    for (int g : {148 * 4, 148 * 8, 148 * 32}) {
        float w0 = t([&] { wr<0><<<g, 256>>>(b, bytes / 16); }), w1 = t([&] { wr<1><<<g, 256>>>(b, bytes / 16); }), w2 = t([&] { wr<2><<<g, 256>>>(b, bytes / 16); });
        float r = t([&] { rd<<<g, 256>>>(a, bytes / 16, out); }), c = t([&] { cp<<<g, 256>>>(a, b, bytes / 16); });
        printf("grid %5d: write st %.0f  st.cs %.0f  st.wt %.0f GB/s | read %.0f GB/s | copy (r+w) %.0f GB/s\n", g, bytes / w0 / 1e6, bytes / w1 / 1e6, bytes / w2 / 1e6, bytes / r / 1e6, 2.0 * bytes / c / 1e6);
    }
    float m = t([&] { cudaMemsetAsync(b, 0, bytes); });
    printf("cudaMemsetAsync: %.0f GB/s\n", bytes / m / 1e6);
    return 0;
}
