// Microbenchmark: HBM read bandwidth of a persistent CTA streaming tiles through an S-stage TMA (cp.async.bulk) ring,
// versus plain coalesced float4 loads.  Informs the stage count / tile size of nested_project_kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(b)), "r"(c)); }
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void tma1d(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(su32(d)), "l"(s), "r"(n), "r"(su32(b)) : "memory");
}
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(su32(b)), "r"(ph) : "memory");
}
template <int STAGES>
__global__ void __launch_bounds__(256) stream_tma(const float* x, int64_t tiles, int tile_bytes, float* out, int work) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * tile_bytes);
    if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int tf = tile_bytes / 4;
    if (threadIdx.x == 0) for (int s = 0; s < STAGES; ++s) { int64_t t = blockIdx.x + (int64_t)s * gridDim.x; if (t < tiles) { expect_tx(&bars[s], tile_bytes); tma1d(sm + (size_t)s * tile_bytes, x + t * tf, tile_bytes, &bars[s]); } }
    float acc = 0.f; uint32_t ph = 0; int it = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        mwait(&bars[s], (ph >> s) & 1u); ph ^= 1u << s;
        const float* a = reinterpret_cast<const float*>(sm + (size_t)s * tile_bytes);
        for (int w = 0; w < work; ++w) for (int e = threadIdx.x; e < tf; e += 256) acc += a[e];   // `work` passes over the tile
        __syncthreads();
        if (threadIdx.x == 0) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); int64_t tn = t + (int64_t)STAGES * gridDim.x; if (tn < tiles) { expect_tx(&bars[s], tile_bytes); tma1d(sm + (size_t)s * tile_bytes, x + tn * tf, tile_bytes, &bars[s]); } }
    }
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void __launch_bounds__(256) stream_ldg(const float4* x, int64_t n4, float* out) {
    float acc = 0.f;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) { float4 v = __ldcs(x + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 12345.678f) out[0] = acc;
}
template <int S> float run_tma(const float* x, int64_t bytes, int tile_bytes, int ctas_per_sm, float* out, int work) {
    size_t smem = (size_t)S * tile_bytes + 8 * S + 64;
    if (smem * ctas_per_sm > 226 * 1024) return -1.f;
    cudaFuncSetAttribute(stream_tma<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float best = 1e9f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); stream_tma<S><<<148 * ctas_per_sm, 256, smem>>>(x, bytes / tile_bytes, tile_bytes, out, work); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    if (cudaGetLastError() != cudaSuccess) return -2.f;
    return best;
}
int main() {
    const int64_t bytes = 860160LL * 1024;   // 881 MB (= 2^20 rows of 840 bytes)
    float *x, *out; cudaMalloc(&x, bytes); cudaMalloc(&out, 4); cudaMemset(x, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int g : {148 * 4, 148 * 8, 148 * 16}) { float best = 1e9f; for (int r = 0; r < 5; ++r) { cudaEventRecord(e0); stream_ldg<<<g, 256>>>((const float4*)x, bytes / 16, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; } printf("ldg float4 grid %5d: %.3f ms  %.0f GB/s\n", g, best, bytes / best / 1e6); }
    for (int work : {1, 4}) for (int rows : {16, 32, 64}) { const int tb = rows * 840;
        for (int cps : {1, 2}) {
            float a = run_tma<2>(x, bytes, tb, cps, out, work), b = run_tma<3>(x, bytes, tb, cps, out, work), c = run_tma<4>(x, bytes, tb, cps, out, work), d = run_tma<6>(x, bytes, tb, cps, out, work);
            printf("work %d tile %2d rows (%5d B) ctas/SM %d: S=2 %.0f  S=3 %.0f  S=4 %.0f  S=6 %.0f GB/s\n", work, rows, tb, cps, a > 0 ? bytes / a / 1e6 : 0, b > 0 ? bytes / b / 1e6 : 0, c > 0 ? bytes / c / 1e6 : 0, d > 0 ? bytes / d / 1e6 : 0);
        } }
    return 0;
}
