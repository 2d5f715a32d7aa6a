// Microbenchmark: shared -> global bulk store (cp.async.bulk.global.shared::cta) bandwidth: one CTA per SM writes `tiles` tiles of
// `tile_bytes` from shared memory to distinct global addresses, `pieces` bulk copies per tile (issued by `pieces` threads),
// `depth` tiles in flight (wait_group.read depth - 1 before reusing a buffer).  Compare with st.global.cs from registers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace gabo;
template <int DEPTH>
__global__ void __launch_bounds__(256, 1) k(double* out, int tiles, int tile_bytes, int pieces) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < DEPTH * tile_bytes / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = i;
    fence_proxy_async();
    __syncthreads();
    const int piece = tile_bytes / pieces;
    for (int t = 0; t < tiles; ++t) {
        const int b = t % DEPTH;
        if (threadIdx.x < pieces) {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
            char* dst = reinterpret_cast<char*>(out) + (static_cast<size_t>(t) * gridDim.x + blockIdx.x) * tile_bytes + threadIdx.x * piece;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                         "r"(smem_u32(smem + b * tile_bytes + threadIdx.x * piece)), "r"(piece) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x < pieces) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void __launch_bounds__(256, 4) kst(double* out, size_t n) {
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        asm volatile("st.global.cs.v2.f64 [%0], {%1, %1};" ::"l"(out + 2 * i), "d"(1.0) : "memory");
}
template <int DEPTH> void run(double* out, int tile_bytes, int pieces, int tiles) {
    cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * tile_bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0); k<DEPTH><<<148, 256, DEPTH * tile_bytes>>>(out, tiles, tile_bytes, pieces); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("bulk store: tile %6d B in %2d pieces, depth %d: %.3f ms  %.0f GB/s  (%s)\n", tile_bytes, pieces, DEPTH, best,
           148.0 * tiles * tile_bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const int tiles = 14;
    double* out; cudaMalloc(&out, static_cast<size_t>(148) * tiles * 102400 * 4);
    run<1>(out, 102400, 1, tiles); run<1>(out, 102400, 32, tiles); run<2>(out, 102400, 1, tiles); run<2>(out, 102400, 32, tiles);
    run<2>(out, 51200, 16, 2 * tiles); run<4>(out, 51200, 16, 2 * tiles); run<4>(out, 25600, 8, 4 * tiles); run<8>(out, 25600, 8, 4 * tiles);
    const size_t n = static_cast<size_t>(148) * tiles * 102400 / 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); kst<<<148 * 4, 256>>>(out, n); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("st.global.cs.v2.f64 of the same %zu MB: %.3f ms  %.0f GB/s\n", n * 16 >> 20, best, n * 16 / best / 1e6);
    return 0;
}
