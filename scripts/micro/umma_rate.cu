// Microbenchmark: issue rate of tcgen05.mma.kind::tf32 (SS mode: both operands from shared memory) on sm_100a as a function of
// the instruction shape M x N x 8 and of the shared-memory layout (no swizzle / 32B / 128B swizzle), with the accumulators
// rotating over `nacc` TMEM regions.  One CTA per SM, thread 0 issues `iters` products back to back, one commit at the end;
// cycles per product = clock64 delta / iters.  Operand data are zeros: only the rate is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gabotorch_b200/csrc -o scripts/micro/umma_rate scripts/micro/umma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace gabo;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return static_cast<uint64_t>((addr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout) << 61);
}

__global__ void __launch_bounds__(128, 1) k(int M, int N, int layout, int nacc, int iters, int a_tmem, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    fence_proxy_async();
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
    // K = 8 tf32 = 32 bytes per row.  no swizzle: core matrices 8 rows x 16 B, SBO = 128 B between 8-row groups, LBO between the
    // two 16-byte k-chunks; swizzled: rows of 32 B (32B swizzle) or 128 B (128B swizzle), SBO = 8 rows
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    if (layout == 0) { a_sbo = 128; a_lbo = (M / 8) * 128; b_sbo = 128; b_lbo = (N / 8) * 128; }
    else if (layout == 6) { a_sbo = 256; a_lbo = 0; b_sbo = 256; b_lbo = 0; }
    else { a_sbo = 1024; a_lbo = 0; b_sbo = 1024; b_lbo = 0; }
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32 * 1024;
    const uint64_t da = desc(a_addr, a_lbo, a_sbo, layout), db = desc(b_addr, b_lbo, b_sbo, layout);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + static_cast<uint32_t>((i % nacc) * N) % 512u;
            if (a_tmem) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(tmem + 448u), "l"(db), "r"(idesc), "r"(i >= nacc ? 1 : 0) : "memory");
            } else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(i >= nacc ? 1 : 0) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int iters = 2048;
    long long h[148];
    for (int a_tmem : {0, 1})
        for (int layout : {0, 6, 2})
            for (int M : {64, 128})
                for (int N : {16, 32, 64, 128, 256}) {
                    for (int nacc : {1, 2}) {
                        if (nacc * N > 448) continue;
                        k<<<148, 128, 160 * 1024>>>(M, N, layout, nacc, iters, a_tmem, out);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("M %d N %d layout %d: %s\n", M, N, layout, cudaGetErrorString(e)); return 1; }
                        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                        long long mx = 0;
                        for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
                        const double cyc = double(mx) / iters;
                        printf("A %s layout %d  M %3d N %3d nacc %d: %7.1f cycles/mma  %.0f FLOP/clk/SM  operand B/clk %.0f\n",
                               a_tmem ? "tmem" : "smem", layout, M, N, nacc, cyc, 2.0 * M * N * 8 / cyc, ((a_tmem ? 0 : M) + N) * 32.0 / cyc);
                    }
                }
    return 0;
}
