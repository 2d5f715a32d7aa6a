// Microbenchmark 2: tcgen05.mma issue rate with the CUTLASS-style issue pattern: warp 0, one elected lane, products unrolled
// x8 with compile-time accumulate flags, operand addresses advancing per product (k-steps of a 64-step ring), optional
// kind::f16 (K = 16) to calibrate against the known dense peak.  cycles per product = clock64 delta / products.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gabotorch_b200/csrc -o scripts/micro/umma_rate2 scripts/micro/umma_rate2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace gabo;

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((addr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
template <int F16>
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (F16)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

template <int F16, int ADVANCE>
__global__ void __launch_bounds__(128, 1) k(int M, int N, int groups, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
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
    // instruction descriptor: D f32; A, B = tf32 (2) or f16 (0); K-major both
    const uint32_t fmt = F16 ? 0u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
    // one product reads 32 bytes per row (8 tf32 or 16 f16): two 16-byte k-chunks; no-swizzle canonical layout
    const uint32_t a_sbo = 128, a_lbo = (M / 8) * 128, b_sbo = 128, b_lbo = (N / 8) * 128;
    const uint32_t a_step = ADVANCE ? 2 * a_lbo : 0, b_step = ADVANCE ? 2 * b_lbo : 0;   // next k-step: two k-chunks further
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 64 * 1024;
    if (threadIdx.x < 32) {
        uint32_t elected;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(elected));
        if (elected) {
            const long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint64_t da = desc(a_addr + u * a_step, a_lbo, a_sbo), db = desc(b_addr + u * b_step, b_lbo, b_sbo);
                    mma<F16>(tmem, da, db, idesc, (g | u) ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            mbar_wait(&bar, 0);
            out[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int F16, int ADVANCE>
void run(long long* out) {
    cudaFuncSetAttribute(k<F16, ADVANCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[148];
    const int groups = 256;
    for (int grid : {1, 148})
        for (int M : {64, 128})
            for (int N : {16, 64, 128, 256}) {
                if (ADVANCE && 8 * 2 * (N / 8) * 128 > 128 * 1024) continue;
                k<F16, ADVANCE><<<grid, 128, 200 * 1024>>>(M, N, groups, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("M %d N %d: %s\n", M, N, cudaGetErrorString(e)); return; }
                cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                const double cyc = double(mx) / (groups * 8);
                printf("%s advance %d grid %3d  M %3d N %3d: %7.1f cycles/mma  %.0f FLOP/clk/SM\n", F16 ? "f16 k16 " : "tf32 k8 ",
                       ADVANCE, grid, M, N, cyc, 2.0 * M * N * (F16 ? 16 : 8) / cyc);
            }
}

int main() {
    long long* out;
    cudaMalloc(&out, 148 * 8);
    run<0, 0>(out);
    run<0, 1>(out);
    run<1, 0>(out);
    run<1, 1>(out);
    return 0;
}
