// Shared device/host helpers for the gabotorch_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gabo_b200.h"

namespace gabo {

// ---------------------------------------------------------------------------------------------
// Host-side error plumbing: every C-ABI entry returns 0 or a negative code and never throws.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define GABO_REQUIRE(cond, code, ...)        \
    do {                                     \
        if (!(cond)) {                       \
            ::gabo::set_error(__VA_ARGS__);  \
            return (code);                   \
        }                                    \
    } while (0)

int sm_count();

__host__ __device__ __forceinline__ int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// --- TMA (bulk async copy engine) 1-D tile staging: global -> shared, completion on an mbarrier ---
// SASS: UBLKCP.S.G + SYNCS.ARRIVE.TRANS64.  Requirements: 16-byte aligned src/dst, bytes % 16 == 0.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk store shared -> global (the reverse TMA direction): one thread issues, completion tracked by bulk groups
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int kPending>
__device__ __forceinline__ void tma_store_wait_read() {   // at most kPending bulk groups still READING shared memory
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    while (!mbar_try_wait(bar, phase)) {
    }
}

// Stage `bytes` from global into shared.  TMA when the alignment rules hold, cooperative loads otherwise
// (ragged tail tiles).  Every thread of the CTA must call it; `phase` is the caller-tracked mbarrier parity and
// is flipped when the barrier was used.  Ends with the data visible to all threads.
__device__ __forceinline__ void stage_tile(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                           uint32_t& phase) {
    const bool bulk = ((bytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(src_gmem) & 15u) == 0) && bytes > 0;
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(dst_smem, src_gmem, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
    } else {
        const uint32_t n8 = bytes >> 3;  // all our payloads are arrays of 8-byte words
        const double* s = reinterpret_cast<const double*>(src_gmem);
        double* d = reinterpret_cast<double*>(dst_smem);
        for (uint32_t i = threadIdx.x; i < n8; i += blockDim.x) d[i] = s[i];
        __syncthreads();
    }
}

// Streaming (evict-first) stores for write-once outputs: keeps the Gram matrix from thrashing L2.
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(double* p, double v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs4(float* p, float a, float b, float c, float d) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void st_cs2(double* p, double a, double b) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
}

// MUFU wrappers (1-2 ulp, flush-to-zero): the IEEE-rounded sqrtf/exp2f paths cost ~8 extra instructions per call.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2 on the MUFU without the denormal pre-scaling of __log2f (3 extra instructions per call); inputs here are
// eigenvalue-sized (never denormal); a flushed zero gives -inf like the slow path would for 0.
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Exact float -> double widening of a NON-NEGATIVE finite float on the ALU pipe (the converting instruction F2F.F64.F32
// runs on the 16-lane XU pipe, which the per-pair kernels saturate first): exponent re-biased by +896 in front of the
// mantissa shifted down by 3.  A float zero (a kernel value below 1.2e-38 that the MUFU flushed) widens to 2^-127 instead
// of 0 -- the mathematical value exp(-beta d^2) is positive anyway, and the branch-free form saves a compare + select.
__device__ __forceinline__ double widen_nonneg(float v) {
    const unsigned b = __float_as_uint(v);
    return __hiloint2double(static_cast<int>((b >> 3) + 0x38000000u), static_cast<int>(b << 29));
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exp(a) for a <= 0 given in double: range reduction in fp64 (exact to ~1e-16), 2^f on the MUFU in fp32.
// Relative error ~2 ulp(fp32); the argument never sees an fp32 rounding, which keeps K = exp(-beta d^2)
// within 1e-6 of the fp64 reference even for arguments near -14 (K ~ 1e-6).
__device__ __forceinline__ float exp_neg_arg(double a) {
    double t = a * 1.4426950408889634074;
    t = fmax(t, -200.0);
    const double n = rint(t);
    const float f = static_cast<float>(t - n);
    const float r = ex2_approx(f);
    const int ni = static_cast<int>(n);
    // 2^ni for ni in [-200, 0]: split so both factors stay normal.
    const int n1 = max(ni, -100);
    const int n2 = ni - n1;
    return r * __int_as_float((n1 + 127) << 23) * __int_as_float((n2 + 127) << 23);
}

// Packed fp32x2 arithmetic (sm_100a FFMA2 / FMUL2 / FADD2): one issue slot for two fp32 lanes of work.  The FMA pipe
// spends two cycles on it, so peak FLOP/s is unchanged -- the gain is for ISSUE-bound kernels (measured on B200:
// 3.8 FFMA or 2.0 FFMA2 warp-instructions per clock per SM, scripts/micro/ffma2_bench.cu).
__device__ __forceinline__ unsigned long long f2_bits(float2 v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ float2 f2_from(unsigned long long v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)), "l"(f2_bits(c)));
    return f2_from(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_bits(a)), "l"(f2_bits(b)));
    return f2_from(d);
}
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

// asin^2(sqrt(w)) = w * (1 + w * P(w)),  w in [0, 0.5]: degree-8 minimax, relative error 1.8e-7.  With
// w = (1 - |c|) / 2 this is (theta / 2)^2 for the angle theta between two unit vectors with |cos| = |c| -- the squared
// geodesic distance without acos and without the cancellation of acos near c = 1.
__device__ __forceinline__ float asin2_sqrt(float w) {
    float p = 0.3292977809906006f;
    p = fmaf(p, w, -0.3745849132537842f);
    p = fmaf(p, w, 0.28826069831848145f);
    p = fmaf(p, w, -0.03355207294225693f);
    p = fmaf(p, w, 0.07700732350349426f);
    p = fmaf(p, w, 0.07967597246170044f);
    p = fmaf(p, w, 0.1143670305609703f);
    p = fmaf(p, w, 0.17777620255947113f);
    p = fmaf(p, w, 0.3333333432674408f);
    return w * fmaf(w, p, 1.0f);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace gabo
