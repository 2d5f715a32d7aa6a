// Fused sphere Gram: out[i,j] = f(acos(clamp(<x1_i, x2_j>))) in one pass (G1 + G2 of SURVEY.md section 8).
//
// Replaces sphere_distance_torch (BoManifolds/Riemannian_utils/sphere_utils_torch.py:12-55: two materialised
// (N1,N2,D) broadcasts + an N1*N2-batch bmm + clamp + acos) and the d^2 / exp passes of
// SphereGaussianKernel.forward (kernel_utils/kernels_sphere.py:89-94).
//
// Roofline: HBM write, 4 B (fp32 out) per pair; inputs are O(N*D) and stay in L2.
// Layout: persistent CTAs walk a contiguous range of (column-block, row-block) tiles; a thread owns 4 consecutive
// columns (x2 points held in registers as fp64) and streams down the rows of the tile, whose x1 points are staged in
// shared memory by the TMA bulk-copy engine (double-buffered, mbarrier completion) and read as broadcasts.
// Each row a warp writes 512 contiguous bytes with streaming float4 stores.
//
// Numerics: the inner product is accumulated in fp64 (D DFMAs), so 1-|c| is exact and the reference's
// clamp epsilon (1e-15) keeps its meaning; everything after that is fp32:
//   w = (1-|c|)/2,  r^2 = 4 asin^2(sqrt w) = 4w(1 + w P(w))   (degree-8 minimax P on [0, 1/2], no sqrt / acos needed),
//   d^2 = r^2 for c >= 0,  (pi - r)^2 for c < 0,   K = 2^(d^2 * (-beta log2 e)).
// Relative error of d^2 is ~1.8e-7, of K about (beta d^2) * 2e-7 + 1e-6.
#include <cmath>

#include "common.cuh"

namespace gabo {

namespace {

// Tuning knobs (defaults = the measured best on B200; scripts/micro/sphere_variants.cu rebuilds this file with others)
#ifndef GABO_SG_TILEM
#define GABO_SG_TILEM 32
#endif
// measured at N = 32768, D = 3 (profiles/r02_sphere_variants.log): 4 rows in flight at <= 102 registers (5 CTAs per SM)
// gives 4494 / 6348 GB/s (fp32 / fp64 out) against 4396 / 5991 for 2 rows at 56 registers; Estrin's scheme is slower
// (the kernel is bound by instruction issue, not by the latency of the Horner chain)
#ifndef GABO_SG_UNROLL
#define GABO_SG_UNROLL 4
#endif
#ifndef GABO_SG_MINBLOCKS
#define GABO_SG_MINBLOCKS 5
#endif
#define GABO_SG_BOUNDS __launch_bounds__(kThreads, GABO_SG_MINBLOCKS)
constexpr int kThreads = 128;
constexpr int kVec = 4;                    // columns per thread
constexpr int kTileN = kThreads * kVec;    // 512 columns per tile
constexpr int kTileM = GABO_SG_TILEM;      // rows per tile
constexpr int kUnroll = GABO_SG_UNROLL;    // rows in flight per thread (fast row loop)

struct TailParams {
    float k_hi, k_lo;  // k = -param * log2(e) split in two floats (Gauss / Laplace)
    // Gaussian kind, fused form: the exponent t = k d^2 of K = 2^t straight from w2 = 1 - |c|, with k folded into the
    // polynomial on the host:  near side  t = w2 (2k + w2 sum_j (k a_j / 2^j) w2^j)  (= k * 4 asin^2(sqrt(w2 / 2))),
    // far side (c < 0)  t = -(sqrt|k| pi - sqrt(-t_near))^2.  Saves the three scalings (x 1/2, x 4, x k) per pair.
    float pc[10];      // pc[0..8] = k a_j / 2^j for j = 8 .. 0,  pc[9] = 2k
    float skpi_hi, skpi_lo;  // sqrt|k| * pi in two floats
};

// d^2 (and optionally d) of the geodesic distance from the fp64 inner product.
// The reference clamps c to [-1+1e-15, 1-1e-15] (sphere_utils_torch.py:53); 1-|c| is exact in fp64, so the same clamp is
// applied to w2 = 1-|c| AFTER the conversion to fp32 (one compare+select instead of two fp64 min/max, which cost ~12
// instructions per pair on sm_100a).  kClampW = 1 - (1 - 1e-15) evaluated in fp64.  A NaN inner product stays NaN.
template <int KIND>
__device__ __forceinline__ float tail(double c, const TailParams& tp) {
    constexpr float kClampW = 9.992007221626409e-16f;
    float w2 = static_cast<float>(1.0 - fabs(c));
    w2 = (w2 < kClampW) ? kClampW : w2;
    const float w = 0.5f * w2;
    const float r2 = 4.0f * asin2_sqrt(w);           // squared distance to the nearer pole (+-x)
    const bool neg = __double2hiint(c) < 0;           // sign bit of c (c = -0.0 gives w = 1/2: both branches agree)
    if (KIND == GABO_KIND_GAUSS) {
        float d2 = r2;
        if (neg) {
            const float r = sqrt_approx(r2);
            const float d = (3.14159274101257324f - r) + (-8.74227765734758577e-8f);
            d2 = d * d;
        }
        const float t = fmaf(d2, tp.k_hi, d2 * tp.k_lo);
        return ex2_approx(t);
    } else {
        const float r = sqrt_approx(r2);
        const float d = neg ? (3.14159274101257324f - r) + (-8.74227765734758577e-8f) : r;
        if (KIND == GABO_KIND_DIST) return d;
        const float t = fmaf(d, tp.k_hi, d * tp.k_lo);
        return ex2_approx(t);
    }
}

// Two pairs at a time on the packed fp32x2 pipe: the polynomial and the exponent scaling are the bulk of the fp32
// instructions of a pair, and the kernel is issue-bound (ncu: 78 % issue-active, HBM 40 % before this change).
__device__ __forceinline__ float2 asin2_sqrt2(float2 w) {
    float2 p = splat2(0.3292977809906006f);
    p = fma2(p, w, splat2(-0.3745849132537842f));
    p = fma2(p, w, splat2(0.28826069831848145f));
    p = fma2(p, w, splat2(-0.03355207294225693f));
    p = fma2(p, w, splat2(0.07700732350349426f));
    p = fma2(p, w, splat2(0.07967597246170044f));
    p = fma2(p, w, splat2(0.1143670305609703f));
    p = fma2(p, w, splat2(0.17777620255947113f));
    p = fma2(p, w, splat2(0.3333333432674408f));
    return mul2(w, fma2(w, p, splat2(1.0f)));
}

__device__ __forceinline__ float far_side(float r2) {   // (pi - sqrt(r2)), the distance when <x,y> < 0
    const float r = sqrt_approx(r2);
    return (3.14159274101257324f - r) + (-8.74227765734758577e-8f);
}

// fp64 -> fp32 of w2 = 1 - |c| WITHOUT the conversion unit.  ncu on the N = 32768 launch: the XU pipe (MUFU and
// F2F.F32.F64, 16 lanes per SM) was the busiest unit of the fp32-output kernel -- 3 XU operations per pair (conversion,
// far-side root, 2^t) cap the kernel at 5.3 pairs per clock per SM, below what HBM takes (5.6).  The conversion is done
// on the ALU pipe instead: clamp the high word (signed compare: negative, zero and tiny values all land on the
// reference's clamp 1 - (1 - 1e-15) = 0x3CD20000'00000000), then the float pattern is the double's exponent re-biased
// (E - 896; for E in [897, 1023] that is "clear the two top bits of the 9 low exponent bits") in front of the top 23
// mantissa bits: one funnel shift and one add.  Truncation instead of rounding: relative error < 1.2e-7 on w2.
// Only for FINITE inputs: a NaN would come out as 1.5, so the callers take this path only for tiles whose points they
// have checked (warp-uniform), everything else runs the converting form above.
__device__ __forceinline__ float w2_bits(double c) {
    const double w = 1.0 - fabs(c);
    const int hi = max(__double2hiint(w), 0x3CD20000);
    return __uint_as_float(__funnelshift_l(static_cast<unsigned>(__double2loint(w)), static_cast<unsigned>(hi), 3) +
                           0x40000000u);
}

// Same arithmetic as tail<KIND>() for the two inner products (c0, c1).  kFast: see w2_bits.
template <int KIND, bool kFast = false>
__device__ __forceinline__ float2 tail2(double c0, double c1, const TailParams& tp) {
    constexpr float kClampW = 9.992007221626409e-16f;
    float w0, w1;
    if (kFast) {
        w0 = w2_bits(c0);
        w1 = w2_bits(c1);
    } else {
        w0 = static_cast<float>(1.0 - fabs(c0));
        w1 = static_cast<float>(1.0 - fabs(c1));
        w0 = (w0 < kClampW) ? kClampW : w0;
        w1 = (w1 < kClampW) ? kClampW : w1;
    }
    if (KIND == GABO_KIND_GAUSS) {   // fused exponent (see TailParams): 10 packed FMAs + the far-side fix-up
        const float2 w = make_float2(w0, w1);
#ifdef GABO_SG_ESTRIN
        // Estrin's scheme: 9 FMA + 4 MUL at depth 6 instead of 9 FMA + 1 MUL at depth 10 (coefficient of w^j = pc[9-j])
        const float2 wa = mul2(w, w), wb = mul2(wa, wa), wc = mul2(wb, wb);
        const float2 e0 = fma2(splat2(tp.pc[8]), w, splat2(tp.pc[9])), e1 = fma2(splat2(tp.pc[6]), w, splat2(tp.pc[7]));
        const float2 e2 = fma2(splat2(tp.pc[4]), w, splat2(tp.pc[5])), e3 = fma2(splat2(tp.pc[2]), w, splat2(tp.pc[3]));
        const float2 e4 = fma2(splat2(tp.pc[0]), w, splat2(tp.pc[1]));
        const float2 f0 = fma2(e1, wa, e0), f1 = fma2(e3, wa, e2);
        const float2 p = fma2(e4, wc, fma2(f1, wb, f0));
#else
        float2 p = splat2(tp.pc[0]);
#pragma unroll
        for (int j = 1; j <= 9; ++j) p = fma2(p, w, splat2(tp.pc[j]));
#endif
        const float2 tn = mul2(w, p);                                         // k d^2 on the near side (<= 0)
        const float2 s = make_float2(sqrt_approx(-tn.x), sqrt_approx(-tn.y));
        const float2 u = add2(sub2(splat2(tp.skpi_hi), s), splat2(tp.skpi_lo));  // sqrt|k| (pi - r)
        const float2 tf = mul2(make_float2(-u.x, -u.y), u);
        const float tx = (__double2hiint(c0) < 0) ? tf.x : tn.x;
        const float ty = (__double2hiint(c1) < 0) ? tf.y : tn.y;
        return make_float2(ex2_approx(tx), ex2_approx(ty));
    }
    const float2 r2 = mul2(splat2(4.0f), asin2_sqrt2(mul2(splat2(0.5f), make_float2(w0, w1))));
    const bool n0 = __double2hiint(c0) < 0, n1 = __double2hiint(c1) < 0;
    float2 v;
    if (KIND == GABO_KIND_GAUSS) {
        v = r2;
        if (n0) { const float d = far_side(r2.x); v.x = d * d; }
        if (n1) { const float d = far_side(r2.y); v.y = d * d; }
    } else {
        v.x = n0 ? far_side(r2.x) : sqrt_approx(r2.x);
        v.y = n1 ? far_side(r2.y) : sqrt_approx(r2.y);
        if (KIND == GABO_KIND_DIST) return v;
    }
    const float2 t = fma2(v, splat2(tp.k_hi), mul2(v, splat2(tp.k_lo)));
    return make_float2(ex2_approx(t.x), ex2_approx(t.y));
}

template <typename OutT>
__device__ __forceinline__ void store4(OutT* p, const float (&v)[kVec], int valid, bool vec_ok);

template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[kVec], int valid, bool vec_ok) {
    if (vec_ok && valid == kVec) {
        st_cs4(p, v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int q = 0; q < kVec; ++q)
            if (q < valid) st_cs(p + q, v[q]);
    }
}
template <>
__device__ __forceinline__ void store4<double>(double* p, const float (&v)[kVec], int valid, bool vec_ok) {
    if (vec_ok && valid == kVec) {
        st_cs2(p, static_cast<double>(v[0]), static_cast<double>(v[1]));
        st_cs2(p + 2, static_cast<double>(v[2]), static_cast<double>(v[3]));
    } else {
#pragma unroll
        for (int q = 0; q < kVec; ++q)
            if (q < valid) st_cs(p + q, static_cast<double>(v[q]));
    }
}

// Two consecutive columns per store.  The fixed-dimension kernel gives a thread columns {2 lane, 2 lane + 1} of each
// 64-column half of its warp's 128-column group, so that ONE warp store instruction covers a contiguous 256 B (fp32)
// or 512 B (fp64) run.  (With 4 consecutive columns per thread the fp64 result needed two 16-byte stores per thread
// that each touched only half of every 32-byte sector: the fp64-output Gram -- the dtype the reference API returns --
// ran at 55 % of HBM where the fp32-output one reached 59 %.)
template <typename OutT>
__device__ __forceinline__ void store2(OutT* p, float v0, float v1, int valid, bool vec_ok);
template <>
__device__ __forceinline__ void store2<float>(float* p, float v0, float v1, int valid, bool vec_ok) {
    if (vec_ok && valid == 2) {
        __stcs(reinterpret_cast<float2*>(p), make_float2(v0, v1));
    } else {
        if (valid > 0) st_cs(p, v0);
        if (valid > 1) st_cs(p + 1, v1);
    }
}
template <>
__device__ __forceinline__ void store2<double>(double* p, float v0, float v1, int valid, bool vec_ok) {
    if (vec_ok && valid == 2) {
        st_cs2(p, static_cast<double>(v0), static_cast<double>(v1));
    } else {
        if (valid > 0) st_cs(p, static_cast<double>(v0));
        if (valid > 1) st_cs(p + 1, static_cast<double>(v1));
    }
}

template <int D, typename OutT, int KIND>
__global__ void GABO_SG_BOUNDS sphere_gram_kernel(const double* __restrict__ x1, int64_t n1,
                                                               const double* __restrict__ x2, int64_t n2,
                                                               TailParams tp, OutT* __restrict__ out, int64_t ld_out,
                                                               int64_t tiles_i, int64_t tiles_total, bool vec_ok) {
    __shared__ __align__(16) double xs[2][kTileM * D];
    __shared__ __align__(8) uint64_t bar[2];

    // contiguous tile range for this CTA; tile id t = jb * tiles_i + ib (rows fastest: x2 registers are reused)
    const int64_t per = (tiles_total + gridDim.x - 1) / gridDim.x;
    const int64_t t_begin = static_cast<int64_t>(blockIdx.x) * per;
    const int64_t t_end = min(t_begin + per, tiles_total);
    if (t_begin >= t_end) return;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t phase[2] = {0u, 0u};
    auto issue = [&](int64_t t, int buf) {
        const int64_t ib = t % tiles_i;
        const int64_t i0 = ib * kTileM;
        const int rows = static_cast<int>(imin(kTileM, n1 - i0));
        const uint32_t bytes = static_cast<uint32_t>(rows) * D * sizeof(double);
        const double* src = x1 + i0 * D;
        const bool bulk = ((bytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
        if (bulk) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar[buf], bytes);
                tma_load_1d(&xs[buf][0], src, bytes, &bar[buf]);
            }
        } else {
            for (int e = threadIdx.x; e < rows * D; e += kThreads) xs[buf][e] = src[e];
        }
        return bulk;
    };

    bool bulk_cur = issue(t_begin, 0);
    int64_t jb_loaded = -1;
    double b[kVec][D];
    // fp64 output (the dtype the reference API returns): two-column stores, see store2; fp32 output: one float4 store
    constexpr bool kPairLayout = sizeof(OutT) == 8;
    int valid = 0, valid_hi = 0;
    int64_t j_first = 0;
    bool b_finite = true;   // the x2 columns of this thread are finite (fast row loop, see below)

    for (int64_t t = t_begin; t < t_end; ++t) {
        const int buf = static_cast<int>((t - t_begin) & 1);
        const int64_t jb = t / tiles_i;
        const int64_t ib = t % tiles_i;
        const int64_t i0 = ib * kTileM;
        const int rows = static_cast<int>(imin(kTileM, n1 - i0));

        // the buffer we are about to refill was last read two tiles ago
        __syncthreads();
        bool bulk_next = false;
        if (t + 1 < t_end) bulk_next = issue(t + 1, buf ^ 1);

        if (jb != jb_loaded) {
            jb_loaded = jb;
            if (kPairLayout) {   // columns of this thread: j_first + {0, 1} and j_first + 64 + {0, 1}
                j_first = jb * kTileN + static_cast<int64_t>(threadIdx.x >> 5) * 128 + 2 * (threadIdx.x & 31);
                valid = static_cast<int>(imax(0, imin(2, n2 - j_first)));
                valid_hi = static_cast<int>(imax(0, imin(2, n2 - (j_first + 64))));
            } else {             // four consecutive columns: one 16-byte store per row
                j_first = jb * kTileN + static_cast<int64_t>(threadIdx.x) * kVec;
                valid = static_cast<int>(imax(0, imin(kVec, n2 - j_first)));
            }
#pragma unroll
            for (int q = 0; q < kVec; ++q) {
                const int64_t j = imin(j_first + (kPairLayout ? (q >> 1) * 64 + (q & 1) : q), n2 - 1);
#pragma unroll
                for (int k = 0; k < D; ++k) b[q][k] = __ldg(x2 + j * D + k);
            }
            b_finite = true;
#pragma unroll
            for (int q = 0; q < kVec; ++q)
#pragma unroll
                for (int k = 0; k < D; ++k)
                    b_finite = b_finite && ((__double2hiint(b[q][k]) & 0x7ff00000) != 0x7ff00000);
        }

        if (bulk_cur) {
            mbar_wait(&bar[buf], phase[buf]);
            phase[buf] ^= 1u;
        } else {
            __syncthreads();
        }

        // Warp-uniform choice of the row loop.  Fast form: every lane of the warp owns only valid columns, vector stores
        // are possible and neither the x2 columns of this warp nor the rows of this tile hold a NaN / Inf (checked here:
        // each lane looks at rows * D / 32 staged values) -- it converts on the ALU pipe (w2_bits) and has no per-row
        // edge branches.  Everything else (ragged edges, unaligned output, non-finite points) takes the general form.
        bool fast = vec_ok && b_finite && (kPairLayout ? (valid == 2 && valid_hi == 2) : (valid == kVec));
        {
            bool row_ok = true;
            for (int e = threadIdx.x & 31; e < rows * D; e += 32)
                row_ok = row_ok && ((__double2hiint(xs[buf][e]) & 0x7ff00000) != 0x7ff00000);
            fast = __all_sync(0xffffffffu, fast && row_ok);
        }
        OutT* orow = out + i0 * ld_out + j_first;
        if (fast) {
#pragma unroll kUnroll
            for (int i = 0; i < rows; ++i, orow += ld_out) {
                double a[D];
#pragma unroll
                for (int k = 0; k < D; ++k) a[k] = xs[buf][i * D + k];
                double c[kVec];
#pragma unroll
                for (int q = 0; q < kVec; ++q) {
                    c[q] = a[0] * b[q][0];
#pragma unroll
                    for (int k = 1; k < D; ++k) c[q] = fma(a[k], b[q][k], c[q]);
                }
                const float2 v01 = tail2<KIND, true>(c[0], c[1], tp), v23 = tail2<KIND, true>(c[2], c[3], tp);
                if (kPairLayout) {
                    st_cs2(reinterpret_cast<double*>(orow), static_cast<double>(v01.x), static_cast<double>(v01.y));
                    st_cs2(reinterpret_cast<double*>(orow) + 64, static_cast<double>(v23.x), static_cast<double>(v23.y));
                } else {
                    st_cs4(reinterpret_cast<float*>(orow), v01.x, v01.y, v23.x, v23.y);
                }
            }
        } else if (valid > 0) {
#pragma unroll 1
            for (int i = 0; i < rows; ++i, orow += ld_out) {
                double a[D];
#pragma unroll
                for (int k = 0; k < D; ++k) a[k] = xs[buf][i * D + k];
                double c[kVec];
#pragma unroll
                for (int q = 0; q < kVec; ++q) {
                    c[q] = a[0] * b[q][0];
#pragma unroll
                    for (int k = 1; k < D; ++k) c[q] = fma(a[k], b[q][k], c[q]);
                }
                const float2 v01 = tail2<KIND>(c[0], c[1], tp), v23 = tail2<KIND>(c[2], c[3], tp);
                if (kPairLayout) {
                    store2<OutT>(orow, v01.x, v01.y, valid, vec_ok);
                    store2<OutT>(orow + 64, v23.x, v23.y, valid_hi, vec_ok);
                } else {
                    const float v[kVec] = {v01.x, v01.y, v23.x, v23.y};
                    store4<OutT>(orow, v, valid, vec_ok);
                }
            }
        }
        bulk_cur = bulk_next;
    }
}

// Generic ambient dimension (D > 16, up to GABO_MAX_SPHERE_DIM): one thread per 4 columns, x2 re-read through L1.
template <typename OutT, int KIND>
__global__ void __launch_bounds__(kThreads) sphere_gram_generic_kernel(const double* __restrict__ x1, int64_t n1,
                                                                       const double* __restrict__ x2, int64_t n2,
                                                                       int dim, TailParams tp, OutT* __restrict__ out,
                                                                       int64_t ld_out, int64_t tiles_i,
                                                                       int64_t tiles_total) {
    extern __shared__ __align__(16) double xs_dyn[];  // kTileM * dim
    for (int64_t t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        const int64_t jb = t / tiles_i;
        const int64_t ib = t % tiles_i;
        const int64_t i0 = ib * kTileM;
        const int rows = static_cast<int>(imin(kTileM, n1 - i0));
        __syncthreads();
        for (int e = threadIdx.x; e < rows * dim; e += kThreads) xs_dyn[e] = x1[i0 * dim + e];
        __syncthreads();
        const int64_t j_first = jb * kTileN + static_cast<int64_t>(threadIdx.x) * kVec;
        const int valid = static_cast<int>(imax(0, imin(kVec, n2 - j_first)));
        if (valid <= 0) continue;
        for (int i = 0; i < rows; ++i) {
            double c[kVec] = {0.0, 0.0, 0.0, 0.0};
            for (int k = 0; k < dim; ++k) {
                const double a = xs_dyn[i * dim + k];
#pragma unroll
                for (int q = 0; q < kVec; ++q) {
                    const int64_t j = imin(j_first + q, n2 - 1);
                    c[q] = fma(a, __ldg(x2 + j * dim + k), c[q]);
                }
            }
            float v[kVec];
#pragma unroll
            for (int q = 0; q < kVec; ++q) v[q] = tail<KIND>(c[q], tp);
            store4<OutT>(out + (i0 + i) * ld_out + j_first, v, valid, false);
        }
    }
}

template <typename OutT, int KIND>
__global__ void sphere_gram_diag_kernel(const double* __restrict__ x1, const double* __restrict__ x2, int64_t n,
                                        int dim, TailParams tp, OutT* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c = 0.0;
    for (int k = 0; k < dim; ++k) c = fma(x1[i * dim + k], x2[i * dim + k], c);
    out[i] = static_cast<OutT>(tail<KIND>(c, tp));
}

TailParams make_tail(double param, int kind) {
    TailParams tp;
    const double k = (kind == GABO_KIND_DIST) ? 0.0 : -param * 1.4426950408889634074;
    tp.k_hi = static_cast<float>(k);
    tp.k_lo = static_cast<float>(k - static_cast<double>(tp.k_hi));
    // coefficients of asin2_sqrt (common.cuh), highest power first; a_j is the coefficient of w^j in P(w)
    static const double a_hi_first[9] = {0.3292977809906006, -0.3745849132537842, 0.28826069831848145,
                                         -0.03355207294225693, 0.07700732350349426, 0.07967597246170044,
                                         0.1143670305609703, 0.17777620255947113, 0.3333333432674408};
    for (int i = 0; i < 9; ++i) tp.pc[i] = static_cast<float>(k * a_hi_first[i] / static_cast<double>(1 << (8 - i)));
    tp.pc[9] = static_cast<float>(2.0 * k);
    const double skpi = sqrt(fabs(k)) * 3.14159265358979323846;
    tp.skpi_hi = static_cast<float>(skpi);
    tp.skpi_lo = static_cast<float>(skpi - static_cast<double>(tp.skpi_hi));
    return tp;
}

template <int D, typename OutT, int KIND>
int launch_fixed(const double* x1, int64_t n1, const double* x2, int64_t n2, TailParams tp, void* out, int64_t ld_out,
                 cudaStream_t stream) {
    const int64_t tiles_i = (n1 + kTileM - 1) / kTileM;
    const int64_t tiles_j = (n2 + kTileN - 1) / kTileN;
    const int64_t tiles = tiles_i * tiles_j;
    int occ = 8;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sphere_gram_kernel<D, OutT, KIND>, kThreads, 0);
    if (occ < 1) occ = 1;
    const int64_t grid = imin(tiles, static_cast<int64_t>(sm_count()) * occ);
    const bool vec_ok = (ld_out % kVec == 0) && aligned16(out);
    sphere_gram_kernel<D, OutT, KIND><<<static_cast<unsigned>(grid), kThreads, 0, stream>>>(
        x1, n1, x2, n2, tp, static_cast<OutT*>(out), ld_out, tiles_i, tiles, vec_ok);
    return check_launch("sphere_gram_kernel");
}

template <typename OutT, int KIND>
int launch_dim(const double* x1, int64_t n1, const double* x2, int64_t n2, int dim, TailParams tp, void* out,
               int64_t ld_out, cudaStream_t stream) {
    switch (dim) {
#define GABO_CASE(DD) \
    case DD:          \
        return launch_fixed<DD, OutT, KIND>(x1, n1, x2, n2, tp, out, ld_out, stream);
        GABO_CASE(2)
        GABO_CASE(3)
        GABO_CASE(4)
        GABO_CASE(5)
        GABO_CASE(6)
        GABO_CASE(7)
        GABO_CASE(8)
        GABO_CASE(9)
        GABO_CASE(10)
        GABO_CASE(11)
        GABO_CASE(12)
#undef GABO_CASE
        default: {
            const int64_t tiles_i = (n1 + kTileM - 1) / kTileM;
            const int64_t tiles_j = (n2 + kTileN - 1) / kTileN;
            const int64_t tiles = tiles_i * tiles_j;
            const int64_t grid = imin(tiles, static_cast<int64_t>(sm_count()) * 8);
            const size_t smem = sizeof(double) * kTileM * dim;
            sphere_gram_generic_kernel<OutT, KIND><<<static_cast<unsigned>(grid), kThreads, smem, stream>>>(
                x1, n1, x2, n2, dim, tp, static_cast<OutT*>(out), ld_out, tiles_i, tiles);
            return check_launch("sphere_gram_generic_kernel");
        }
    }
}

template <typename OutT>
int launch_kind(const double* x1, int64_t n1, const double* x2, int64_t n2, int dim, double param, int kind, void* out,
                int64_t ld_out, cudaStream_t stream) {
    const TailParams tp = make_tail(param, kind);
    switch (kind) {
        case GABO_KIND_GAUSS:
            return launch_dim<OutT, GABO_KIND_GAUSS>(x1, n1, x2, n2, dim, tp, out, ld_out, stream);
        case GABO_KIND_LAPLACE:
            return launch_dim<OutT, GABO_KIND_LAPLACE>(x1, n1, x2, n2, dim, tp, out, ld_out, stream);
        default:
            return launch_dim<OutT, GABO_KIND_DIST>(x1, n1, x2, n2, dim, tp, out, ld_out, stream);
    }
}

template <typename OutT>
int launch_diag(const double* x1, const double* x2, int64_t n, int dim, double param, int kind, void* out,
                cudaStream_t stream) {
    const TailParams tp = make_tail(param, kind);
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    OutT* o = static_cast<OutT*>(out);
    switch (kind) {
        case GABO_KIND_GAUSS:
            sphere_gram_diag_kernel<OutT, GABO_KIND_GAUSS><<<grid, 256, 0, stream>>>(x1, x2, n, dim, tp, o);
            break;
        case GABO_KIND_LAPLACE:
            sphere_gram_diag_kernel<OutT, GABO_KIND_LAPLACE><<<grid, 256, 0, stream>>>(x1, x2, n, dim, tp, o);
            break;
        default:
            sphere_gram_diag_kernel<OutT, GABO_KIND_DIST><<<grid, 256, 0, stream>>>(x1, x2, n, dim, tp, o);
            break;
    }
    return check_launch("sphere_gram_diag_kernel");
}

}  // namespace

}  // namespace gabo

extern "C" int gabo_sphere_gram(const double* x1, int64_t n1, const double* x2, int64_t n2, int dim, double param,
                                int kind, void* out, int out_dtype, int64_t ld_out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n1 >= 0 && n2 >= 0, GABO_E_ARG, "gabo_sphere_gram: negative size");
    if (n1 == 0 || n2 == 0) return GABO_OK;
    GABO_REQUIRE(x1 && x2 && out, GABO_E_ARG, "gabo_sphere_gram: null pointer");
    GABO_REQUIRE(dim >= 1 && dim <= GABO_MAX_SPHERE_DIM, GABO_E_ARG, "gabo_sphere_gram: dim %d outside [1, %d]", dim,
                 GABO_MAX_SPHERE_DIM);
    GABO_REQUIRE(kind >= GABO_KIND_GAUSS && kind <= GABO_KIND_DIST, GABO_E_ARG, "gabo_sphere_gram: bad kind %d", kind);
    GABO_REQUIRE(out_dtype == GABO_F32 || out_dtype == GABO_F64, GABO_E_ARG, "gabo_sphere_gram: bad out_dtype");
    GABO_REQUIRE(ld_out >= n2, GABO_E_ARG, "gabo_sphere_gram: ld_out < n2");
    // 8-byte (element) alignment is enough: the TMA bulk copy of an x1 tile is used only when the tile is 16-byte aligned
    // (cooperative loads otherwise) and x2 is read with scalar loads, so batch slices b*N*D with odd N*D are accepted
    GABO_REQUIRE(((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2)) & 7u) == 0, GABO_E_ALIGN,
                 "gabo_sphere_gram: inputs must be 8-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dim == 1) dim = 1;  // handled by the generic kernel
    if (out_dtype == GABO_F32) return launch_kind<float>(x1, n1, x2, n2, dim, param, kind, out, ld_out, s);
    return launch_kind<double>(x1, n1, x2, n2, dim, param, kind, out, ld_out, s);
}

extern "C" int gabo_sphere_gram_diag(const double* x1, const double* x2, int64_t n, int dim, double param, int kind,
                                     void* out, int out_dtype, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_sphere_gram_diag: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x1 && x2 && out, GABO_E_ARG, "gabo_sphere_gram_diag: null pointer");
    GABO_REQUIRE(dim >= 1 && dim <= GABO_MAX_SPHERE_DIM, GABO_E_ARG, "gabo_sphere_gram_diag: bad dim %d", dim);
    GABO_REQUIRE(kind >= GABO_KIND_GAUSS && kind <= GABO_KIND_DIST, GABO_E_ARG, "gabo_sphere_gram_diag: bad kind");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (out_dtype == GABO_F32) return launch_diag<float>(x1, x2, n, dim, param, kind, out, s);
    if (out_dtype == GABO_F64) return launch_diag<double>(x1, x2, n, dim, param, kind, out, s);
    set_error("gabo_sphere_gram_diag: bad out_dtype");
    return GABO_E_ARG;
}
