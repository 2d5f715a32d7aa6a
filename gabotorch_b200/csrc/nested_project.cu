// Nested SPD projection Y = W^T X W of HD-GaBO (P1 of SURVEY.md section 8) as ONE dense contraction in Mandel
// coordinates:  y[n x dvl] = x[n x dvh] * P^T  with  P[(a,b),(p,q)] = m_ab (W_pa W_qb + [p != q] W_qa W_pb) / m_pq.
//
// Replaces projection_from_spd_to_nested_spd (BoManifolds/nested_mappings/nested_spd_utils.py:13-48: W repeated n
// times, two bmm) together with the Mandel unpack / pack either side of it (kernels_nested_spd.py:122-127).
//
// Roofline: HBM read, 4 (dvh + dvl) bytes per matrix (900 B for SPD(20) -> SPD(5)) against 2 dvh dvl flop; the FP32
// pipe alone cannot keep up with HBM at that intensity, so the products run on the tensor cores as 3xTF32
// (a_hi b_hi + a_lo b_hi + a_hi b_lo, error ~ 2^-21) with fp32 accumulation -- the only GEMM-shaped op on the path.
// Layout: persistent CTAs (one per SM), 64-row tiles of x streamed through a 3-stage shared-memory ring by the TMA
// bulk-copy engine (rows are contiguous, a tile is one 1-D copy), sixteen warps = 4 row groups (m16) x 4 parts of k,
// P pre-split into hi/lo and pre-arranged per lane (one 16-byte value per 8 output columns and k-step), kept in
// registers when it fits.
#include "spd_common.cuh"

namespace gabo {
namespace {

constexpr int kTileRows = 64;
constexpr int kStages = 3;
constexpr int kMaxStages = 6;                 // ring depth limit (shared memory decides: 4 stages for SPD(20))
constexpr int kSplitK = 4;                    // warps sharing a row group, each with a quarter of the k range
constexpr int kThreadsP = 4 * kSplitK * 32;   // 16 warps

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
          "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

__device__ __forceinline__ void mandel_rc_dev(int d, int pos, int& r, int& c) {
    int k = 0, len = d;
    while (pos >= len) {
        pos -= len;
        --len;
        ++k;
    }
    r = pos;
    c = pos + k;
}

// pack[(s * 32 + lane) * NT + j][4] = { hi(b0), hi(b1), lo(b0), lo(b1) },  b0 = P[8j + g][8s + 2t], b1 = P[8j + g][8s + 2t + 1]
__global__ void projection_pack_kernel(const double* __restrict__ w, int D, int d, int ksteps, int nt,
                                       float* __restrict__ pack) {
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    const int total = ksteps * 32 * nt * 2;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int which = e & 1;
        const int j = (e >> 1) % nt;
        const int lane = ((e >> 1) / nt) & 31;
        const int s = ((e >> 1) / nt) >> 5;
        const int g = lane >> 2, t = lane & 3;
        const int o = 8 * j + g;           // low Mandel index (output column)
        const int i = 8 * s + 2 * t + which;  // high Mandel index (k)
        double v = 0.0;
        if (o < dvl && i < dvh) {
            int a, b, p, q;
            mandel_rc_dev(d, o, a, b);
            mandel_rc_dev(D, i, p, q);
            const double mab = (a == b) ? 1.0 : 1.4142135623730951;
            if (p == q) v = mab * w[p * d + a] * w[p * d + b];
            else v = mab * (w[p * d + a] * w[q * d + b] + w[q * d + a] * w[p * d + b]) / 1.4142135623730951;
        }
        const float vf = static_cast<float>(v);
        const float hi = to_tf32(vf);
        const float lo = to_tf32(static_cast<float>(v - static_cast<double>(hi)));
        float* dst = pack + (static_cast<int64_t>((s * 32 + lane) * nt + j)) * 4;
        dst[which] = hi;
        dst[2 + which] = lo;
    }
}

// Work split inside a CTA (one CTA per SM, 16 warps): warp w handles 16 rows (group w & 3) of the 64-row tile and a
// QUARTER of the k range (w >> 2); the partial accumulators of a row group are summed through shared memory.  ncu on
// the 4- and 8-warp versions: no pipe above 35 %, every warp waiting on its own fixed-latency dependencies (~650
// instructions per tile at ~6 cycles each) -- the cure is more warps per scheduler, each with a shorter stream.
// 3xTF32 split of the streamed operand: a_hi = a with the low 13 mantissa bits cleared (exactly a tf32 number, one
// LOP3), a_lo = a - a_hi (exact, one FADD; its own low bits are dropped by the tensor core: error 2^-21 |a|).
// The projection operator P was split (round-to-nearest) when it was packed; when a warp's half of the k range is at
// most KH steps its fragments of P live in REGISTERS for the whole kernel (KH > 0), otherwise they are read from
// shared memory every tile (KH = 0).
template <int NT, bool EVEN, int KH>
__global__ void __launch_bounds__(kThreadsP, 1)
    nested_project_kernel(const float* __restrict__ x, int64_t n, int dvh, int dvl, int ksteps, int nstages,
                          const float* __restrict__ pack, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_floats = kTileRows * dvh;  // 64 * dvh * 4 bytes: a multiple of 256
    const int os_floats = (kTileRows * dvl + 3) & ~3;
    float* As = reinterpret_cast<float*>(smem_raw);
    // KH > 0: the operator fragments live in registers, so their shared-memory image is only needed while they are
    // loaded -- it borrows the LAST ring stage, whose first TMA copy is issued after that.  The space this frees buys a
    // fourth ring stage (SPD(20): 4 x 53.8 KB): with three stages only ~107 KB per SM were in flight while a tile was
    // being processed, short of the ~90-130 KB Little's law asks for at 44 GB/s per SM and 2-3 us loaded HBM latency.
    float* Bp = (KH > 0) ? As + (nstages - 1) * stage_floats : As + nstages * stage_floats;
    float* Os = (KH > 0) ? As + nstages * stage_floats : Bp + ksteps * 32 * NT * 4;   // kSplitK partial output tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(Os + kSplitK * os_floats);

    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    for (int e = threadIdx.x; e < ksteps * 32 * NT; e += kThreadsP)
        reinterpret_cast<float4*>(Bp)[e] = __ldg(reinterpret_cast<const float4*>(pack) + e);
    __syncthreads();

    const uint32_t full_bytes = static_cast<uint32_t>(stage_floats) * 4u;
    auto issue = [&](int64_t tile, int stage) {  // thread 0 only; full tiles only
        if (tile < tiles && (tile + 1) * kTileRows <= n) {
            mbar_expect_tx(&bars[stage], full_bytes);
            tma_load_1d(As + stage * stage_floats, x + tile * kTileRows * dvh, full_bytes, &bars[stage]);
        }
    };
    const int early = (KH > 0) ? nstages - 1 : nstages;   // stages whose first copy can start right away
    if (threadIdx.x == 0) {
        for (int s = 0; s < early; ++s) issue(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(s) * gridDim.x, s);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowgrp = warp & 3, khalf = warp >> 2;   // khalf: which part of the k range (0 .. kSplitK-1)
    const int g = lane >> 2, t = lane & 3;
    // Fragment row -> tile row.  The rows are 4 dvh bytes apart (a whole tile is ONE bulk copy, so they cannot be
    // padded) and with the natural mapping the 64-bit fragment loads of a half-warp (4 rows x 4 column pairs) collide
    // two ways for dvh = 210 (ncu: 52 % of the shared-memory wavefronts were conflicts and the LSU pipe, at 70 %, was
    // the limiter).  Rows are independent outputs, so fragment rows g / g+8 are mapped to tile rows
    // 4 (g & 3) + (g >> 2) and that + 2: a half-warp then touches rows 4 apart, whose bank offsets are 8 words apart.
    const int row_lo = 4 * (g & 3) + (g >> 2);
    const int ksplit = (ksteps + kSplitK - 1) / kSplitK;
    const int s_begin = static_cast<int>(imin(khalf * ksplit, ksteps));
    const int s_end = static_cast<int>(imin(s_begin + ksplit, ksteps));
    // the last k-step reaches past the end of a row when dvh is not a multiple of 8: the half that owns it masks it
    const bool has_ragged = (8 * ksteps != dvh) && s_begin < s_end && s_end == ksteps;
    float4 breg[KH > 0 ? KH : 1][NT];
    if (KH > 0) {
#pragma unroll
        for (int i = 0; i < (KH > 0 ? KH : 1); ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j)
                breg[i][j] = (s_begin + i < s_end)
                                 ? reinterpret_cast<const float4*>(Bp)[(s_begin + i) * 32 * NT + lane * NT + j]
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();                                   // everyone has its fragments: the borrowed stage is free
        if (threadIdx.x == 0) {
            fence_proxy_async();                           // generic-proxy reads of Bp before the async-proxy write
            issue(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(nstages - 1) * gridDim.x, nstages - 1);
        }
    }
    uint32_t phase_bits = 0u;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int stage = it % nstages;
        const int rows = static_cast<int>(imin(kTileRows, n - tile * kTileRows));
        float* A = As + stage * stage_floats;
        if (rows == kTileRows) {
            mbar_wait(&bars[stage], (phase_bits >> stage) & 1u);
            phase_bits ^= (1u << stage);
        } else {  // ragged last tile: plain cooperative copy
            const float* src = x + tile * kTileRows * dvh;
            for (int e = threadIdx.x; e < rows * dvh; e += kThreadsP) A[e] = src[e];
            for (int e = rows * dvh + threadIdx.x; e < stage_floats; e += kThreadsP) A[e] = 0.0f;
            __syncthreads();
        }

        // Three accumulators per output tile (hi*hi, lo*hi, hi*lo): independent mma.sync chains; small terms summed first.
        float acc[NT][4], acc_lh[NT][4], acc_hl[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[j][q] = 0.0f;
                acc_lh[j][q] = 0.0f;
                acc_hl[j][q] = 0.0f;
            }
        const float* r0 = A + (rowgrp * 16 + row_lo) * dvh;
        const float* r1 = r0 + 2 * dvh;
        const float4* bp = reinterpret_cast<const float4*>(Bp) + lane * NT;
        auto kstep = [&](const float (&a)[4], const float4 (&b)[NT]) {  // a0:(lo row, col) a1:(hi row, col) a2/a3: col+1
            float ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ah[q] = __uint_as_float(__float_as_uint(a[q]) & 0xffffe000u);
                al[q] = a[q] - ah[q];
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {   // b = {hi b0, hi b1, lo b0, lo b1}
                mma_tf32(acc_lh[j], al, b[j].x, b[j].y);
                mma_tf32(acc_hl[j], ah, b[j].z, b[j].w);
                mma_tf32(acc[j], ah, b[j].x, b[j].y);
            }
        };
        auto load_a = [&](int s, float (&a)[4], bool ragged) {
            const int col = 8 * s + 2 * t;
            if (ragged) {   // columns past the end of the row are zero, never read
                a[0] = (col < dvh) ? r0[col] : 0.0f;
                a[1] = (col < dvh) ? r1[col] : 0.0f;
                a[2] = (col + 1 < dvh) ? r0[col + 1] : 0.0f;
                a[3] = (col + 1 < dvh) ? r1[col + 1] : 0.0f;
            } else if (EVEN) {
                const float2 u = *reinterpret_cast<const float2*>(r0 + col);
                const float2 v = *reinterpret_cast<const float2*>(r1 + col);
                a[0] = u.x; a[2] = u.y; a[1] = v.x; a[3] = v.y;
            } else {
                a[0] = r0[col]; a[2] = r0[col + 1]; a[1] = r1[col]; a[3] = r1[col + 1];
            }
        };
        if (KH > 0) {   // operator fragments in registers: the k loop is fully unrolled
#pragma unroll
            for (int i = 0; i < (KH > 0 ? KH : 1); ++i) {
                const int s = s_begin + i;
                if (s < s_end) {
                    float a[4];
                    load_a(s, a, has_ragged && s == s_end - 1);
                    kstep(a, breg[KH > 0 ? i : 0]);
                }
            }
        } else {
#pragma unroll 2
            for (int s = s_begin; s < s_end; ++s) {
                float a[4];
                load_a(s, a, has_ragged && s == s_end - 1);
                float4 b[NT];
#pragma unroll
                for (int j = 0; j < NT; ++j) b[j] = bp[s * 32 * NT + j];
                kstep(a, b);
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] += acc_lh[j][q] + acc_hl[j][q];

        __syncthreads();  // every warp is done with this stage (and with the previous tile's output staging)
        if (threadIdx.x == 0) {
            fence_proxy_async();
            issue(tile + static_cast<int64_t>(nstages) * gridDim.x, stage);
        }
        float* o0 = Os + khalf * os_floats + (rowgrp * 16 + row_lo) * dvl;
        float* o1 = o0 + 2 * dvl;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c0 = 8 * j + 2 * t;
            if (c0 < dvl) { o0[c0] = acc[j][0]; o1[c0] = acc[j][2]; }
            if (c0 + 1 < dvl) { o0[c0 + 1] = acc[j][1]; o1[c0 + 1] = acc[j][3]; }
        }
        __syncthreads();
        float* dst = y + tile * kTileRows * dvl;
        for (int e = threadIdx.x; e < rows * dvl; e += kThreadsP) {   // sum of the k-range partials, coalesced store
            float v = Os[e];
#pragma unroll
            for (int q = 1; q < kSplitK; ++q) v += Os[q * os_floats + e];
            __stcs(dst + e, v);
        }
    }
}

int ksteps_for(int dvh) { return (dvh + 7) / 8; }
int ntiles_for(int dvl) { return (dvl + 7) / 8; }

template <int NT>
int launch_nt(const float* x, int64_t n, int dvh, int dvl, const float* pack, float* y, cudaStream_t s) {
    const int ksteps = ksteps_for(dvh);
    // ring depth: kStages when it fits the 227 KB of shared memory, otherwise 2 (long Mandel vectors)
    constexpr int kRegSteps = (NT <= 2) ? 7 : 0;
    const bool in_regs = kRegSteps > 0 && (ksteps + kSplitK - 1) / kSplitK <= kRegSteps;
    // operator in registers: no resident shared-memory image (it borrows the last stage at start-up), ring as deep as fits
    // up to kMaxStages; otherwise the image stays and the ring is kStages deep when that fits, 2 for long Mandel vectors
    auto smem_for = [&](int stages) {
        const size_t op = in_regs ? 0 : static_cast<size_t>(ksteps) * 32 * NT * 4;
        return sizeof(float) * (static_cast<size_t>(stages) * kTileRows * dvh + op + kSplitK * ((kTileRows * dvl + 3) & ~3)) +
               8 * kMaxStages + 16;
    };
    int nstages = in_regs ? kMaxStages : kStages;
    while (nstages > 2 && smem_for(nstages) > 227 * 1024) --nstages;
    const size_t smem = smem_for(nstages);
    GABO_REQUIRE(smem <= 227 * 1024, GABO_E_UNSUPPORTED,
                 "gabo_nested_spd_project: Mandel length %d needs %zu bytes of shared memory (> 227 KB)", dvh, smem);
    GABO_REQUIRE(!in_regs || static_cast<size_t>(ksteps) * 32 * NT * 4 <= static_cast<size_t>(kTileRows) * dvh, GABO_E_UNSUPPORTED,
                 "gabo_nested_spd_project: operator image does not fit a ring stage");
    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    const unsigned grid = static_cast<unsigned>(imin(tiles, sm_count()));
    const bool even = (dvh % 2) == 0;
    // operator fragments in registers when a warp's part of the k range is at most 7 steps and NT <= 2
    // (SPD(20) -> SPD(5): 27 steps / 4 parts = 7, NT = 2 -> 56 registers); shared memory otherwise
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kThreadsP, smem, s>>>(x, n, dvh, dvl, ksteps, nstages, pack, y);
    };
    if (in_regs) {
        if (even) go(nested_project_kernel<NT, true, kRegSteps>);
        else go(nested_project_kernel<NT, false, kRegSteps>);
    } else {
        if (even) go(nested_project_kernel<NT, true, 0>);
        else go(nested_project_kernel<NT, false, 0>);
    }
    return check_launch("nested_project_kernel");
}

}  // namespace
}  // namespace gabo

namespace gabo {
namespace {
// Reference-precision path of the projection for the nested kernels (kernels_nested_spd.py:122-127): fp64 in, fp64 out,
// one thread per (matrix, output Mandel entry).  The tensor-core path above is fp32 (3xTF32, ~1e-6 relative), which is
// right for streaming millions of raw samples but not for the affine-invariant distance of ill-conditioned projected
// matrices (its error is amplified by their condition number); GP training sets are tens of points.
__global__ void nested_project_f64_kernel(const double* __restrict__ x, int64_t n, int D, int d,
                                          const double* __restrict__ w, double* __restrict__ y) {
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n * dvl) return;
    const int64_t i = e / dvl;
    const int o = static_cast<int>(e % dvl);
    int a, b;
    mandel_rc_dev(d, o, a, b);
    const double* xi = x + i * dvh;
    // y_ab = sum_pq W_pa X_pq W_qb, X_pq read from the Mandel vector (off-diagonal entries carry a factor sqrt 2)
    double acc = 0.0;
    for (int p = 0; p < D; ++p) {
        double t = 0.0;
        for (int q = 0; q < D; ++q) {
            const int lo = p < q ? p : q, hi = p < q ? q : p;
            const double xv = xi[mandel_pos(D, lo, hi)];
            t = fma((p == q) ? xv : xv * 0.70710678118654752440, w[q * d + b], t);
        }
        acc = fma(w[p * d + a], t, acc);
    }
    y[e] = (a == b) ? acc : acc * 1.41421356237309504880;
}
}  // namespace
}  // namespace gabo

extern "C" int gabo_nested_spd_project_f64(const double* x_mandel, int64_t n, int D, int d, const double* w,
                                           double* y_mandel, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_spd_project_f64: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x_mandel && w && y_mandel, GABO_E_ARG, "gabo_nested_spd_project_f64: null pointer");
    GABO_REQUIRE(D >= 1 && d >= 1 && d <= D, GABO_E_ARG, "gabo_nested_spd_project_f64: need 1 <= d <= D, got D=%d d=%d",
                 D, d);
    const int64_t total = n * (static_cast<int64_t>(d) * (d + 1) / 2);
    nested_project_f64_kernel<<<static_cast<unsigned>((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x_mandel, n, D, d, w, y_mandel);
    return check_launch("nested_project_f64_kernel");
}

extern "C" int64_t gabo_nested_projection_pack_size(int D, int d) {
    if (D < 1 || d < 1 || d > D || d > GABO_MAX_SPD_DIM) return -1;
    return static_cast<int64_t>(gabo::ksteps_for(D * (D + 1) / 2)) * 32 * gabo::ntiles_for(d * (d + 1) / 2) * 4;
}

extern "C" int gabo_nested_projection_matrix(const double* w, int D, int d, float* p_pack, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(w && p_pack, GABO_E_ARG, "gabo_nested_projection_matrix: null pointer");
    GABO_REQUIRE(D >= 1 && d >= 1 && d <= D && d <= GABO_MAX_SPD_DIM, GABO_E_ARG,
                 "gabo_nested_projection_matrix: need 1 <= d <= min(D, %d), got D=%d d=%d", GABO_MAX_SPD_DIM, D, d);
    GABO_REQUIRE(aligned16(p_pack), GABO_E_ALIGN, "gabo_nested_projection_matrix: pack must be 16-byte aligned");
    const int ksteps = ksteps_for(D * (D + 1) / 2), nt = ntiles_for(d * (d + 1) / 2);
    const int total = ksteps * 32 * nt * 2;
    projection_pack_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, D, d, ksteps, nt,
                                                                                              p_pack);
    return check_launch("projection_pack_kernel");
}

extern "C" int gabo_nested_spd_project(const float* x_mandel, int64_t n, int D, int d, const float* p_pack,
                                       float* y_mandel, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_spd_project: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x_mandel && p_pack && y_mandel, GABO_E_ARG, "gabo_nested_spd_project: null pointer");
    GABO_REQUIRE(D >= 1 && d >= 1 && d <= D && d <= GABO_MAX_SPD_DIM, GABO_E_ARG,
                 "gabo_nested_spd_project: need 1 <= d <= min(D, %d), got D=%d d=%d", GABO_MAX_SPD_DIM, D, d);
    GABO_REQUIRE(aligned16(x_mandel) && aligned16(p_pack), GABO_E_ALIGN,
                 "gabo_nested_spd_project: x and pack must be 16-byte aligned");
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (ntiles_for(dvl)) {
        case 1: return launch_nt<1>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 2: return launch_nt<2>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 3: return launch_nt<3>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 4: return launch_nt<4>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 5: return launch_nt<5>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
    }
    return GABO_E_ARG;
}
