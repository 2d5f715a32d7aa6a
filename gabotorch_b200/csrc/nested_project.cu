// Nested SPD projection Y = W^T X W of HD-GaBO (P1 of SURVEY.md section 8) as ONE dense contraction in Mandel
// coordinates:  y[n x dvl] = x[n x dvh] * P^T  with  P[(a,b),(p,q)] = m_ab (W_pa W_qb + [p != q] W_qa W_pb) / m_pq.
//
// Replaces projection_from_spd_to_nested_spd (BoManifolds/nested_mappings/nested_spd_utils.py:13-48: W repeated n
// times, two bmm) together with the Mandel unpack / pack either side of it (kernels_nested_spd.py:122-127).
//
// Roofline: HBM read, 4 (dvh + dvl) bytes per matrix (900 B for SPD(20) -> SPD(5)) against 2 dvh dvl flop; the FP32
// pipe alone cannot keep up with HBM at that intensity, so the products run on the tensor cores as 3xTF32
// (a_hi b_hi + a_lo b_hi + a_hi b_lo, error ~ 2^-21) with fp32 accumulation -- the only GEMM-shaped op on the path.
// Layout: persistent CTAs (one per SM), 64-row tiles of x streamed through a shared-memory ring by the TMA bulk-copy
// engine (rows are contiguous, a tile is one 1-D copy), sixteen warps = 4 groups of 16 rows x 4 parts of k.
// Operand roles (round 2): the OPERATOR is the 16-row A fragment of mma.m16n8k8 (16 output Mandel entries per tile, pre-
// split into hi/lo and pre-arranged per lane, kept in registers when it fits) and the STREAMED rows are the 8-column B
// fragment, whose two values per lane are adjacent in memory: one 64-bit shared load and two LOP3 + two FADD (the
// 3xTF32 split) feed three MMAs with no register shuffling.  The round-1 form had the streamed rows as the A fragment:
// its four values per lane come from two rows, and interleaving them cost 62 IMAD.MOV per loop body -- 21 % of all
// instructions (ncu), which is what kept the kernel at 70 % of HBM (6.7 k warp-instructions per tile at 2 per clock).
#include <cstdlib>

#include "spd_common.cuh"
#include "nested_project_tc.cuh"

namespace gabo {
namespace {

#ifndef GABO_NP_ROWGROUPS
#define GABO_NP_ROWGROUPS 2
#endif
constexpr int kRowGroups = GABO_NP_ROWGROUPS;       // 16-row groups per tile
constexpr int kTileRows = 16 * kRowGroups;          // 32-row tiles, TWO CTAs per SM: while one CTA reduces and stores its
#ifndef GABO_NP_CTAS
#define GABO_NP_CTAS ((GABO_NP_ROWGROUPS <= 2) ? 2 : 1)
#endif
constexpr int kCtasPerSm = GABO_NP_CTAS;   // tile the other one feeds the tensor pipe (one 64-row CTA per SM
constexpr int kStages = 3;                          // left the pipe idle during every epilogue: 0.73 -> see profiles/)
constexpr int kMaxStages = 6;                 // ring depth limit (shared memory decides: 4 stages for SPD(20))
constexpr int kSplitK = 4;                    // warps sharing a row group, each with a quarter of the k range
constexpr int kThreadsP = kRowGroups * kSplitK * 32;   // 8 warps per CTA

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4& a, float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
          "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// m16n8k16 bf16 product (same issue cost as the TF32 k8 product on the legacy pipe, scripts/micro/mma_rate.cu)
__device__ __forceinline__ void mma_bf16(float (&c)[4], const float4& a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
          "r"(__float_as_uint(a.w)), "r"(b0), "r"(b1));
}
// {lower half, upper half} = bf16(lo), bf16(hi), round to nearest
__device__ __forceinline__ uint32_t pack_bf16(float lower, float upper) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(upper), "f"(lower));
    return r;
}

// GABO_PROJECT_KERNEL: unset / "3xtf32" = three TF32 products per k-step; "bf16" = one TF32 product + ONE bf16 k16 product
// for both correction terms; "tc" = the tcgen05 kernel.  Read once; the operator pack depends on it.
int project_mode() {
    static const int mode = [] {
        const char* e = std::getenv("GABO_PROJECT_KERNEL");
        if (e == nullptr) return 0;
        if (e[0] == 't') return 2;
        if (e[0] == 'b') return 1;
        return 0;
    }();
    return mode;
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

// pack[((s * 32 + lane) * MT + j) * 8 + {0..3, 4..7}] = {hi, lo} of the A fragment (a0, a1, a2, a3) of m-tile j, k-step s:
//   a0 = P[16j + g][8s + 2t], a1 = P[16j + g + 8][8s + 2t], a2 = P[16j + g][8s + 2t + 1], a3 = P[16j + g + 8][8s + 2t + 1]
// (fragment k index t <-> memory column 8s + 2t and t + 4 <-> 8s + 2t + 1: a permutation of k shared with the B loads).
__global__ void projection_pack_kernel(const double* __restrict__ w, int D, int d, int ksteps, int mt, int c16,
                                       float* __restrict__ pack) {
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    const int total = ksteps * 32 * mt;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int j = e % mt;
        const int lane = (e / mt) & 31;
        const int s = (e / mt) >> 5;
        const int g = lane >> 2, t = lane & 3;
        float hi[4], lo[4];
        double full[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                  // a0 .. a3
            const int o = 16 * j + g + ((q & 1) ? 8 : 0);   // low Mandel index (output column)
            const int i = 8 * s + 2 * t + (q >> 1);         // high Mandel index (k)
            double v = 0.0;
            if (o < dvl && i < dvh) {
                int a, b, p, qq;
                mandel_rc_dev(d, o, a, b);
                mandel_rc_dev(D, i, p, qq);
                const double mab = (a == b) ? 1.0 : 1.4142135623730951;
                if (p == qq) v = mab * w[p * d + a] * w[p * d + b];
                else v = mab * (w[p * d + a] * w[qq * d + b] + w[qq * d + a] * w[p * d + b]) / 1.4142135623730951;
            }
            full[q] = v;
            hi[q] = to_tf32(static_cast<float>(v));
            lo[q] = static_cast<float>(v - static_cast<double>(hi[q]));
        }
        float* dst = pack + static_cast<int64_t>(e) * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = hi[q];
        if (c16) {
            // A fragment of the m16n8k16 correction product: k slots (2t, 2t + 1) pair with lo(x) of the lane's two columns
            // (operator entry in bf16), slots (2t + 8, 2t + 9) with hi(x) (operator residual lo = P - tf32(P) in bf16)
            uint32_t* c = reinterpret_cast<uint32_t*>(dst + 4);
            c[0] = pack_bf16(static_cast<float>(full[0]), static_cast<float>(full[2]));   // row g
            c[1] = pack_bf16(static_cast<float>(full[1]), static_cast<float>(full[3]));   // row g + 8
            c[2] = pack_bf16(lo[0], lo[2]);
            c[3] = pack_bf16(lo[1], lo[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[4 + q] = to_tf32(lo[q]);
        }
    }
}

// Work split inside a CTA (one CTA per SM, 16 warps): warp w handles 16 rows (two 8-row B groups, rows 16 (w & 3) ..)
// of the 64-row tile and a QUARTER of the k range (w >> 2); the partial accumulators are summed through shared memory.
// 3xTF32 split of the streamed operand: b_hi = b with the low 13 mantissa bits cleared (exactly a tf32 number, one
// LOP3), b_lo = b - b_hi (exact, one FADD; its own low bits are dropped by the tensor core: error 2^-21 |b|).
// The operator was split (round-to-nearest) when it was packed; when a warp's part of the k range is at most KH steps
// (and there is one m-tile) its fragments live in REGISTERS for the whole kernel (KH > 0), otherwise they are read from
// shared memory every tile (KH = 0).
template <int MT, bool EVEN, int KH, bool C16>
__global__ void __launch_bounds__(kThreadsP, kCtasPerSm)
    nested_project_kernel(const float* __restrict__ x, int64_t n, int dvh, int dvl, int ksteps, int nstages,
                          const float* __restrict__ pack, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_floats = kTileRows * dvh;  // 32 * dvh * 4 bytes: a multiple of 128
    const int os_floats = (kTileRows * dvl + 3) & ~3;
    float* As = reinterpret_cast<float*>(smem_raw);
    // KH > 0: the operator fragments live in registers, so their shared-memory image is only needed while they are
    // loaded -- it borrows the LAST ring stage, whose first TMA copy is issued after that; the space buys a deeper ring.
    const int borrow = (KH > 0) ? (ksteps * 32 * MT * 8 + stage_floats - 1) / stage_floats : 0;   // stages lent to the image
    float* Bp = As + (nstages - borrow) * stage_floats;
    float* Os = (KH > 0) ? As + nstages * stage_floats : Bp + ksteps * 32 * MT * 8;   // kSplitK partial output tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(Os + kSplitK * os_floats);

    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    for (int e = threadIdx.x; e < ksteps * 32 * MT * 2; e += kThreadsP)
        reinterpret_cast<float4*>(Bp)[e] = __ldg(reinterpret_cast<const float4*>(pack) + e);
    __syncthreads();

    const uint32_t full_bytes = static_cast<uint32_t>(stage_floats) * 4u;
    auto issue = [&](int64_t tile, int stage) {  // thread 0 only; full tiles only
        if (tile < tiles && (tile + 1) * kTileRows <= n) {
            mbar_expect_tx(&bars[stage], full_bytes);
            tma_load_1d(As + stage * stage_floats, x + tile * kTileRows * dvh, full_bytes, &bars[stage]);
        }
    };
    const int early = nstages - borrow;                   // stages whose first copy can start right away
    if (threadIdx.x == 0) {
        for (int s = 0; s < early; ++s) issue(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(s) * gridDim.x, s);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowgrp = warp % kRowGroups, khalf = warp / kRowGroups;   // khalf: which part of the k range (0 .. kSplitK-1)
    const int g = lane >> 2, t = lane & 3;
    const int ksplit = (ksteps + kSplitK - 1) / kSplitK;
    const int s_begin = static_cast<int>(imin(khalf * ksplit, ksteps));
    const int s_end = static_cast<int>(imin(s_begin + ksplit, ksteps));
    // the last k-step reaches past the end of a row when dvh is not a multiple of 8: the part that owns it masks it
    const bool has_ragged = (8 * ksteps != dvh) && s_begin < s_end && s_end == ksteps;
    float4 breg[KH > 0 ? KH : 1][MT][2];      // [k-step][m-tile][hi, lo]
    if (KH > 0) {
#pragma unroll
        for (int i = 0; i < (KH > 0 ? KH : 1); ++i)
#pragma unroll
            for (int j = 0; j < MT; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    breg[i][j][h] = (s_begin + i < s_end)
                                        ? reinterpret_cast<const float4*>(Bp)[((s_begin + i) * 32 + lane) * MT * 2 + j * 2 + h]
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();                                   // everyone has its fragments: the borrowed stage is free
        if (threadIdx.x == 0) {
            fence_proxy_async();                           // generic-proxy reads of Bp before the async-proxy write
            for (int s = early; s < nstages; ++s)
                issue(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(s) * gridDim.x, s);
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

        // Per 8-row group q and m-tile j three accumulators (hi*hi, hi*lo, lo*hi): independent mma.sync chains; the small
        // terms are summed first.
        float acc[2][MT][4], acc_a[2][MT][4], acc_b[2][MT][4];
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < MT; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[q][j][c] = 0.0f;
                    acc_a[q][j][c] = 0.0f;
                    acc_b[q][j][c] = 0.0f;
                }
        const float* r0 = A + (rowgrp * 16 + g) * dvh;     // B-fragment column n = g  <->  tile row 16 rowgrp + 8 q + g
        const float* r1 = r0 + 8 * dvh;
        const float4* bp = reinterpret_cast<const float4*>(Bp) + lane * MT * 2;
        auto kstep = [&](const float (&b)[2][2], const float4 (&p)[MT][2]) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float bh0 = __uint_as_float(__float_as_uint(b[q][0]) & 0xffffe000u);
                const float bh1 = __uint_as_float(__float_as_uint(b[q][1]) & 0xffffe000u);
                const float bl0 = b[q][0] - bh0, bl1 = b[q][1] - bh1;
#pragma unroll
                for (int j = 0; j < MT; ++j) {
#if !defined(GABO_NP_ABLATE)
                    if (C16) {   // both correction terms in ONE bf16 k16 product: slots 0-7 = P lo(x), slots 8-15 = lo(P) hi(x)
                        mma_bf16(acc_a[q][j], p[j][1], pack_bf16(bl0, bl1), pack_bf16(bh0, bh1));
                    } else {
                        mma_tf32(acc_a[q][j], p[j][0], bl0, bl1);     // hi(P) lo(x)
                        mma_tf32(acc_b[q][j], p[j][1], bh0, bh1);     // lo(P) hi(x)
                    }
                    mma_tf32(acc[q][j], p[j][0], bh0, bh1);       // hi(P) hi(x)
#elif GABO_NP_ABLATE == 1   // ablation (scripts/micro/project_variants.cu): everything but the MMAs
                    acc_a[q][j][0] += p[j][0].x * bl0; acc_b[q][j][1] += p[j][1].y * bh1; acc[q][j][2] += bl1 + bh0;
#else                       // ablation: plain TF32, no split
                    mma_tf32(acc[q][j], p[j][0], b[q][0], b[q][1]);
#endif
                }
            }
        };
        auto load_b = [&](int s, float (&b)[2][2], bool ragged) {
            const int col = 8 * s + 2 * t;
            if (ragged) {   // columns past the end of the row are zero, never read
                b[0][0] = (col < dvh) ? r0[col] : 0.0f;
                b[1][0] = (col < dvh) ? r1[col] : 0.0f;
                b[0][1] = (col + 1 < dvh) ? r0[col + 1] : 0.0f;
                b[1][1] = (col + 1 < dvh) ? r1[col + 1] : 0.0f;
            } else if (EVEN) {
                const float2 u = *reinterpret_cast<const float2*>(r0 + col);
                const float2 v = *reinterpret_cast<const float2*>(r1 + col);
                b[0][0] = u.x; b[0][1] = u.y; b[1][0] = v.x; b[1][1] = v.y;
            } else {
                b[0][0] = r0[col]; b[0][1] = r0[col + 1]; b[1][0] = r1[col]; b[1][1] = r1[col + 1];
            }
        };
        if (KH > 0) {   // operator fragments in registers: the k loop is fully unrolled
#pragma unroll
            for (int i = 0; i < (KH > 0 ? KH : 1); ++i) {
                const int s = s_begin + i;
                if (s < s_end) {
                    float b[2][2];
                    load_b(s, b, has_ragged && s == s_end - 1);
                    kstep(b, breg[KH > 0 ? i : 0]);
                }
            }
        } else {
#pragma unroll 2
            for (int s = s_begin; s < s_end; ++s) {
                float b[2][2];
                load_b(s, b, has_ragged && s == s_end - 1);
                float4 p[MT][2];
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                    p[j][0] = bp[s * 32 * MT * 2 + j * 2];
                    p[j][1] = bp[s * 32 * MT * 2 + j * 2 + 1];
                }
                kstep(b, p);
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int j = 0; j < MT; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[q][j][c] += acc_a[q][j][c] + acc_b[q][j][c];

        __syncthreads();  // every warp is done with this stage (and with the previous tile's output staging)
        if (threadIdx.x == 0) {
            fence_proxy_async();
            issue(tile + static_cast<int64_t>(nstages) * gridDim.x, stage);
        }
        // C fragment: c0 = (m = g, n = 2t), c1 = (m = g, n = 2t + 1), c2 = (m = g + 8, n = 2t), c3 = (m = g + 8, n = 2t + 1);
        // m = output Mandel entry inside m-tile j, n = row inside the 8-row group q
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float* o0 = Os + khalf * os_floats + (rowgrp * 16 + q * 8 + 2 * t) * dvl;
            float* o1 = o0 + dvl;
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const int m0 = 16 * j + g;
                if (m0 < dvl) { o0[m0] = acc[q][j][0]; o1[m0] = acc[q][j][1]; }
                if (m0 + 8 < dvl) { o0[m0 + 8] = acc[q][j][2]; o1[m0 + 8] = acc[q][j][3]; }
            }
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

// Canonical (K-major, no-swizzle UMMA core-matrix) image of the operator for the tcgen05 kernel: [hi | lo], 16 x kp floats each,
// float index of (n, k) = ((k / 4) * 2 + n / 8) * 32 + (n % 8) * 4 + k % 4; rows n >= dvl and columns k >= dvh are zero.
__global__ void projection_pack_canonical_kernel(const double* __restrict__ w, int D, int d, int kp, float* __restrict__ out) {
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    const int total = tc::kN * kp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int o = e / kp, i = e % kp;
        double v = 0.0;
        if (o < dvl && i < dvh) {
            int a, b, p, q;
            mandel_rc_dev(d, o, a, b);
            mandel_rc_dev(D, i, p, q);
            const double mab = (a == b) ? 1.0 : 1.4142135623730951;
            if (p == q) v = mab * w[p * d + a] * w[p * d + b];
            else v = mab * (w[p * d + a] * w[q * d + b] + w[q * d + a] * w[p * d + b]) / 1.4142135623730951;
        }
        const float hi = to_tf32(static_cast<float>(v));
        const float lo = to_tf32(static_cast<float>(v - static_cast<double>(hi)));
        const int idx = ((i / 4) * 2 + o / 8) * 32 + (o % 8) * 4 + i % 4;
        out[idx] = hi;
        out[total + idx] = lo;
    }
}

int ksteps_for(int dvh) { return (dvh + 7) / 8; }
// tcgen05 path: at most 16 output Mandel entries (d <= 5) and a tile ring + chunk buffers + operator inside 227 KB
bool tc_eligible(int dvh, int dvl) { return dvl <= tc::kN && tc::Smem(dvh).total <= 227u * 1024u; }
int64_t fragment_pack_floats(int dvh, int dvl);
int ntiles_for(int dvl) { return (dvl + 15) / 16; }   // m16 tiles of output Mandel entries
int64_t fragment_pack_floats(int dvh, int dvl) { return static_cast<int64_t>(ksteps_for(dvh)) * 32 * ntiles_for(dvl) * 8; }

template <int MT>
int launch_nt(const float* x, int64_t n, int dvh, int dvl, const float* pack, float* y, cudaStream_t s) {
    const int ksteps = ksteps_for(dvh);
    // ring depth: kStages when it fits the 227 KB of shared memory, otherwise 2 (long Mandel vectors)
    constexpr int kRegSteps = (MT == 1) ? 7 : 0;
    const bool in_regs = kRegSteps > 0 && (ksteps + kSplitK - 1) / kSplitK <= kRegSteps;
    // operator in registers: no resident shared-memory image (it borrows the last stage at start-up), ring as deep as fits
    // up to kMaxStages; otherwise the image stays and the ring is kStages deep when that fits, 2 for long Mandel vectors
    auto smem_for = [&](int stages) {
        const size_t op = in_regs ? 0 : static_cast<size_t>(ksteps) * 32 * MT * 8;
        return sizeof(float) * (static_cast<size_t>(stages) * kTileRows * dvh + op + kSplitK * ((kTileRows * dvl + 3) & ~3)) +
               8 * kMaxStages + 16;
    };
    int nstages = in_regs ? kMaxStages : kStages;
    // per-CTA budget that still lets kCtasPerSm CTAs share the 228 KB of an SM (1 KB per CTA is reserved by the system)
    const size_t smem_cap = (kCtasPerSm == 2) ? 113 * 1024 + 512 : (kCtasPerSm > 2 ? (228 * 1024) / kCtasPerSm - 1024 : 227 * 1024);
    while (nstages > 2 && smem_for(nstages) > smem_cap) --nstages;
    const size_t smem = smem_for(nstages);
    GABO_REQUIRE(smem <= 227 * 1024, GABO_E_UNSUPPORTED,
                 "gabo_nested_spd_project: Mandel length %d needs %zu bytes of shared memory (> 227 KB)", dvh, smem);
    GABO_REQUIRE(!in_regs || static_cast<size_t>(ksteps) * 32 * MT * 8 <= static_cast<size_t>(nstages - 1) * kTileRows * dvh,
                 GABO_E_UNSUPPORTED, "gabo_nested_spd_project: operator image does not fit the borrowed ring stages");
    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    const unsigned grid = static_cast<unsigned>(imin(tiles, static_cast<int64_t>(sm_count()) * kCtasPerSm));
    const bool even = (dvh % 2) == 0;
    // operator fragments in registers when a warp's part of the k range is at most 7 steps and there is one m-tile
    // (SPD(20) -> SPD(5): 27 steps / 4 parts = 7, 8 floats each -> 56 registers); shared memory otherwise
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kThreadsP, smem, s>>>(x, n, dvh, dvl, ksteps, nstages, pack, y);
    };
    const bool c16 = project_mode() == 1;
    if (in_regs) {
        if (even) { if (c16) go(nested_project_kernel<MT, true, kRegSteps, true>); else go(nested_project_kernel<MT, true, kRegSteps, false>); }
        else { if (c16) go(nested_project_kernel<MT, false, kRegSteps, true>); else go(nested_project_kernel<MT, false, kRegSteps, false>); }
    } else {
        if (even) { if (c16) go(nested_project_kernel<MT, true, 0, true>); else go(nested_project_kernel<MT, true, 0, false>); }
        else { if (c16) go(nested_project_kernel<MT, false, 0, true>); else go(nested_project_kernel<MT, false, 0, false>); }
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
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    // [mma.sync fragment image | canonical UMMA image (hi, lo) when the tcgen05 kernel can take the shape]
    return gabo::fragment_pack_floats(dvh, dvl) + (gabo::tc_eligible(dvh, dvl) ? 2 * gabo::tc::kN * gabo::tc::Smem(dvh).kp : 0);
}

extern "C" int gabo_nested_projection_matrix(const double* w, int D, int d, float* p_pack, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(w && p_pack, GABO_E_ARG, "gabo_nested_projection_matrix: null pointer");
    GABO_REQUIRE(D >= 1 && d >= 1 && d <= D && d <= GABO_MAX_SPD_DIM, GABO_E_ARG,
                 "gabo_nested_projection_matrix: need 1 <= d <= min(D, %d), got D=%d d=%d", GABO_MAX_SPD_DIM, D, d);
    GABO_REQUIRE(aligned16(p_pack), GABO_E_ALIGN, "gabo_nested_projection_matrix: pack must be 16-byte aligned");
    const int ksteps = ksteps_for(D * (D + 1) / 2), nt = ntiles_for(d * (d + 1) / 2);
    const int total = ksteps * 32 * nt * 4;
    projection_pack_kernel<<<(total / 4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, D, d, ksteps, nt, project_mode() == 1 ? 1 : 0, p_pack);
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    if (tc_eligible(dvh, dvl)) {
        const int kp = tc::Smem(dvh).kp;
        projection_pack_canonical_kernel<<<(tc::kN * kp + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            w, D, d, kp, p_pack + fragment_pack_floats(dvh, dvl));
    }
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
    // The tcgen05 / TMEM kernel (nested_project_tc.cuh) is parity-green but slower than the mma.sync kernel on this skinny
    // shape (N = 2^20: 0.64 ms against 0.196 ms; ~110 cycles per 64 x 16 x 8 tcgen05.mma, see its header): opt-in with
    // GABO_PROJECT_KERNEL=tc, kept as the measured answer to "would tcgen05 help this GEMM as it stands?".
    const bool want_tc = project_mode() == 2;
    if (tc_eligible(dvh, dvl) && want_tc && n >= tc::kRows) {
        const tc::Smem L(dvh);
        const cudaError_t e = cudaFuncSetAttribute(tc::nested_project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(L.total));
        GABO_REQUIRE(e == cudaSuccess, GABO_E_CUDA, "nested_project_tc_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        const int64_t tiles = (n + tc::kRows - 1) / tc::kRows;
        const unsigned grid = static_cast<unsigned>(imin(tiles, sm_count()));
        tc::nested_project_tc_kernel<<<grid, tc::kThreads, L.total, s>>>(x_mandel, n, dvh, dvl,
                                                                         p_pack + fragment_pack_floats(dvh, dvl), y_mandel);
        return check_launch("nested_project_tc_kernel");
    }
    switch (ntiles_for(dvl)) {
        case 1: return launch_nt<1>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 2: return launch_nt<2>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 3: return launch_nt<3>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
    }
    return GABO_E_ARG;
}
