// SPD affine-invariant Gram (G3 + G4 + G5 of SURVEY.md section 8).
//
// Replaces vector_to_symmetric_matrix_mandel_torch (Riemannian_utils/spd_utils_torch.py:159-194, a Python loop over
// matrices), affine_invariant_distance_torch (:53-120, one torch.symeig call per PAIR inside a Python loop, :109-110)
// and the d^2 / exp passes of SpdAffineInvariantGaussianKernel.forward (kernel_utils/kernels_spd.py:90-100).
//
//   gabo_spd_factor  : per point, fp64: Mandel unpack -> Cholesky L -> A = L^-1          (O(N d^3), negligible)
//   gabo_spd_ai_gram : per pair: G = A_i L_j (lower-triangular product, fp64 FMAs), one-sided Jacobi on G in
//                      registers (fp32 or fp64), lambda = |g_k|^2, then the reference's fp32 tail
//                      d = sqrt(sum log(lambda)^2 + 1e-15) and K = exp(-beta d^2).
// One THREAD per pair: N^2 pairs give far more parallelism than the chip has lanes, so a per-thread register-resident
// Jacobi needs no shuffles and no shared-memory traffic (the warp-cooperative form is used by the optimiser, where
// parallelism is scarce).  A thread owns one column j (L_j in registers for d <= 5) and walks down the rows of the
// tile, whose A_i are staged in shared memory by TMA bulk copies and read as broadcasts; a warp stores 128 contiguous
// bytes per row.  Roofline: FP32/FP64 pipe, not HBM (d=3: ~0.5 kFLOP per 4 output bytes).
#include <atomic>

#include "spd_common.cuh"

namespace gabo {

namespace {

constexpr int kThreads = 128;  // columns per tile
#ifndef GABO_TILE_OVERHEAD_ROWS
#define GABO_TILE_OVERHEAD_ROWS 4
#endif
constexpr int kTileOverheadRows = GABO_TILE_OVERHEAD_ROWS;  // tile-size model of launch_pair: per-tile overhead in rows

// ------------------------------------------------------------------------------------------------------------
// per-point factorisation
// ------------------------------------------------------------------------------------------------------------
// Two point sets per launch (x2 / fac2 may be null with n2 = 0): both operands of a Gram build are factorised by ONE
// launch -- at N = 2048 a factorisation launch is ~4 us of pure latency next to a ~105 us Gram kernel.
template <int d>
__global__ void spd_factor_kernel(const double* __restrict__ x1, int64_t n1, const double* __restrict__ x2, int64_t n2,
                                  int is_mandel, double* __restrict__ fac1, double* __restrict__ fac2,
                                  int32_t* __restrict__ flags) {
    constexpr int TRI = tri_size(d);
    constexpr int FS = factor_stride(d);
    // let a dependent launch (the pair kernel, programmatic stream serialisation) start its CTAs now; it still waits for
    // this grid to finish before it reads the records (griddepcontrol.wait)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n1 + n2) return;
    const double* x = x1;
    double* fac = fac1;
    if (i >= n1) {
        i -= n1;
        x = x2;
        fac = fac2;
    }
    double L[TRI], A[TRI];
    bool ok;
    if (is_mandel) {
        const double* v = x + i * TRI;
        // spd_utils_torch.py:181-187: diagonal first, then the k-th super-diagonal divided by sqrt(2)
        auto acc = [&](int r, int c) {  // r >= c
            const double e = v[mandel_pos(d, c, r)];
            return (r == c) ? e : e * 0.70710678118654752440;
        };
        ok = chol_inv<d>(acc, L, A);
    } else {
        const double* m = x + i * d * d;
        auto acc = [&](int r, int c) { return m[r * d + c]; };
        ok = chol_inv<d>(acc, L, A);
    }
    double* o = fac + i * FS;
#pragma unroll
    for (int e = 0; e < TRI; ++e) o[e] = L[e];
#pragma unroll
    for (int e = 0; e < TRI; ++e) o[TRI + e] = A[e];
    if (FS > 2 * TRI) o[2 * TRI] = 0.0;
    if (!ok && flags) atomicOr(flags, 1);
}

// ------------------------------------------------------------------------------------------------------------
// per-pair kernel
// ------------------------------------------------------------------------------------------------------------
// K = exp(-param d^2) / exp(-param d) / d.  `kp` = -param log2(e) split into two floats by the launcher.
// fp32 compute path: d^2 and the product with (k_hi + k_lo) in fp32 (relative error of the exponent 6e-8, i.e. of K at
// most 6e-8 * param d^2), 2^t on the MUFU.  fp64 ("reference-grade") path: the exponent is reduced in fp64.
struct ExpParams {
    float k_hi, k_lo;
    double param;
};

template <int KIND, typename T>
__device__ __forceinline__ float finish(float dist, const ExpParams& kp) {
    if (KIND == GABO_KIND_DIST) return dist;
    if (sizeof(T) == 4) {
        const float v = (KIND == GABO_KIND_GAUSS) ? dist * dist : dist;   // kernels_spd.py:96-98 / :185
        return ex2_approx(fmaf(v, kp.k_hi, v * kp.k_lo));
    }
    const double dd = static_cast<double>(dist);
    if (KIND == GABO_KIND_GAUSS) return exp_neg_arg(-kp.param * dd * dd);
    return exp_neg_arg(-kp.param * dd);
}

// Tile enumeration.  General: t = jb * tiles_i + ib (rows fastest, so a CTA keeps its column block).  Symmetric
// (x1 is x2): only the tiles that touch the upper triangle are enumerated, so the contiguous per-CTA ranges stay
// balanced: for column block jb the row super-blocks sb = 0..jb (128 rows each, R = 128 / tile_m tiles per super-block).
struct TileMap {
    int64_t tiles_i;
    int tile_m;
    int symmetric;
    __host__ __device__ void decode(int64_t t, int64_t& jb, int64_t& i0) const {
        if (!symmetric) {
            jb = t / tiles_i;
            i0 = (t % tiles_i) * tile_m;
        } else {
            const int64_t R = kThreads / tile_m;
            const int64_t u = t / R;
            const int64_t r = t % R;
            int64_t q = static_cast<int64_t>((sqrt(8.0 * static_cast<double>(u) + 1.0) - 1.0) * 0.5);
            while ((q + 1) * (q + 2) / 2 <= u) ++q;
            while (q * (q + 1) / 2 > u) --q;
            jb = q;
            i0 = ((u - q * (q + 1) / 2) * R + r) * tile_m;
        }
    }
};

// (column block, first row) of a tile id drawn from the dynamic scheduler.
struct TileCursor {
    int64_t jb, i0, i_end;   // column block, first row of the tile (i_end unused)
    // 32-bit form for the dynamic scheduler (tile ids and tile counts are < 2^31; tile_m is a power of two)
    __device__ void start32(const TileMap& map, unsigned int t) {
        if (!map.symmetric) {
            const unsigned int ti = static_cast<unsigned int>(map.tiles_i);
            const unsigned int q = t / ti;
            jb = q;
            i0 = static_cast<int64_t>(t - q * ti) * map.tile_m;
        } else {
            const unsigned int R = kThreads / map.tile_m;
            const unsigned int u = t / R, r = t - u * R;
            unsigned int q = static_cast<unsigned int>((sqrtf(8.0f * static_cast<float>(u) + 1.0f) - 1.0f) * 0.5f);
            while ((q + 1) * (q + 2) / 2 <= u) ++q;
            while (q * (q + 1) / 2 > u) --q;
            jb = q;
            i0 = static_cast<int64_t>((u - q * (q + 1) / 2) * R + r) * map.tile_m;
        }
        i_end = 0;
    }
};

// Jacobi route for the (rare) pairs whose scales fall outside closed_form_in_range: kept out of line so that the hot
// loop of the closed-form kernels does not carry its registers.
template <int d>
__device__ __noinline__ float2 jacobi_log2_sq_x2_slow(float2 g00, float2 g10, float2 g11, float2 g20, float2 g21,
                                                      float2 g22) {   // by value: G stays in registers at the call site
    const float2 z = make_float2(0.0f, 0.0f);
    float2 G[d][d], lam[d];
    G[0][0] = g00;
    G[0][1] = z;
    G[1][0] = g10;
    G[1][1] = g11;
    if constexpr (d == 3) {
        G[0][2] = z;
        G[1][2] = z;
        G[2][0] = g20;
        G[2][1] = g21;
        G[2][2] = g22;
    }
    jacobi_onesided_x2<d>(G, lam);
    return sum_log2_sq_x2<d>(lam);
}

template <int d>
struct PairCfg {
    static constexpr int kMaxTileM = (d <= 5) ? 32 : 8;  // keeps static shared memory under 48 KB for d = 8
    static constexpr bool kPacked = (d <= 5);            // G for two problems fits the register file
};

#ifndef GABO_SPD_MINBLOCKS
#define GABO_SPD_MINBLOCKS 1
#endif
template <int d, typename T, typename OutT, int KIND>
__global__ void __launch_bounds__(kThreads, GABO_SPD_MINBLOCKS)
    spd_ai_gram_kernel(const double* __restrict__ fac1, int64_t n1, const double* __restrict__ fac2, int64_t n2,
                       ExpParams kp, OutT* __restrict__ out, int64_t ld_out, TileMap map, int64_t tiles_total,
                       unsigned int* __restrict__ sched) {
    constexpr int TRI = tri_size(d);
    constexpr int FS = factor_stride(d);
    constexpr bool kLInRegs = (d <= 5);
    constexpr bool kX2 = PairCfg<d>::kPacked && sizeof(T) == 4;   // two pairs per thread on the fp32x2 pipe
    // d = 2, 3 in fp32: closed-form eigenvalues (spd_common.cuh, closed_form_log2_sq_x2) instead of Jacobi sweeps
    constexpr bool kClosed = kX2 && (d == 2 || d == 3);
    constexpr int kMaxTileM = PairCfg<d>::kMaxTileM;
    __shared__ __align__(16) double fs[2][kMaxTileM * FS];              // staged x1 records (we read the A halves)
    __shared__ double ls[kLInRegs ? 1 : TRI * kThreads];                // L_j, entry-major (conflict-free), d >= 6 only
    __shared__ __align__(8) uint64_t bar[2];

    // Dynamic tile scheduling: the per-pair Jacobi takes a data-dependent number of sweeps, so equal tile COUNTS per CTA
    // left the SMs idle 11 % of the launch (ncu: smsp__cycles_active 0.89 of elapsed at N = 2048).  Thread 0 draws tile
    // ids from a global ticket counter one tile ahead (the id travels through s_tile and the existing per-tile barrier,
    // the tile's rows through the TMA pipeline); the last CTA to leave resets the counters for the next launch.
    __shared__ unsigned int s_tile[2];
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
        s_tile[0] = blockIdx.x;     // the first two tiles of a CTA are static (no ticket round trip on the start-up path);
    }                               // tickets are drawn from the third tile on, offset by 2 gridDim.x
    __syncthreads();
    // Programmatic dependent launch: this kernel may have been started while the factorisation kernel before it in the
    // stream was still running (its launch latency and ramp-up are hidden); nothing above reads that kernel's output.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    uint32_t phase_bits = 0u;
    // stage the rows of the tile under cursor `c` (an empty edge tile stages nothing and nothing is waited for)
    auto issue = [&](const TileCursor& c, int buf) {
        const int rows = static_cast<int>(imax(0, imin(map.tile_m, n1 - c.i0)));
        if (rows > 0 && threadIdx.x == 0) {
            const uint32_t bytes = static_cast<uint32_t>(rows) * FS * sizeof(double);  // FS even -> multiple of 16
            mbar_expect_tx(&bar[buf], bytes);
            tma_load_1d(&fs[buf][0], fac1 + c.i0 * FS, bytes, &bar[buf]);
        }
    };

    // prologue: tile 0 is staged into buffer 0, tile 1 into buffer 1 (ids through s_tile[0], s_tile[1])
    TileCursor cur;
    unsigned int t_cur = s_tile[0];
    if (threadIdx.x == 0) {
        if (t_cur < tiles_total) {
            cur.start32(map, t_cur);
            issue(cur, 0);
        }
        const unsigned int t1 = blockIdx.x + gridDim.x;
        s_tile[1] = t1;
        if (t1 < tiles_total) {
            TileCursor nx;
            nx.start32(map, t1);
            issue(nx, 1);
        }
    }
    int64_t jb_loaded = -1;
    double Lreg[kLInRegs ? TRI : 1];
    int64_t j = 0;
    bool jvalid = false;

    for (int it = 0; t_cur < tiles_total; ++it) {
        const int buf = it & 1;
        cur.start32(map, t_cur);
        const int64_t jb = cur.jb, i0 = cur.i0;
        const int rows = static_cast<int>(imax(0, imin(map.tile_m, n1 - i0)));
        // one barrier per tile, at the END of the iteration (see end_of_tile): it releases fs[buf] / ls and publishes
        // the id of the tile after next
        auto end_of_tile = [&]() {
            __syncthreads();
            const unsigned int t_next = s_tile[buf ^ 1];
            if (threadIdx.x == 0) {
                const unsigned int t_nn = atomicAdd(&sched[0], 1u) + 2u * gridDim.x;
                s_tile[buf] = t_nn;            // read by everyone after the NEXT barrier
                if (t_nn < tiles_total) {
                    TileCursor nx;
                    nx.start32(map, t_nn);
                    issue(nx, buf);
                }
            }
            t_cur = t_next;
        };
        if (rows <= 0) {
            end_of_tile();
            continue;
        }

        if (jb != jb_loaded) {
            jb_loaded = jb;
            j = jb * kThreads + threadIdx.x;
            jvalid = j < n2;
            const double* src = fac2 + imin(j, n2 - 1) * FS;
            if (kLInRegs) {
#pragma unroll
                for (int e = 0; e < (kLInRegs ? TRI : 1); ++e) Lreg[e] = __ldg(src + e);
            } else {
#pragma unroll
                for (int e = 0; e < TRI; ++e) ls[e * kThreads + threadIdx.x] = __ldg(src + e);
            }
        }

        mbar_wait(&bar[buf], (phase_bits >> buf) & 1u);
        phase_bits ^= (1u << buf);

        auto store = [&](int64_t gi, float v) {
            if (!jvalid) return;
            if (!map.symmetric) {
                st_cs(out + gi * ld_out + j, static_cast<OutT>(v));
            } else if (j >= gi) {
                out[gi * ld_out + j] = static_cast<OutT>(v);
                if (j > gi) out[j * ld_out + gi] = static_cast<OutT>(v);
            }
        };
        auto Lj = [&](int e) { return kLInRegs ? Lreg[kLInRegs ? e : 0] : ls[e * kThreads + threadIdx.x]; };

        if constexpr (kClosed) {
            // two rows of the tile per step, same column; the triangular product G = A_i L_j (and, for d = 3, the one
            // adjugate entry with a cancellation) in fp64, everything else packed fp32x2 (closed_form_log2_sq_x2)
            const bool sym = map.symmetric != 0;
            OutT* o1 = out + i0 * ld_out + j;
            OutT* o2 = out + j * ld_out + i0;               // mirrored entry (symmetric builds only)
            auto widen = [](float v) -> OutT {
                if constexpr (sizeof(OutT) == 8) return widen_nonneg(v);   // K, d >= 0: ALU-pipe widening
                else return v;
            };
            for (int i = 0; i < rows; i += 2, o1 += 2 * ld_out, o2 += 2) {
                const int i1 = min(i + 1, rows - 1);            // odd tail: the partner repeats row i (not stored)
                const double* A0 = &fs[buf][i * FS + TRI];
                const double* A1 = &fs[buf][i1 * FS + TRI];
                double g0[TRI], g1[TRI];
#pragma unroll
                for (int r = 0; r < d; ++r) {
#pragma unroll
                    for (int c = 0; c <= r; ++c) {
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int k = c; k <= r; ++k) {
                            const double l = Lj(tri_idx(k, c));
                            s0 = fma(A0[tri_idx(r, k)], l, s0);
                            s1 = fma(A1[tri_idx(r, k)], l, s1);
                        }
                        g0[tri_idx(r, c)] = s0;
                        g1[tri_idx(r, c)] = s1;
                    }
                }
                float2 G[TRI];
#pragma unroll
                for (int e = 0; e < TRI; ++e) G[e] = make_float2(static_cast<float>(g0[e]), static_cast<float>(g1[e]));
                float2 x = make_float2(0.0f, 0.0f);
                if constexpr (d == 3) {
                    x = make_float2(static_cast<float>(fma(g0[1], g0[4], -(g0[3] * g0[2]))),
                                    static_cast<float>(fma(g1[1], g1[4], -(g1[3] * g1[2]))));
                }
                bool ok0, ok1;
                float2 ls = closed_form_log2_sq_x2<d>(G, x, ok0, ok1);
                if (!(ok0 && ok1)) {                                        // out-of-range scales: rare
                    if constexpr (d == 3) ls = jacobi_log2_sq_x2_slow<d>(G[0], G[1], G[2], G[3], G[4], G[5]);
                    else ls = jacobi_log2_sq_x2_slow<d>(G[0], G[1], G[2], G[0], G[0], G[0]);
                }
                // spd_utils_torch.py:117-120 in fp32: d^2 = sum log(lambda)^2 + 1e-15, log = ln2 * log2 (MUFU)
                const float2 d2 = fma2(ls, splat2(0.48045301391820142f), splat2(1e-15f));
                float2 v;
                if (KIND == GABO_KIND_GAUSS) {
                    const float2 t = fma2(d2, splat2(kp.k_hi), mul2(d2, splat2(kp.k_lo)));   // kernels_spd.py:96-98
                    v = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                } else {
                    v = make_float2(sqrt_approx(d2.x), sqrt_approx(d2.y));
                    if (KIND == GABO_KIND_LAPLACE) {
                        const float2 t = fma2(v, splat2(kp.k_hi), mul2(v, splat2(kp.k_lo)));  // kernels_spd.py:185
                        v = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                    }
                }
                if (jvalid) {
                    if (!sym) {
                        st_cs(o1, widen(v.x));
                        if (i1 != i) st_cs(o1 + ld_out, widen(v.y));
                    } else {
                        const int64_t gi = i0 + i;
                        if (j >= gi) {
                            o1[0] = widen(v.x);
                            if (j > gi) o2[0] = widen(v.x);
                        }
                        if (i1 != i && j > gi) {
                            o1[ld_out] = widen(v.y);
                            if (j > gi + 1) o2[1] = widen(v.y);
                        }
                    }
                }
            }
        } else if constexpr (kX2) {
            // two rows of the tile per step, same column: (A_i, A_i+1) x L_j on the packed fp32x2 pipe
            for (int i = 0; i < rows; i += 2) {
                const int i1 = min(i + 1, rows - 1);            // odd tail: the partner repeats row i (not stored)
                const double* A0 = &fs[buf][i * FS + TRI];
                const double* A1 = &fs[buf][i1 * FS + TRI];
                float2 G[d][d];
#pragma unroll
                for (int r = 0; r < d; ++r) {
#pragma unroll
                    for (int c = 0; c < d; ++c) {
                        if (c <= r) {
                            double s0 = 0.0, s1 = 0.0;
#pragma unroll
                            for (int k = c; k <= r; ++k) {
                                const double l = Lj(tri_idx(k, c));
                                s0 = fma(A0[tri_idx(r, k)], l, s0);
                                s1 = fma(A1[tri_idx(r, k)], l, s1);
                            }
                            G[r][c] = make_float2(static_cast<float>(s0), static_cast<float>(s1));
                        } else {
                            G[r][c] = make_float2(0.0f, 0.0f);
                        }
                    }
                }
                float2 lam[d];
                jacobi_onesided_x2<d>(G, lam);
                // spd_utils_torch.py:117-120 in fp32: d^2 = sum log(lambda)^2 + 1e-15, log = ln2 * log2 (MUFU)
                const float2 d2 = fma2(sum_log2_sq_x2<d>(lam), splat2(0.48045301391820142f), splat2(1e-15f));
                float2 v;
                if (KIND == GABO_KIND_GAUSS) {
                    const float2 t = fma2(d2, splat2(kp.k_hi), mul2(d2, splat2(kp.k_lo)));   // kernels_spd.py:96-98
                    v = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                } else {
                    v = make_float2(sqrt_approx(d2.x), sqrt_approx(d2.y));
                    if (KIND == GABO_KIND_LAPLACE) {
                        const float2 t = fma2(v, splat2(kp.k_hi), mul2(v, splat2(kp.k_lo)));  // kernels_spd.py:185
                        v = make_float2(ex2_approx(t.x), ex2_approx(t.y));
                    }
                }
                store(i0 + i, v.x);
                if (i1 != i) store(i0 + i1, v.y);
            }
        } else {
            for (int i = 0; i < rows; ++i) {
                const double* Ai = &fs[buf][i * FS + TRI];
                T G[d][d];
                tri_product<d, T>([&](int e) { return Ai[e]; }, Lj, G);
                T lam[d];
                // throughput-bound here: the row-cyclic order with its per-rotation skip does less work than the
                // branch-free round-robin form (which wins where latency binds, in the acquisition kernel)
                if constexpr (sizeof(T) == 4 && d >= 8) {
                    jacobi_onesided_cyclic_rows2<d>(G, lam);     // fp32: rows packed on the fp32x2 pipe
                } else {
                    jacobi_onesided_cyclic<d, T>(G, lam);
                }
                store(i0 + i, finish<KIND, T>(ai_distance_from_eigs<d, T>(lam), kp));
            }
        }
        end_of_tile();
    }
    // the last CTA to leave resets the ticket and exit counters (every CTA's final draw is already behind it)
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
            sched[0] = 0u;
            sched[1] = 0u;
            __threadfence();
        }
    }
}

// Ticket / exit counters of the dynamic tile scheduler: 64 slots used round-robin so that launches in flight on
// different streams do not share a pair; a slot is zero whenever no launch owns it (the kernel resets it on exit).
__device__ unsigned int g_sched[64][2];

unsigned int* sched_slot() {
    static std::atomic<unsigned int> next{0};
    void* p = nullptr;   // per-device address of the symbol (a lookup in the runtime's module table, no device work)
    if (cudaGetSymbolAddress(&p, g_sched) != cudaSuccess) return nullptr;
    return static_cast<unsigned int*>(p) + 2 * (next.fetch_add(1) % 64);
}

template <int d, typename T, typename OutT, int KIND>
int launch_pair(const double* fac1, int64_t n1, const double* fac2, int64_t n2, double param, void* out,
                int64_t ld_out, int symmetric, cudaStream_t stream) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spd_ai_gram_kernel<d, T, OutT, KIND>, kThreads, 0);
    if (occ < 1) occ = 1;
    const int64_t slots = static_cast<int64_t>(sm_count()) * occ;
    const int64_t tiles_j = (n2 + kThreads - 1) / kThreads;
    auto count = [&](int tm) {
        if (!symmetric) return ((n1 + tm - 1) / tm) * tiles_j;
        return (kThreads / tm) * (tiles_j * (tiles_j + 1) / 2);
    };
    // rows per tile from a two-term cost model: waves of tiles over the resident CTAs (quantisation) x rows per tile plus a
    // per-tile overhead worth ~kTileOverheadRows rows (barrier, ticket, TMA turn-around).  Measured at SPD(3) with
    // scripts/micro/spd_variants.cu: N = 2048 takes 16-row tiles (32.2 us) over 8-row (33.8) and 4-row (37.2) ones.
    int tile_m = 2;
    int64_t best_cost = -1;
    for (int tm = PairCfg<d>::kMaxTileM; tm >= 2; tm >>= 1) {
        const int64_t waves = (count(tm) + slots - 1) / slots;
        const int64_t cost = waves * (tm + kTileOverheadRows);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            tile_m = tm;
        }
    }
    TileMap map;
    map.tile_m = tile_m;
    map.tiles_i = (n1 + tile_m - 1) / tile_m;
    map.symmetric = symmetric ? 1 : 0;
    const int64_t tiles = count(tile_m);
    const int64_t grid = imin(tiles, slots);
    ExpParams kp;
    const double k2 = -param * 1.4426950408889634074;
    kp.k_hi = static_cast<float>(k2);
    kp.k_lo = static_cast<float>(k2 - static_cast<double>(kp.k_hi));
    kp.param = param;
    unsigned int* sched = sched_slot();
    GABO_REQUIRE(sched != nullptr, GABO_E_CUDA, "spd_ai_gram: cannot resolve the scheduler counters");
    GABO_REQUIRE(tiles < (1ll << 31), GABO_E_UNSUPPORTED, "spd_ai_gram: %lld tiles exceed the 32-bit tile id", (long long)tiles);
    // launched with programmatic stream serialisation: the CTAs may start under the tail of the previous kernel in the stream
    // (the factorisation) and wait at griddepcontrol.wait before their first read of its output
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, spd_ai_gram_kernel<d, T, OutT, KIND>, fac1, n1, fac2, n2, kp,
                                              static_cast<OutT*>(out), ld_out, map, tiles, sched);
    GABO_REQUIRE(le == cudaSuccess, GABO_E_CUDA, "spd_ai_gram_kernel: %s", cudaGetErrorString(le));
    return check_launch("spd_ai_gram_kernel");
}

template <int d, typename T, typename OutT>
int launch_kind(const double* fac1, int64_t n1, const double* fac2, int64_t n2, double param, int kind, void* out,
                int64_t ld_out, int symmetric, cudaStream_t stream) {
    switch (kind) {
        case GABO_KIND_GAUSS:
            return launch_pair<d, T, OutT, GABO_KIND_GAUSS>(fac1, n1, fac2, n2, param, out, ld_out, symmetric, stream);
        case GABO_KIND_LAPLACE:
            return launch_pair<d, T, OutT, GABO_KIND_LAPLACE>(fac1, n1, fac2, n2, param, out, ld_out, symmetric,
                                                              stream);
        default:
            return launch_pair<d, T, OutT, GABO_KIND_DIST>(fac1, n1, fac2, n2, param, out, ld_out, symmetric, stream);
    }
}

template <int d>
int launch_types(const double* fac1, int64_t n1, const double* fac2, int64_t n2, double param, int kind, int compute,
                 void* out, int out_dtype, int64_t ld_out, int symmetric, cudaStream_t stream) {
    if (compute == GABO_F32) {
        if (out_dtype == GABO_F32)
            return launch_kind<d, float, float>(fac1, n1, fac2, n2, param, kind, out, ld_out, symmetric, stream);
        return launch_kind<d, float, double>(fac1, n1, fac2, n2, param, kind, out, ld_out, symmetric, stream);
    }
    if (out_dtype == GABO_F32)
        return launch_kind<d, double, float>(fac1, n1, fac2, n2, param, kind, out, ld_out, symmetric, stream);
    return launch_kind<d, double, double>(fac1, n1, fac2, n2, param, kind, out, ld_out, symmetric, stream);
}

}  // namespace

}  // namespace gabo

extern "C" int64_t gabo_spd_factor_stride(int d) {
    if (d < 1 || d > GABO_MAX_SPD_DIM) return -1;
    return gabo::factor_stride(d);
}

namespace gabo {
namespace {
int launch_factor(const double* x1, int64_t n1, const double* x2, int64_t n2, int d, int input_is_mandel, double* fac1,
                  double* fac2, int32_t* flags, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned>((n1 + n2 + 127) / 128);
    switch (d) {
#define GABO_CASE(DD)                                                                                       \
    case DD:                                                                                                \
        spd_factor_kernel<DD><<<grid, 128, 0, s>>>(x1, n1, x2, n2, input_is_mandel, fac1, fac2, flags);      \
        break;
        GABO_CASE(1)
        GABO_CASE(2)
        GABO_CASE(3)
        GABO_CASE(4)
        GABO_CASE(5)
        GABO_CASE(6)
        GABO_CASE(7)
        GABO_CASE(8)
#undef GABO_CASE
    }
    return check_launch("spd_factor_kernel");
}
}  // namespace
}  // namespace gabo

extern "C" int gabo_spd_factor(const double* x, int64_t n, int d, int input_is_mandel, double* fac, int32_t* flags,
                               void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_spd_factor: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x && fac, GABO_E_ARG, "gabo_spd_factor: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_factor: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    GABO_REQUIRE(aligned16(fac), GABO_E_ALIGN, "gabo_spd_factor: fac must be 16-byte aligned");
    return launch_factor(x, n, nullptr, 0, d, input_is_mandel, fac, nullptr, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_spd_factor2(const double* x1, int64_t n1, const double* x2, int64_t n2, int d, int input_is_mandel,
                                double* fac1, double* fac2, int32_t* flags, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n1 >= 0 && n2 >= 0, GABO_E_ARG, "gabo_spd_factor2: negative size");
    if (n1 + n2 == 0) return GABO_OK;
    GABO_REQUIRE((n1 == 0 || (x1 && fac1)) && (n2 == 0 || (x2 && fac2)), GABO_E_ARG, "gabo_spd_factor2: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_factor2: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    GABO_REQUIRE(aligned16(fac1) && aligned16(fac2), GABO_E_ALIGN, "gabo_spd_factor2: fac must be 16-byte aligned");
    return launch_factor(x1, n1, x2, n2, d, input_is_mandel, fac1, fac2, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_spd_ai_gram(const double* fac1, int64_t n1, const double* fac2, int64_t n2, int d, double param,
                                int kind, int compute, int symmetric, void* out, int out_dtype, int64_t ld_out,
                                void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n1 >= 0 && n2 >= 0, GABO_E_ARG, "gabo_spd_ai_gram: negative size");
    if (n1 == 0 || n2 == 0) return GABO_OK;
    GABO_REQUIRE(fac1 && fac2 && out, GABO_E_ARG, "gabo_spd_ai_gram: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_ai_gram: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    GABO_REQUIRE(kind >= GABO_KIND_GAUSS && kind <= GABO_KIND_DIST, GABO_E_ARG, "gabo_spd_ai_gram: bad kind %d", kind);
    GABO_REQUIRE(compute == GABO_F32 || compute == GABO_F64, GABO_E_ARG, "gabo_spd_ai_gram: bad compute dtype");
    GABO_REQUIRE(out_dtype == GABO_F32 || out_dtype == GABO_F64, GABO_E_ARG, "gabo_spd_ai_gram: bad out_dtype");
    GABO_REQUIRE(ld_out >= n2, GABO_E_ARG, "gabo_spd_ai_gram: ld_out < n2");
    GABO_REQUIRE(aligned16(fac1) && aligned16(fac2), GABO_E_ALIGN, "gabo_spd_ai_gram: factors must be 16-byte aligned");
    GABO_REQUIRE(!symmetric || (fac1 == fac2 && n1 == n2), GABO_E_ARG,
                 "gabo_spd_ai_gram: symmetric needs fac1 == fac2 and n1 == n2");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
#define GABO_CASE(DD) \
    case DD:          \
        return launch_types<DD>(fac1, n1, fac2, n2, param, kind, compute, out, out_dtype, ld_out, symmetric, s);
        GABO_CASE(1)
        GABO_CASE(2)
        GABO_CASE(3)
        GABO_CASE(4)
        GABO_CASE(5)
        GABO_CASE(6)
        GABO_CASE(7)
        GABO_CASE(8)
#undef GABO_CASE
    }
    return GABO_E_ARG;
}
