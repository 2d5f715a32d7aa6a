// Small-matrix device routines shared by the SPD Gram kernel, the batched manifold operations and the acquisition
// optimiser.  Everything is templated on the matrix size d (<= 8) so that the state lives in registers.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace gabo {

__host__ __device__ constexpr int tri_size(int d) { return d * (d + 1) / 2; }
// per-point record written by gabo_spd_factor: [L (tri) | A = L^-1 (tri) | pad to an even count]
__host__ __device__ constexpr int factor_stride(int d) { return ((2 * tri_size(d) + 1) / 2) * 2; }
__host__ __device__ constexpr int tri_idx(int r, int c) { return r * (r + 1) / 2 + c; }  // c <= r

// Mandel position of entry (r, c), r <= c: diagonal-by-diagonal layout of spd_utils_torch.py:181-187.
__host__ __device__ constexpr int mandel_pos(int d, int r, int c) {
    // k = c - r is the diagonal; diagonals 0..k-1 hold d, d-1, ..., d-k+1 entries
    return (c - r) * d - ((c - r) * ((c - r) - 1)) / 2 + r;
}

// Cholesky X = L L^T and A = L^-1, fp64, packed lower-triangular row-major.  Returns false when a pivot is not
// positive (torch.cholesky raises there, spd_utils_torch.py:87).  X is read through the accessor x(r, c), r >= c.
template <int d, typename Acc>
__device__ __forceinline__ bool chol_inv(Acc x, double (&L)[tri_size(d)], double (&A)[tri_size(d)]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        double s = x(j, j);
#pragma unroll
        for (int k = 0; k < j; ++k) s = fma(-L[tri_idx(j, k)], L[tri_idx(j, k)], s);
        ok = ok && (s > 0.0);
        const double ljj = sqrt(s);
        const double inv = 1.0 / ljj;
        L[tri_idx(j, j)] = ljj;
        A[tri_idx(j, j)] = inv;
#pragma unroll
        for (int i = j + 1; i < d; ++i) {
            double t = x(i, j);
#pragma unroll
            for (int k = 0; k < j; ++k) t = fma(-L[tri_idx(i, k)], L[tri_idx(j, k)], t);
            L[tri_idx(i, j)] = t * inv;
        }
    }
    // A = L^-1 by forward substitution, column by column
#pragma unroll
    for (int j = 0; j < d; ++j) {
#pragma unroll
        for (int i = j + 1; i < d; ++i) {
            double t = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) t = fma(L[tri_idx(i, k)], A[tri_idx(k, j)], t);
            A[tri_idx(i, j)] = -t * A[tri_idx(i, i)];
        }
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi on the columns of G (d x d): G V = U Sigma.  Works on the Cholesky-factor product
// G = L_x^-1 L_y instead of the whitened matrix W = G G^T, which keeps high RELATIVE accuracy of the small
// eigenvalues (error ~ eps * cond(G) = eps * sqrt(cond(W))).  On return the columns of G are u_k sigma_k and
// lam[k] = sigma_k^2 are the eigenvalues of W; log/exp maps are then sum_k f(lam_k)/lam_k g_k g_k^T.
// Converged rotations are skipped, so the result of a pair does not depend on how many extra sweeps its warp runs.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct JacobiTraits;

// XOR the sign bit of `s` into `v` (one LOP3).
__device__ __forceinline__ float xor_sign(float v, float s) {
    return __int_as_float(__float_as_int(v) ^ (__float_as_int(s) & 0x80000000));
}

template <>
struct JacobiTraits<float> {
    static constexpr int kMaxSweeps = 10;
    static __device__ __forceinline__ float tol2() { return 3.6e-15f; }  // (6e-8)^2
    // Jacobi rotation for the 2x2 Gram block [[a, c], [c, b]] with TWO MUFU ops (no division):
    //   h = b - a, r = sqrt(h^2 + 4c^2):  cos(2 theta) = |h| / r,  2 cos^2(theta) = (|h| + r) / r =: q2,
    //   cs = sqrt(q2 / 2),  sn = sign(h) c / (r cs)   (|theta| <= pi/4, the inner rotation),
    // both from rsqrt.approx of (h^2 + 4c^2) and of q2.  The MUFU results carry ~2 ulp, so (cs, sn) is renormalised
    // with one Newton step on cs^2 + sn^2 = 1: the rotation is orthogonal to ~1e-14 and does not rescale the columns.
    // tc = tan(theta) * c is the amount the squared column norms move: a' = a - tc, b' = b + tc.
    static __device__ __forceinline__ void rotation(float a, float b, float c, float& cs, float& sn, float& tc) {
        const float h = b - a;
        const float c2 = c + c;
        const float m = fmaf(h, h, c2 * c2);
        const float ri = rsqrt_approx(m);
        const float q2 = fmaf(m, ri, fabsf(h)) * ri;               // in [1, 2]
        const float iq = rsqrt_approx(q2);                          // 1 / (sqrt2 cs)
        const float cs0 = (q2 * iq) * 0.70710678118654752f;
        const float sn0 = xor_sign((c2 * ri) * (iq * 0.70710678118654752f), h);
        const float f = fmaf(-0.5f, fmaf(cs0, cs0, sn0 * sn0), 1.5f);
        cs = cs0 * f;
        sn = sn0 * f;
        tc = (sn0 * (iq * 1.41421356237309505f)) * c;
    }
    static __device__ __forceinline__ float log_(float x) { return __logf(x); }          // lg2.approx: |err| <= 1.7e-7
    static __device__ __forceinline__ float dist_(float s) { return sqrt_approx(s + 1e-15f); }
};

template <>
struct JacobiTraits<double> {
    static constexpr int kMaxSweeps = 12;
    static __device__ __forceinline__ double tol2() { return 1e-26; }  // (1e-13)^2
    static __device__ __forceinline__ void rotation(double a, double b, double c, double& cs, double& sn, double& tc) {
        // The angle only has to be approximately the Jacobi angle (it sets the convergence rate, not the accuracy), so
        // it is computed on the fp32 MUFU path; cs is then normalised in fp64 so that the rotation is orthogonal to 1e-16.
        const float zf = static_cast<float>(b - a) * rcp_approx(2.0f * static_cast<float>(c));
        const float tf = copysignf(1.0f, zf) * rcp_approx(fabsf(zf) + sqrt_approx(fmaf(zf, zf, 1.0f)));
        const double t = static_cast<double>(tf);
        cs = rsqrt(fma(t, t, 1.0));
        sn = t * cs;
        // exact norm update for an inexact angle: a' = cs^2 (a + t^2 b - 2 t c); expressed as a - tc_p below would
        // need two values, so the caller recomputes the norms from the columns in fp64 (see jacobi_onesided).
        tc = t * c;
    }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float dist_(float s) { return sqrtf(s + 1e-15f); }
};

// Round-robin ("tournament") ordering of the index pairs of a sweep: m - 1 rounds of m / 2 DISJOINT pairs
// (m = d rounded up to even; pairs that involve the dummy index d are skipped).  The rotations of one round touch
// different columns, so a thread can overlap them -- the row-cyclic order (0,1),(0,2),... makes every rotation depend on
// the previous one, and both the SPD(8) Gram kernel and the acquisition kernel are bound by that dependency chain.
template <int d>
struct RoundRobin {
    static constexpr int m = (d % 2 == 0) ? d : d + 1;
    static constexpr int kRounds = m - 1;
    static constexpr int kPairs = m / 2;
    static __host__ __device__ constexpr int player(int round, int i) {
        return i == 0 ? 0 : 1 + ((i - 1 + round) % (m - 1));
    }
    static __host__ __device__ constexpr int lo(int round, int j) {
        return player(round, j) < player(round, m - 1 - j) ? player(round, j) : player(round, m - 1 - j);
    }
    static __host__ __device__ constexpr int hi(int round, int j) {
        return player(round, j) < player(round, m - 1 - j) ? player(round, m - 1 - j) : player(round, j);
    }
};

// One-sided Jacobi with the round-robin ordering, branch-free inside a round: the m / 2 inner products, rotation
// parameters and column updates of a round are independent instruction streams (rotations that are already converged
// get the exact identity by selection, so a lane's result still does not depend on its warp neighbours).
template <int d, typename T>
__device__ __forceinline__ void jacobi_onesided_rr(T (&G)[d][d], T (&lam)[d]) {
    using Tr = JacobiTraits<T>;
    using RR = RoundRobin<d>;
    constexpr bool kF32 = sizeof(T) == 4;
#pragma unroll
    for (int k = 0; k < d; ++k) {
        T s = T(0);
#pragma unroll
        for (int r = 0; r < d; ++r) s = fma(G[r][k], G[r][k], s);
        lam[k] = s;
    }
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int round = 0; round < RR::kRounds; ++round) {
            T c[RR::kPairs];
            bool need[RR::kPairs];
            bool any = false;
#pragma unroll
            for (int j = 0; j < RR::kPairs; ++j) {
                constexpr int dummy = d;  // index d does not exist when d is odd
                const int p = RR::lo(round, j), q = RR::hi(round, j);
                c[j] = T(0);
                need[j] = false;
                if (q < dummy) {
                    T s = T(0);
#pragma unroll
                    for (int r = 0; r < d; ++r) s = fma(G[r][p], G[r][q], s);
                    c[j] = s;
                    need[j] = s * s > Tr::tol2() * (lam[p] * lam[q]);
                    any = any || need[j];
                }
            }
            if (any) {
                rotated = true;
#pragma unroll
                for (int j = 0; j < RR::kPairs; ++j) {
                    const int p = RR::lo(round, j), q = RR::hi(round, j);
                    if (q < d) {
                        const T a = lam[p], b = lam[q];
                        T cs, sn, tc;
                        Tr::rotation(a, b, c[j], cs, sn, tc);
                        cs = need[j] ? cs : T(1);
                        sn = need[j] ? sn : T(0);
                        tc = need[j] ? tc : T(0);
#pragma unroll
                        for (int r = 0; r < d; ++r) {
                            const T gp = G[r][p], gq = G[r][q];
                            G[r][p] = fma(cs, gp, -sn * gq);
                            G[r][q] = fma(sn, gp, cs * gq);
                        }
                        if (kF32) {
                            lam[p] = a - tc;
                            lam[q] = b + tc;
                        } else {  // the fp64 angle is approximate: exact expression for the rotated norms
                            const T c2 = cs * cs, s2 = sn * sn, x = T(2) * cs * sn * c[j];
                            lam[p] = fma(c2, a, fma(s2, b, -x));
                            lam[q] = fma(s2, a, fma(c2, b, x));
                        }
                    }
                }
            }
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
#pragma unroll
    for (int k = 0; k < d; ++k) {
        T s = T(0);
#pragma unroll
        for (int r = 0; r < d; ++r) s = fma(G[r][k], G[r][k], s);
        lam[k] = s;
    }
}

// Compact sweep for kernels that inline the solve into a large body (the acquisition kernels): ODD-EVEN TRANSPOSITION
// ordering with rotate-and-interchange.  A sweep is m / 2 passes over the same two rounds,
//     even round: (0,1) (2,3) ... (m-2,m-1)        odd round: (1,2) (3,4) ... (m-3,m-2),
// and every rotation also EXCHANGES its two columns (free: the rotated columns are simply stored crosswise).  As in
// odd-even transposition sort, after m rounds the column order is reversed and every pair of columns has met exactly
// once.  The loop body is two rounds (~500 instructions, m - 1 rotations) and needs no register moves; the fully
// unrolled round-robin sweep is ~2000 instructions and did not fit the instruction caches once inlined in the
// acquisition kernel (ncu: stall_no_instruction 4-6 warps per issue), and a round-robin loop with an explicit column
// permutation spent 29 % of its instructions on MOVs.  Branch-free inside a round (independent rotations overlap);
// converged pairs get the exact identity (+ exchange).  Odd d gets a zero dummy column.
template <int d, typename T>
__device__ __forceinline__ void jacobi_onesided_loop(T (&G)[d][d], T (&lam)[d]) {
    using Tr = JacobiTraits<T>;
    constexpr bool kF32 = sizeof(T) == 4;
    constexpr int m = d + (d & 1);
    T C[d][m], l[m];
#pragma unroll
    for (int k = 0; k < m; ++k) {
        T s = T(0);
#pragma unroll
        for (int r = 0; r < d; ++r) {
            C[r][k] = (k < d) ? G[r][k < d ? k : 0] : T(0);
            s = fma(C[r][k], C[r][k], s);
        }
        l[k] = s;
    }
    // one round over the pairs (first, first+1), (first+2, first+3), ...; returns whether this lane rotated
    auto round = [&](auto first_tag) -> bool {
        constexpr int first = decltype(first_tag)::value;
        constexpr int np = (m - first) / 2;
        T c[np > 0 ? np : 1];
        bool need[np > 0 ? np : 1];
        bool any = false;
#pragma unroll
        for (int j = 0; j < np; ++j) {
            const int p = first + 2 * j, q = p + 1;
            T s = T(0);
#pragma unroll
            for (int r = 0; r < d; ++r) s = fma(C[r][p], C[r][q], s);
            c[j] = s;
            need[j] = s * s > Tr::tol2() * (l[p] * l[q]);
            any = any || need[j];
        }
#pragma unroll
        for (int j = 0; j < np; ++j) {
            const int p = first + 2 * j, q = p + 1;
            const T a = l[p], b = l[q];
            T cs, sn, tc;
            Tr::rotation(a, b, c[j], cs, sn, tc);
            cs = need[j] ? cs : T(1);
            sn = need[j] ? sn : T(0);
            tc = need[j] ? tc : T(0);
#pragma unroll
            for (int r = 0; r < d; ++r) {
                const T gp = C[r][p], gq = C[r][q];
                C[r][q] = fma(cs, gp, -sn * gq);      // rotated column p, stored at q
                C[r][p] = fma(sn, gp, cs * gq);       // rotated column q, stored at p
            }
            if (kF32) {
                l[q] = a - tc;
                l[p] = b + tc;
            } else {  // the fp64 angle is approximate: exact expression for the rotated norms
                const T c2 = cs * cs, s2 = sn * sn, x = T(2) * cs * sn * c[j];
                l[q] = fma(c2, a, fma(s2, b, -x));
                l[p] = fma(s2, a, fma(c2, b, x));
            }
        }
        return any;
    };
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll 1
        for (int pass = 0; pass < m / 2; ++pass) {
            rotated = round(std::integral_constant<int, 0>{}) || rotated;
            rotated = round(std::integral_constant<int, 1>{}) || rotated;
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
    // the dummy column of an odd d is wherever the exchanges left it: it is the (only) zero column
#pragma unroll
    for (int k = 0; k < d; ++k) lam[k] = T(0);
    if (d == m) {
#pragma unroll
        for (int k = 0; k < d; ++k) {
            T s = T(0);
#pragma unroll
            for (int r = 0; r < d; ++r) {
                G[r][k] = C[r][k];
                s = fma(C[r][k], C[r][k], s);
            }
            lam[k] = s;
        }
    } else {
        // after an even number of sweeps the order is the original one, after an odd number it is reversed: the dummy
        // sits at position m-1 or 0; pick the d real columns accordingly (warp-uniform: every lane ran the same sweeps)
        T s0 = T(0);
#pragma unroll
        for (int r = 0; r < d; ++r) s0 = fma(C[r][0], C[r][0], s0);
        const bool dummy_first = (s0 == T(0));
#pragma unroll
        for (int k = 0; k < d; ++k) {
            T s = T(0);
#pragma unroll
            for (int r = 0; r < d; ++r) {
                const T v = dummy_first ? C[r][k + 1 < m ? k + 1 : 0] : C[r][k];
                G[r][k] = v;
                s = fma(v, v, s);
            }
            lam[k] = s;
        }
    }
}

// Stopping rule: a rotation is applied while c^2 > tol2 * a * b; the sweeps end when no lane of the warp rotated.
// (Stopping a sweep earlier on the strength of quadratic convergence is NOT safe here: for nearly identical matrices
// W = I + E the iteration converges relative to |E|, not to the diagonal, and d^2 ~ |E|_F^2 needs the off-diagonal part
// resolved to ~1e-7 of the DIAGONAL to keep |d - d_ref| <= 1e-6.)
// Converged rotations are skipped per lane, so the result of a pair does not depend on the other pairs of its warp.
template <int d, typename T>
__device__ __forceinline__ void jacobi_onesided_cyclic(T (&G)[d][d], T (&lam)[d]) {
    using Tr = JacobiTraits<T>;
    constexpr bool kF32 = sizeof(T) == 4;
#pragma unroll
    for (int k = 0; k < d; ++k) {
        T s = T(0);
#pragma unroll
        for (int r = 0; r < d; ++r) s = fma(G[r][k], G[r][k], s);
        lam[k] = s;
    }
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < d - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < d; ++q) {
                T c = T(0);
#pragma unroll
                for (int r = 0; r < d; ++r) c = fma(G[r][p], G[r][q], c);
                const T a = lam[p], b = lam[q];
                if (c * c > Tr::tol2() * (a * b)) {
                    rotated = true;
                    T cs, sn, tc;
                    Tr::rotation(a, b, c, cs, sn, tc);
#pragma unroll
                    for (int r = 0; r < d; ++r) {
                        const T gp = G[r][p], gq = G[r][q];
                        G[r][p] = fma(cs, gp, -sn * gq);
                        G[r][q] = fma(sn, gp, cs * gq);
                    }
                    if (kF32) {
                        lam[p] = a - tc;
                        lam[q] = b + tc;
                    } else {  // the fp64 angle is approximate: use the exact expression for the rotated norms
                        const T c2 = cs * cs, s2 = sn * sn, x = T(2) * cs * sn * c;
                        lam[p] = fma(c2, a, fma(s2, b, -x));
                        lam[q] = fma(s2, a, fma(c2, b, x));
                    }
                }
            }
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
#pragma unroll
    for (int k = 0; k < d; ++k) {
        T s = T(0);
#pragma unroll
        for (int r = 0; r < d; ++r) s = fma(G[r][k], G[r][k], s);
        lam[k] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same one-sided Jacobi for TWO independent problems per thread on the packed fp32x2 pipe (.x = problem 0,
// .y = problem 1).  The Gram kernel is issue-bound and ~60 % of its instructions are fp32 FMA/MUL/ADD; packing two
// pairs halves those issue slots (FFMA2 / FMUL2 / FADD2).  A rotation is executed when EITHER problem needs it; the
// other one gets the exact identity (cs = 1, sn = 0, tc = 0), so each problem sees exactly the rotations it would see
// alone and its result does not depend on its partner.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 abs2(float2 v) { return make_float2(fabsf(v.x), fabsf(v.y)); }
__device__ __forceinline__ float2 rsqrt2(float2 v) { return make_float2(rsqrt_approx(v.x), rsqrt_approx(v.y)); }

template <int d>
__device__ __forceinline__ void jacobi_onesided_x2(float2 (&G)[d][d], float2 (&lam)[d]) {
    using Tr = JacobiTraits<float>;
#pragma unroll
    for (int k = 0; k < d; ++k) {
        float2 s = mul2(G[0][k], G[0][k]);
#pragma unroll
        for (int r = 1; r < d; ++r) s = fma2(G[r][k], G[r][k], s);
        lam[k] = s;
    }
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < d - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < d; ++q) {
                float2 c = mul2(G[0][p], G[0][q]);
#pragma unroll
                for (int r = 1; r < d; ++r) c = fma2(G[r][p], G[r][q], c);
                const float2 a = lam[p], b = lam[q];
                const float2 cc = mul2(c, c), thr = mul2(mul2(a, b), splat2(Tr::tol2()));
                const bool n0 = cc.x > thr.x, n1 = cc.y > thr.y;
                if (n0 || n1) {
                    rotated = true;
                    // JacobiTraits<float>::rotation, two at a time
                    const float2 h = sub2(b, a), c2 = add2(c, c);
                    const float2 m = fma2(h, h, mul2(c2, c2));
                    const float2 ri = rsqrt2(m);
                    const float2 q2 = mul2(fma2(m, ri, abs2(h)), ri);
                    const float2 iqs = mul2(rsqrt2(q2), splat2(0.70710678118654752f));
                    const float2 cs0 = mul2(q2, iqs);
                    float2 sn0 = mul2(mul2(c2, ri), iqs);
                    sn0.x = xor_sign(sn0.x, h.x);
                    sn0.y = xor_sign(sn0.y, h.y);
                    const float2 f = fma2(splat2(-0.5f), fma2(cs0, cs0, mul2(sn0, sn0)), splat2(1.5f));
                    float2 cs = mul2(cs0, f), sn = mul2(sn0, f);
                    float2 tc = mul2(mul2(sn0, add2(iqs, iqs)), c);
                    if (!n0) { cs.x = 1.0f; sn.x = 0.0f; tc.x = 0.0f; }
                    if (!n1) { cs.y = 1.0f; sn.y = 0.0f; tc.y = 0.0f; }
                    const float2 nsn = make_float2(-sn.x, -sn.y);
#pragma unroll
                    for (int r = 0; r < d; ++r) {
                        const float2 gp = G[r][p], gq = G[r][q];
                        G[r][p] = fma2(cs, gp, mul2(nsn, gq));
                        G[r][q] = fma2(sn, gp, mul2(cs, gq));
                    }
                    lam[p] = sub2(a, tc);
                    lam[q] = add2(b, tc);
                }
            }
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
#pragma unroll
    for (int k = 0; k < d; ++k) {
        float2 s = mul2(G[0][k], G[0][k]);
#pragma unroll
        for (int r = 1; r < d; ++r) s = fma2(G[r][k], G[r][k], s);
        lam[k] = s;
    }
}

// sum_k log2(lambda_k)^2 for both problems (the natural-log scale ln2^2 is applied by the caller).
template <int d>
__device__ __forceinline__ float2 sum_log2_sq_x2(const float2 (&lam)[d]) {
    float2 s = splat2(0.0f);
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const float2 l = make_float2(__log2f(lam[k].x), __log2f(lam[k].y));
        s = fma2(l, l, s);
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Closed-form eigenvalues for d = 2, 3 with the RELATIVE accuracy the log-distance needs ("bilateral" form).
// The trigonometric formula gives the eigenvalues of a symmetric 3 x 3 matrix M with an ABSOLUTE error of a few
// eps |M| -- fine for the largest eigenvalue, useless for the smallest one of an ill-conditioned W = G G^T (the
// tolerance on d needs ~1e-5 relative on every eigenvalue).  So only LARGEST eigenvalues are taken from it:
//     lambda_max(W)              from M1 = G G^T,
//     det W / lambda_min(W)      = lambda_max(adj(G)^T adj(G)) from M2: adj(G) = det(G) G^-1 of the lower-triangular G is
//                                  five plain products of its entries plus ONE entry with a cancellation,
//                                  x = g10 g21 - g20 g11, which the caller forms in fp64 before rounding,
//     lambda_mid(W)              = det W / (lambda_max lambda_min),  det W = (g00 g11 g22)^2 exactly (G is triangular).
// When the two largest (or two smallest) eigenvalues nearly coincide the angle of the trigonometric formula is
// ill-conditioned (error ~ sqrt(eps)), but the determinant identity moves the middle eigenvalue by the opposite
// relative amount, and sum log^2 is stationary under such a split: the error enters d^2 only as p^2 * eps-ish
// (scripts/dev_closed3.py checks the bound of SURVEY 8(d) in float32 emulation on the benchmark law, near-identical
// pairs, double / triple eigenvalues, cond up to 5000 per matrix; tests/test_gram_gpu.py does it on the device).
// Two problems per thread on the packed fp32x2 pipe (.x / .y).  Intermediate quantities reach |G|^8: the caller
// checks closed_form_in_range() and sends anything else through the Jacobi sweeps.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 neg2(float2 v) { return make_float2(-v.x, -v.y); }
__device__ __forceinline__ float2 log2_2(float2 v) { return make_float2(lg2_approx(v.x), lg2_approx(v.y)); }

// Largest eigenvalue of the symmetric matrix [[m00 m01 m02], [m01 m11 m12], [m02 m12 m22]]:
//   q = tr/3, p^2 = |M - qI|_F^2 / 6, r = det((M - qI)/p)/2 = cos(3 phi), lambda_max = q + 2 p cos(phi),
//   cos(phi) = g(s) with s = sqrt((1 + r)/2) = cos(3 phi / 2):  g(s) = cos(2/3 acos s) is analytic on [0, 1]
//   (degree-8 polynomial, |error| < 1.2e-7 in fp32), so no acos / cos evaluation is needed.  q is returned as well
//   (range check of the caller).
__device__ __forceinline__ float2 sym3_lam_max_x2(float2 m00, float2 m11, float2 m22, float2 m01, float2 m02,
                                                  float2 m12, float2& q) {
    q = mul2(add2(add2(m00, m11), m22), splat2(0.333333343267440796f));
    const float2 b00 = sub2(m00, q), b11 = sub2(m11, q), b22 = sub2(m22, q);
    float2 off = mul2(m01, m01);
    off = fma2(m02, m02, off);
    off = fma2(m12, m12, off);
    float2 dg = mul2(b00, b00);
    dg = fma2(b11, b11, dg);
    dg = fma2(b22, b22, dg);
    float2 p2 = fma2(off, splat2(0.333333343267440796f), mul2(dg, splat2(0.166666671633720398f)));
    p2 = make_float2(fmaxf(p2.x, 1e-37f), fmaxf(p2.y, 1e-37f));
    const float2 ip = rsqrt2(p2);
    // r = det(B) / (2 p^3) with the UNSCALED B = M - qI (three multiplies instead of six; |det B| <= 2 p^3 stays inside
    // the fp32 range because the caller keeps q in [1e-6, 1e9], and a det that underflows belongs to a p below 2e-13,
    // i.e. to a correction of lambda below 2e-7 q)
    const float2 t0 = fma2(b11, b22, neg2(mul2(m12, m12)));
    const float2 t1 = fma2(m12, m02, neg2(mul2(m01, b22)));
    const float2 t2 = fma2(m01, m12, neg2(mul2(b11, m02)));
    const float2 det = fma2(m02, t2, fma2(m01, t1, mul2(b00, t0)));
    // scaled one factor at a time: ip^3 alone overflows when p^2 sits at its floor (M = qI: det = 0, ip = 3e18)
    const float2 s2 = fma2(mul2(mul2(det, ip), ip), mul2(ip, splat2(0.25f)), splat2(0.5f));   // (1 + r) / 2
    const float2 sv = make_float2(sqrt_approx(__saturatef(s2.x)), sqrt_approx(__saturatef(s2.y)));
    float2 g = splat2(2.0f * -6.393143697e-04f);                           // 2 g(s): lambda = q + p (2 cos phi)
    g = fma2(g, sv, splat2(2.0f * 3.707686486e-03f));
    g = fma2(g, sv, splat2(2.0f * -1.032549309e-02f));
    g = fma2(g, sv, splat2(2.0f * 1.967571822e-02f));
    g = fma2(g, sv, splat2(2.0f * -3.196129547e-02f));
    g = fma2(g, sv, splat2(2.0f * 5.328865600e-02f));
    g = fma2(g, sv, splat2(2.0f * -1.110956833e-01f));
    g = fma2(g, sv, splat2(2.0f * 5.773497198e-01f));
    g = fma2(g, sv, splat2(2.0f * 5.000000033e-01f));
    const float2 p = mul2(p2, ip);                                         // sqrt(p2)
    return fma2(p, g, q);
}

// Both means q1 = tr(M1)/3, q2 = tr(M2)/3 of a lane inside [1e-6, 1e9]: then nothing overflows on the way
// (|det B| <= 2 p^3 <= 1e28) and whatever underflows is below 2e-7 of q (see sym3_lam_max_x2).  NaN fails the test.
// M2's eigenvalues are products of two eigenvalues of W, so the window covers eigenvalues of W in about [1e-3, 3e4];
// pairs outside it take the Jacobi sweeps.
__device__ __forceinline__ bool closed_form_in_range(float q1, float q2) {
    return fminf(q1, q2) >= 1e-6f && fmaxf(q1, q2) <= 1e9f;
}

// sum_k log2(lambda_k(G G^T))^2 for lower-triangular G (packed row-major: g00 | g10 g11 | g20 g21 g22);
// d = 3 also takes x = g10 g21 - g20 g11 (formed in fp64 by the caller).  `ok` = closed_form_in_range per problem.
template <int d>
__device__ __forceinline__ float2 closed_form_log2_sq_x2(const float2 (&G)[tri_size(d)], float2 x, bool& ok0, bool& ok1) {
    static_assert(d == 2 || d == 3, "closed forms exist for d = 2, 3");
    if constexpr (d == 2) {
        const float2 a = mul2(G[0], G[0]), b = fma2(G[1], G[1], mul2(G[2], G[2])), c = mul2(G[0], G[1]);
        const float2 h = sub2(a, b), c2 = add2(c, c);
        const float2 disc = fma2(h, h, mul2(c2, c2));
        const float2 root = make_float2(sqrt_approx(disc.x), sqrt_approx(disc.y));
        const float2 tr = add2(a, b);
        const float2 l1 = log2_2(mul2(add2(tr, root), splat2(0.5f)));              // larger eigenvalue: no cancellation
        const float2 dt = mul2(G[0], G[2]);
        const float2 ld = log2_2(dt);
        const float2 l2 = sub2(add2(ld, ld), l1);                                  // log2(det / lambda_1)
        ok0 = closed_form_in_range(tr.x, dt.x);      // trace and determinant of G inside [1e-9, 1e9]
        ok1 = closed_form_in_range(tr.y, dt.y);
        return fma2(l1, l1, mul2(l2, l2));
    } else {
        // M1 = G G^T
        const float2 m00 = mul2(G[0], G[0]);
        const float2 m11 = fma2(G[1], G[1], mul2(G[2], G[2]));
        const float2 m22 = fma2(G[3], G[3], fma2(G[4], G[4], mul2(G[5], G[5])));
        const float2 m01 = mul2(G[0], G[1]), m02 = mul2(G[0], G[3]);
        const float2 m12 = fma2(G[1], G[3], mul2(G[2], G[4]));
        float2 q1, q2;
        const float2 l1 = sym3_lam_max_x2(m00, m11, m22, m01, m02, m12, q1);
        // H = D adj(G) D with D = diag(1, -1, 1) (same singular values, no sign flips):
        //   h00 = g11 g22, h11 = g00 g22, h22 = g00 g11, h10 = g10 g22, h21 = g21 g00, h20 = x;   M2 = H^T H
        const float2 h00 = mul2(G[2], G[5]), h11 = mul2(G[0], G[5]), h22 = mul2(G[0], G[2]);
        const float2 h10 = mul2(G[1], G[5]), h21 = mul2(G[4], G[0]);
        const float2 n00 = fma2(h00, h00, fma2(h10, h10, mul2(x, x)));
        const float2 n11 = fma2(h11, h11, mul2(h21, h21));
        const float2 n22 = mul2(h22, h22);
        const float2 n01 = fma2(h10, h11, mul2(x, h21));
        const float2 n02 = mul2(x, h22), n12 = mul2(h21, h22);
        const float2 mu = sym3_lam_max_x2(n00, n11, n22, n01, n02, n12, q2);       // det W / lambda_min
        ok0 = closed_form_in_range(q1.x, q2.x);
        ok1 = closed_form_in_range(q1.y, q2.y);
        const float2 ld = log2_2(mul2(h22, G[5]));                                 // log2 det G
        const float2 ld2 = add2(ld, ld);                                           // log2 det W
        const float2 a = log2_2(l1);                                               // log2 lambda_max
        const float2 c = sub2(ld2, log2_2(mu));                                    // log2 lambda_min
        const float2 b = sub2(sub2(ld2, a), c);                                    // log2 lambda_mid
        return fma2(a, a, fma2(b, b, mul2(c, c)));
    }
}

// G = A * L for packed lower-triangular A (rows) and L, accumulated in fp64, returned in T.
template <int d, typename T, typename AccA, typename AccL>
__device__ __forceinline__ void tri_product(AccA A, AccL L, T (&G)[d][d]) {
#pragma unroll
    for (int r = 0; r < d; ++r) {
#pragma unroll
        for (int c = 0; c < d; ++c) {
            if (c <= r) {
                double s = 0.0;
#pragma unroll
                for (int k = c; k <= r; ++k) s = fma(A(tri_idx(r, k)), L(tri_idx(k, c)), s);
                G[r][c] = static_cast<T>(s);
            } else {
                G[r][c] = T(0);
            }
        }
    }
}

// Row-cyclic one-sided Jacobi (fp32) with the ROWS of G packed two per register pair: inner products and column
// updates run on the packed fp32x2 pipe (FFMA2 / FMUL2), halving their issue slots.  For the single-problem-per-thread
// sizes (d >= 6, where two problems per thread would not fit the register file) the Gram kernel is issue-bound with
// as many FMULs as FFMAs in the rotation updates (ncu: 63 % issue-active, FMA pipe 49 %).  An odd d gets a zero row.
template <int d>
__device__ __forceinline__ void jacobi_onesided_cyclic_rows2(float (&G)[d][d], float (&lam)[d]) {
    using Tr = JacobiTraits<float>;
    constexpr int R2 = (d + 1) / 2;
    float2 H[R2][d];
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int k = 0; k < d; ++k) H[r][k] = make_float2(G[2 * r][k], (2 * r + 1 < d) ? G[(2 * r + 1 < d) ? 2 * r + 1 : 0][k] : 0.0f);
    auto dot2 = [&](int p, int q) {
        float2 s = mul2(H[0][p], H[0][q]);
#pragma unroll
        for (int r = 1; r < R2; ++r) s = fma2(H[r][p], H[r][q], s);
        return s.x + s.y;
    };
#pragma unroll
    for (int k = 0; k < d; ++k) lam[k] = dot2(k, k);
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < d - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < d; ++q) {
                const float c = dot2(p, q);
                const float a = lam[p], b = lam[q];
                if (c * c > Tr::tol2() * (a * b)) {
                    rotated = true;
                    float cs, sn, tc;
                    Tr::rotation(a, b, c, cs, sn, tc);
                    const float2 cs2 = splat2(cs), sn2 = splat2(sn), nsn2 = splat2(-sn);
#pragma unroll
                    for (int r = 0; r < R2; ++r) {
                        const float2 gp = H[r][p], gq = H[r][q];
                        H[r][p] = fma2(cs2, gp, mul2(nsn2, gq));
                        H[r][q] = fma2(sn2, gp, mul2(cs2, gq));
                    }
                    lam[p] = a - tc;
                    lam[q] = b + tc;
                }
            }
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
#pragma unroll
    for (int k = 0; k < d; ++k) lam[k] = dot2(k, k);
#pragma unroll
    for (int r = 0; r < R2; ++r)
#pragma unroll
        for (int k = 0; k < d; ++k) {
            G[2 * r][k] = H[r][k].x;
            if (2 * r + 1 < d) G[(2 * r + 1 < d) ? 2 * r + 1 : 0][k] = H[r][k].y;
        }
}

// d <= 3 has at most one real pair per round: the row-cyclic form (with its per-rotation skip) is the cheaper one there.
template <int d, typename T>
__device__ __forceinline__ void jacobi_onesided(T (&G)[d][d], T (&lam)[d]) {
    if constexpr (d >= 4) {
        jacobi_onesided_rr<d, T>(G, lam);
    } else {
        jacobi_onesided_cyclic<d, T>(G, lam);
    }
}

// The form for kernels that inline the solve into a large body (acquisition): compact looped rounds for d >= 4.
template <int d, typename T>
__device__ __forceinline__ void jacobi_onesided_compact(T (&G)[d][d], T (&lam)[d]) {
    if constexpr (d >= 4) {
        jacobi_onesided_loop<d, T>(G, lam);
    } else {
        jacobi_onesided_cyclic<d, T>(G, lam);
    }
}

// Reference tail (spd_utils_torch.py:108-120): eigenvalues rounded to fp32, log / square / sum / sqrt(+1e-15) in fp32.
// The fp32 compute path takes the logs and the root on the MUFU (abs. error of d <= 3e-7, inside the 1e-6 floor the
// reference's own float32 eigenvalues leave); the fp64 ("reference-grade") path uses the IEEE-accurate logf / sqrtf.
template <int d, typename T>
__device__ __forceinline__ float ai_distance_from_eigs(const T (&lam)[d]) {
    using Tr = JacobiTraits<T>;
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const float l = Tr::log_(static_cast<float>(lam[k]));
        s = fmaf(l, l, s);
    }
    return Tr::dist_(s);
}


// ---------------------------------------------------------------------------------------------------------------
// Two-sided cyclic Jacobi for a symmetric (possibly indefinite) d x d matrix: S = V diag(lam) V^T.
// Used for the exponential map (expm of a whitened tangent vector) where the matrix is not positive definite, so the
// one-sided form above does not apply.  Only the upper triangle of S is referenced (S[r][c], r <= c).
// Rotations stop when every off-diagonal entry is below tol * ||S||_F: eigenvalues then carry an ABSOLUTE error of
// order eps * ||S||, which is what exp(lam) needs.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct SymJacobiTraits;
template <>
struct SymJacobiTraits<float> {
    static constexpr int kMaxSweeps = 10;
    static __device__ __forceinline__ float tol() { return 6e-8f; }
    static __device__ __forceinline__ float rsqrt_(float x) {
        const float y = rsqrt_approx(x);
        return y * fmaf(-0.5f * x * y, y, 1.5f);
    }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rcp_(float x) { return 1.0f / x; }
};
template <>
struct SymJacobiTraits<double> {
    static constexpr int kMaxSweeps = 14;
    static __device__ __forceinline__ double tol() { return 1e-15; }
    static __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double rcp_(double x) { return 1.0 / x; }
};

#define GABO_SYM(S, r, c) S[((r) < (c)) ? (r) : (c)][((r) < (c)) ? (c) : (r)]

template <int d, typename T, bool kWantV>
__device__ __forceinline__ void jacobi_symmetric(T (&S)[d][d], T (&lam)[d], T (&V)[d][d]) {
    using Tr = SymJacobiTraits<T>;
    if (kWantV) {
#pragma unroll
        for (int r = 0; r < d; ++r)
#pragma unroll
            for (int c = 0; c < d; ++c) V[r][c] = (r == c) ? T(1) : T(0);
    }
    T fro = T(0);
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = r; c < d; ++c) fro = fma((r == c) ? T(1) : T(2), S[r][c] * S[r][c], fro);
    const T thr = Tr::tol() * Tr::sqrt_(fro);
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll
        for (int p = 0; p < d - 1; ++p) {
#pragma unroll
            for (int q = p + 1; q < d; ++q) {
                const T apq = S[p][q];
                if (fabs(apq) > thr) {
                    rotated = true;
                    const T tau = (S[q][q] - S[p][p]) * Tr::rcp_(T(2) * apq);
                    const T t = copysign(T(1), tau) * Tr::rcp_(fabs(tau) + Tr::sqrt_(fma(tau, tau, T(1))));
                    const T cs = Tr::rsqrt_(fma(t, t, T(1)));
                    const T sn = t * cs;
                    S[p][p] = fma(-t, apq, S[p][p]);
                    S[q][q] = fma(t, apq, S[q][q]);
                    S[p][q] = T(0);
#pragma unroll
                    for (int r = 0; r < d; ++r) {
                        if (r != p && r != q) {
                            const T srp = GABO_SYM(S, r, p), srq = GABO_SYM(S, r, q);
                            GABO_SYM(S, r, p) = fma(cs, srp, -sn * srq);
                            GABO_SYM(S, r, q) = fma(sn, srp, cs * srq);
                        }
                    }
                    if (kWantV) {
#pragma unroll
                        for (int r = 0; r < d; ++r) {
                            const T vp = V[r][p], vq = V[r][q];
                            V[r][p] = fma(cs, vp, -sn * vq);
                            V[r][q] = fma(sn, vp, cs * vq);
                        }
                    }
                }
            }
        }
        if (!__any_sync(__activemask(), rotated)) break;
    }
#pragma unroll
    for (int k = 0; k < d; ++k) lam[k] = S[k][k];
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative two-sided Jacobi for ONE symmetric (possibly indefinite) d x d matrix in shared memory, d <= 8:
// lane = 8 j + r works on pair j of the current round-robin round and on row (then column) r.  The rotations of a
// round act on disjoint index pairs, so all column updates (S <- S J, V <- V J) run in parallel, then all row updates
// (S <- J^T S).  ~45 instructions per round instead of the ~3000-instruction unrolled sweep every lane used to run
// redundantly for the same matrix (the whitened search direction of the acquisition solver).
//   Hu : upper triangle, packed row-major (entry (r,c), r <= c, at r d - r (r-1)/2 + c - r);  S : d*d scratch;
//   V  : d*d eigenvectors (row-major, column k = k-th eigenvector);  lam : d eigenvalues.  Call with a full warp.
// ---------------------------------------------------------------------------------------------------------------
template <int d, typename T>
__device__ __forceinline__ void jacobi_symmetric_warp(const T* Hu, T* S, T* V, T* lam, int lane) {
    using Tr = SymJacobiTraits<T>;
    using RR = RoundRobin<d>;
    static_assert(d <= 8, "one row per lane octet");
    __syncwarp();
    T fro = T(0);
    for (int e = lane; e < d * d; e += 32) {
        const int r = e / d, c = e % d;
        const int lo = r < c ? r : c, hi = r < c ? c : r;
        const T v = Hu[lo * d - (lo * (lo - 1)) / 2 + (hi - lo)];
        S[e] = v;
        V[e] = (r == c) ? T(1) : T(0);
        fro = fma(v, v, fro);
    }
    fro = warp_sum(fro);
    const T thr = Tr::tol() * Tr::sqrt_(fro);
    __syncwarp();
    const int j = lane >> 3, r = lane & 7;
    const bool lane_ok = (j < RR::kPairs) && (r < d);
#pragma unroll 1
    for (int sweep = 0; sweep < Tr::kMaxSweeps; ++sweep) {
        bool rotated = false;
#pragma unroll 1
        for (int round = 0; round < RR::kRounds; ++round) {
            int p = 0, q = 1;
            bool act = false;
            T cs = T(1), sn = T(0);
            if (j < RR::kPairs) {
                const int a = RR::player(round, j), b = RR::player(round, RR::m - 1 - j);
                p = a < b ? a : b;
                q = a < b ? b : a;
                if (q < d) {
                    const T apq = S[p * d + q];
                    if (fabs(apq) > thr) {
                        act = true;
                        const T tau = (S[q * d + q] - S[p * d + p]) * Tr::rcp_(T(2) * apq);
                        const T t = copysign(T(1), tau) * Tr::rcp_(fabs(tau) + Tr::sqrt_(fma(tau, tau, T(1))));
                        cs = Tr::rsqrt_(fma(t, t, T(1)));
                        sn = t * cs;
                    }
                }
            }
            rotated = rotated || act;
            __syncwarp();
            if (act && lane_ok) {           // columns p, q of S and V, row r
                const T sp = S[r * d + p], sq = S[r * d + q];
                S[r * d + p] = fma(cs, sp, -sn * sq);
                S[r * d + q] = fma(sn, sp, cs * sq);
                const T vp = V[r * d + p], vq = V[r * d + q];
                V[r * d + p] = fma(cs, vp, -sn * vq);
                V[r * d + q] = fma(sn, vp, cs * vq);
            }
            __syncwarp();
            if (act && lane_ok) {           // rows p, q of S, column r
                const T sp = S[p * d + r], sq = S[q * d + r];
                S[p * d + r] = fma(cs, sp, -sn * sq);
                S[q * d + r] = fma(sn, sp, cs * sq);
            }
            __syncwarp();
        }
        if (!__any_sync(0xffffffffu, rotated)) break;
    }
    for (int k = lane; k < d; k += 32) lam[k] = S[k * d + k];
    __syncwarp();
}

// Full d x d from a packed lower-triangular array (row-major): M[r][c] = tri[r(r+1)/2 + c] for c <= r, else 0.
template <int d, typename T, typename Acc>
__device__ __forceinline__ void tri_expand(Acc tri, T (&M)[d][d]) {
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) M[r][c] = (c <= r) ? static_cast<T>(tri(tri_idx(r, c))) : T(0);
}

// C = A * B (full d x d)
template <int d, typename T>
__device__ __forceinline__ void matmul(const T (&A)[d][d], const T (&B)[d][d], T (&C)[d][d]) {
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < d; ++k) s = fma(A[r][k], B[k][c], s);
            C[r][c] = s;
        }
}
// C = A * B^T
template <int d, typename T>
__device__ __forceinline__ void matmul_nt(const T (&A)[d][d], const T (&B)[d][d], T (&C)[d][d]) {
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < d; ++k) s = fma(A[r][k], B[c][k], s);
            C[r][c] = s;
        }
}
// C = sum_k f[k] * g_k g_k^T for the columns g_k of G  (symmetric; both triangles written)
template <int d, typename T>
__device__ __forceinline__ void weighted_outer(const T (&G)[d][d], const T (&f)[d], T (&C)[d][d]) {
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = r; c < d; ++c) {
            T s = T(0);
#pragma unroll
            for (int k = 0; k < d; ++k) s = fma(f[k] * G[r][k], G[c][k], s);
            C[r][c] = s;
            C[c][r] = s;
        }
}

}  // namespace gabo
