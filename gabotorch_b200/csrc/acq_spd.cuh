// Acquisition value / multi-start Riemannian CG on SPD(d) under the affine-invariant metric (A1 + A4 with M2
// primitives).  One warp per restart; lanes own GP training points, each running the same
// in-register one-sided Jacobi eigen-solve as the Gram kernel (spd_common.cuh).
//
// pymanopt's PositiveDefinite (reference call sites manifold_optimize.py:207-221, gabo_spd.py:98; numpy statements
// Riemannian_utils/spd_utils.py:104-213) gives  inner_X(U,V) = tr(X^-1 U X^-1 V),  retr = exp_X(U) = X expm(X^-1 U),
// transp = identity,  Log_X(Y) = X^1/2 logm(X^-1/2 Y X^-1/2) X^1/2.  The solver is run in WHITENED COORDINATES:
// the iterate is carried as an inverse factor  Finv  with  X = Finv^-1 Finv^-T,  a tangent vector xi as the symmetric
// matrix  Xi = Finv xi Finv^T.  Then (exactly, for any factor):
//     inner_X(xi1, xi2) = <Xi1, Xi2>_F,
//     exp_X(a eta): with the eigen-decomposition  H = V diag(lam) V^T  of the whitened direction, the new inverse
//                   factor is  Finv' = diag(exp(-a lam / 2)) V^T Finv                     (no Cholesky, no expm)
//     identity transport to the new point:  Xi' = E V^T Xi V E,  E = diag(exp(-a lam / 2));  for eta itself
//                   H' = diag(lam exp(-a lam)),
//     Log_X(X_i) whitened = logm(G_i G_i^T) with G_i = Finv L_i (L_i the Cholesky factor of the training point)
//                   = sum_k log(l_k)/l_k g_k g_k^T  after the one-sided Jacobi (columns g_k, l_k = |g_k|^2).
// One symmetric eigen-solve per CG iteration (the direction), one one-sided Jacobi per training point per trial.
// The iterate X is materialised once, at the end.  Same iterates as pymanopt's operations in exact arithmetic.
#pragma once
#include <cstdlib>
#include "acq_common.cuh"
#include "spd_common.cuh"

namespace gabo {
namespace {

__host__ __device__ constexpr int ui(int d, int r, int c) { return r * d - (r * (r - 1)) / 2 + (c - r); }  // r <= c

template <int d, typename T, int NCH>
__global__ void __launch_bounds__(kAcqWarps * 32)
    spd_acq_kernel(GpParams gp, RcgParams opt, int mode, double* __restrict__ x_io, int64_t r,
                   double* __restrict__ value, double* __restrict__ grad_out, int32_t* __restrict__ iters,
                   int32_t* __restrict__ reason, int32_t* __restrict__ flags) {
    constexpr int TRI = tri_size(d);
    constexpr int FS = factor_stride(d);
    constexpr int DD = d * d;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    T* Ls = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * n * TRI));
    T* alpha = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad));
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * n * n));
    double* dbase = reinterpret_cast<double*>(smem_raw + cv.take(sizeof(double) * kAcqWarps * 2 * DD));
    constexpr int kPerWarpT = 3 * DD + 3 * TRI + 2 * d;
    T* tbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kAcqWarps * (kPerWarpT + npad)));

    // entry-major (Ls[e * n + i]): lanes = training points read consecutive words (point-major rows of TRI words
    // collided 4 ways for TRI = 36)
    for (int e = threadIdx.x; e < n * TRI; e += blockDim.x)
        Ls[(e % TRI) * n + (e / TRI)] = static_cast<T>(gp.x_train[static_cast<int64_t>(e / TRI) * FS + (e % TRI)]);
    for (int e = threadIdx.x; e < n; e += blockDim.x) alpha[e] = static_cast<T>(gp.alpha[e]);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Minv[e] = static_cast<T>(gp.minv[e]);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rid = static_cast<int64_t>(blockIdx.x) * kAcqWarps + warp;
    if (rid >= r) return;

    double* Finv = dbase + warp * 2 * DD;  // inverse factor of the iterate (fp64)
    double* Q0 = Finv + DD;                // V^T Finv
    T* wt = tbase + warp * (kPerWarpT + npad);
    T* Qs = wt;                 // Q0 in T (what the lanes read)
    T* Vs = Qs + DD;            // eigenvectors of the whitened direction
    T* tmp = Vs + DD;           // scratch d x d
    T* Om = tmp + DD;           // whitened cost gradient (upper triangle)
    T* Hh = Om + TRI;           // whitened search direction
    T* OmV = Hh + TRI;          // V^T Om V
    T* lamH = OmV + TRI;        // eigenvalues of Hh
    T* Es = lamH + d;           // exp(-a lam / 2) of the trial
    T* ksh = Es + d;

    const T s_out = static_cast<T>(gp.outputscale), beta = static_cast<T>(gp.beta);
    T k_l[NCH], mk_l[NCH];
    T W[NCH][TRI];
    EiScalars<T> sc;

    // ---- cost at the trial point with inverse factor diag(Es) * Qs ------------------------------------------
    auto cost_trial = [&]() -> T {
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T kk = T(0);
            if (i < n) {
                T G[d][d];
#pragma unroll
                for (int rr = 0; rr < d; ++rr) {
                    const T er = Es[rr];
#pragma unroll
                    for (int c = 0; c < d; ++c) {
                        T s = T(0);
#pragma unroll
                        for (int m = c; m < d; ++m) s = fma(Qs[rr * d + m], Ls[tri_idx(m, c) * n + i], s);
                        G[rr][c] = er * s;
                    }
                }
                T lam[d];
                jacobi_onesided_compact<d, T>(G, lam);
                T dsq = T(1e-15);  // spd_utils_torch.py:120
                T f[d];
#pragma unroll
                for (int k = 0; k < d; ++k) {
                    const T l = M<T>::log_(lam[k]);
                    dsq = fma(l, l, dsq);
                    f[k] = l / lam[k];
                }
                kk = s_out * M<T>::exp_(-beta * dsq);
#pragma unroll
                for (int rr = 0; rr < d; ++rr)
#pragma unroll
                    for (int c = rr; c < d; ++c) {
                        T s = T(0);
#pragma unroll
                        for (int k = 0; k < d; ++k) s = fma(f[k] * G[rr][k], G[c][k], s);
                        W[ch][ui(d, rr, c)] = s;
                    }
                ksh[i] = kk;
            } else {
#pragma unroll
                for (int e = 0; e < TRI; ++e) W[ch][e] = T(0);
            }
            k_l[ch] = kk;
        }
        __syncwarp();
        T ka = T(0), kmk = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T mk = T(0);
            if (i < n) {
                T m0 = T(0), m1 = T(0), m2 = T(0), m3 = T(0);   // independent accumulators (latency)
                int j = 0;
                for (; j + 3 < n; j += 4) {
                    m0 = fma(Minv[j * n + i], ksh[j], m0);
                    m1 = fma(Minv[(j + 1) * n + i], ksh[j + 1], m1);
                    m2 = fma(Minv[(j + 2) * n + i], ksh[j + 2], m2);
                    m3 = fma(Minv[(j + 3) * n + i], ksh[j + 3], m3);
                }
                for (; j < n; ++j) m0 = fma(Minv[j * n + i], ksh[j], m0);
                mk = (m0 + m1) + (m2 + m3);
                ka = fma(k_l[ch], alpha[i], ka);
                kmk = fma(k_l[ch], mk, kmk);
            }
            mk_l[ch] = mk;
        }
        ka = warp_sum(ka);
        kmk = warp_sum(kmk);
        sc = ei_scalars<T>(ka, kmk, gp);
        const T cst = -sc.ei;
        return (cst == cst) ? cst : M<T>::inf();
    };

    // ---- whitened cost gradient at the last evaluated trial point -> dst (upper triangle, smem) --------------
    auto assemble_grad = [&](T* dst) {
        T p[TRI];
#pragma unroll
        for (int e = 0; e < TRI; ++e) p[e] = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            if (i < n) {
                const T w = -sc.cdf * alpha[i] - sc.pdf_over_sigma * mk_l[ch];
                const T coef = T(2) * beta * w * k_l[ch];
#pragma unroll
                for (int e = 0; e < TRI; ++e) p[e] = fma(coef, W[ch][e], p[e]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < TRI; ++e) {
            const T s = warp_sum(p[e]);
            if (lane == (e & 31)) dst[e] = -s;  // cost = -EI
        }
        __syncwarp();
    };

    auto sym_inner = [&](const T* a, const T* b) -> T {  // <A, B>_F of two symmetric matrices (upper storage)
        T s = T(0);
        for (int e = lane; e < TRI; e += 32) {
            // entry e is on the diagonal iff e == ui(r, r) for some r
            bool diag = false;
#pragma unroll
            for (int rr = 0; rr < d; ++rr) diag = diag || (e == ui(d, rr, rr));
            s = fma((diag ? T(1) : T(2)) * a[e], b[e], s);
        }
        return warp_sum(s);
    };

    // ---- initial point: Cholesky of X0, Finv = L^-1 -----------------------------------------------------------
    const double* xin = x_io + rid * DD;
    double L0[TRI], A0[TRI];
    const bool ok0 = chol_inv<d>([&](int rr, int c) { return xin[rr * d + c]; }, L0, A0);
    if (!ok0) {
        if (flags && lane == 0) atomicOr(flags, 1);
        if (lane == 0) {
            value[rid] = M<double>::nan();
            if (iters) iters[rid] = 0;
            if (reason) reason[rid] = -1;
        }
        if (mode == 0 && grad_out)
            for (int e = lane; e < DD; e += 32) grad_out[rid * DD + e] = M<double>::nan();
        return;
    }
#pragma unroll
    for (int rr = 0; rr < d; ++rr)
#pragma unroll
        for (int c = 0; c < d; ++c) {
            const int e = rr * d + c;
            if (lane == (e & 31)) {
                const double v = (c <= rr) ? A0[tri_idx(rr, c)] : 0.0;
                Finv[e] = v;
                Qs[e] = static_cast<T>(v);
                tmp[e] = static_cast<T>((c <= rr) ? L0[tri_idx(rr, c)] : 0.0);  // lower factor, for the ambient gradient
            }
        }
    for (int k = lane; k < d; k += 32) Es[k] = T(1);
    __syncwarp();

    // The whole solve is ONE loop with ONE cost_trial() call site: the initial evaluation is its first pass.  cost_trial
    // inlines a fully unrolled d x d Jacobi (thousands of instructions for d = 8); with three call sites the kernel was
    // 27k instructions (440 KB) and ncu showed the warps starved on instruction fetch (stall no_instruction 5.7 per
    // issue, 13 % issue-active).
    T cost = T(0), gPg = T(0), gradnorm = T(0);
    int it = 0, why = 0;
    T stepsize = M<T>::nan();
    T oldalpha = T(-1);
    const T mingrad = static_cast<T>(opt.mingradnorm), minstep = static_cast<T>(opt.minstepsize);
    const T contraction = static_cast<T>(opt.contraction), suff = static_cast<T>(opt.suff_decr);

    auto set_trial = [&](T a) {
        __syncwarp();
        for (int k = lane; k < d; k += 32) Es[k] = M<T>::exp_(T(-0.5) * a * lamH[k]);
        __syncwarp();
    };

    bool first = true;
    while (true) {
        T df0 = T(0), norm_d = T(0), a = T(0);
        if (!first) {
            if (it + 1 >= opt.maxiter) { why = 1; break; }
            if (gradnorm < mingrad) { why = 2; break; }
            if (stepsize < minstep) { why = 3; break; }
            df0 = sym_inner(Om, Hh);
            if (df0 >= T(0)) {
                __syncwarp();
                for (int e = lane; e < TRI; e += 32) Hh[e] = -Om[e];
                __syncwarp();
                df0 = -gPg;
            }
            norm_d = M<T>::sqrt_(sym_inner(Hh, Hh));

            // eigen-decomposition of the whitened direction (every lane, redundantly; warp-uniform result)
            {
                T S[d][d], lam[d], V[d][d];
#pragma unroll
                for (int rr = 0; rr < d; ++rr)
#pragma unroll
                    for (int c = 0; c < d; ++c) S[rr][c] = (c >= rr) ? Hh[ui(d, rr, c)] : T(0);
                jacobi_symmetric<d, T, true>(S, lam, V);
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < d; ++rr)
#pragma unroll
                    for (int c = 0; c < d; ++c) {
                        const int e = rr * d + c;
                        if (lane == (e & 31)) Vs[e] = V[rr][c];
                    }
#pragma unroll
                for (int k = 0; k < d; ++k)
                    if (lane == k) lamH[k] = lam[k];
                __syncwarp();
            }
            // Q0 = V^T Finv (fp64), Qs = (T) Q0;  tmp = Om V;  OmV = V^T tmp
            for (int e = lane; e < DD; e += 32) {
                const int rr = e / d, c = e % d;
                double s = 0.0;
                T t = T(0);
                for (int m = 0; m < d; ++m) {
                    s = fma(static_cast<double>(Vs[m * d + rr]), Finv[m * d + c], s);
                    const int lo = rr < m ? rr : m, hi = rr < m ? m : rr;
                    t = fma(Om[ui(d, lo, hi)], Vs[m * d + c], t);
                }
                Q0[e] = s;
                Qs[e] = static_cast<T>(s);
                tmp[e] = t;
            }
            __syncwarp();
            for (int e = lane; e < DD; e += 32) {
                const int rr = e / d, c = e % d;
                if (c >= rr) {
                    T t = T(0);
                    for (int m = 0; m < d; ++m) t = fma(Vs[m * d + rr], tmp[m * d + c], t);
                    OmV[ui(d, rr, c)] = t;
                }
            }
            __syncwarp();

            a = (oldalpha >= T(0)) ? oldalpha : static_cast<T>(opt.initial_stepsize) / norm_d;
        }
        // LineSearchAdaptive: first trial unconditional, then backtrack while the Armijo test fails (<= ls_maxiter times)
        T newf;
        int evals = 0;
        do {
            if (!first) {
                if (evals > 0) a *= contraction;
                set_trial(a);
            }
            newf = cost_trial();
            ++evals;
        } while (!first && newf > cost + suff * a * df0 && evals <= opt.ls_maxiter);

        if (first) {   // this pass was the evaluation at the starting point
            first = false;
            cost = newf;
            {   // this kernel only serves gabo_ei_eval now (mode 0); the solver is spd_rcg_cta_kernel
                if (lane == 0) value[rid] = static_cast<double>(-cost);
                if (grad_out) {
                    assemble_grad(Om);
                    // ambient Riemannian gradient of EI: L (-Om) L^T
                    for (int e = lane; e < DD; e += 32) {
                        const int rr = e / d, c = e % d;
                        double s = 0.0;
                        for (int a = 0; a <= rr; ++a)
                            for (int b = 0; b <= c; ++b) {
                                const int lo = a < b ? a : b, hi = a < b ? b : a;
                                s += static_cast<double>(tmp[rr * d + a]) * static_cast<double>(Om[ui(d, lo, hi)]) *
                                     static_cast<double>(tmp[c * d + b]);
                            }
                        grad_out[rid * DD + e] = -s;
                    }
                }
                return;
            }
            assemble_grad(Om);
            gPg = sym_inner(Om, Om);
            gradnorm = M<T>::sqrt_(gPg);
            for (int e = lane; e < TRI; e += 32) Hh[e] = -Om[e];
            __syncwarp();
            continue;
        }

        const bool stay = newf > cost;
        if (stay) {
            a = T(0);
            set_trial(a);
            newf = cost;
        }
        stepsize = a * norm_d;
        oldalpha = (evals == 2) ? a : T(2) * a;

        // new gradient (in the coordinates of the accepted point) into tmp[0..TRI)
        if (stay) {
            for (int e = lane; e < TRI; e += 32) tmp[e] = OmV[e];
            __syncwarp();
        } else {
            assemble_grad(tmp);
        }
        // transported old gradient E OmV E and direction diag(lam exp(-a lam)); Hestenes-Stiefel
        T ip = T(0), den = T(0), ngg = T(0);
        for (int e = lane; e < TRI; e += 32) {
            int rr = 0, c = 0;
#pragma unroll
            for (int q = 0; q < d; ++q)
                if (e >= ui(d, q, q)) {
                    rr = q;
                    c = q + (e - ui(d, q, q));
                }
            const T og = Es[rr] * OmV[e] * Es[c];
            const T oe = (rr == c) ? lamH[rr] * Es[rr] * Es[rr] : T(0);
            const T gnew = tmp[e];
            const T df = gnew - og;
            const T wgt = (rr == c) ? T(1) : T(2);
            ip = fma(wgt * gnew, df, ip);
            den = fma(wgt * df, oe, den);
            ngg = fma(wgt * gnew, gnew, ngg);
            Hh[e] = oe;
        }
        ip = warp_sum(ip);
        den = warp_sum(den);
        ngg = warp_sum(ngg);
        const T q = ip / den;
        const T bcg = (q > T(0) && q < M<T>::inf()) ? q : T(0);  // max(0, q); NaN -> 0 like Python's max(0, nan)
        __syncwarp();
        for (int e = lane; e < TRI; e += 32) {
            const T gnew = tmp[e];
            Hh[e] = fma(bcg, Hh[e], -gnew);
            Om[e] = gnew;
        }
        // accept: Finv <- E Q0
        for (int e = lane; e < DD; e += 32)
            Finv[e] = exp(-0.5 * static_cast<double>(a) * static_cast<double>(lamH[e / d])) * Q0[e];
        __syncwarp();
        cost = newf;
        gPg = ngg;
        gradnorm = M<T>::sqrt_(ngg);
        ++it;
    }

    // materialise X = Finv^-1 Finv^-T = (Finv^T Finv)^-1: Y = Finv^T Finv, Y = Ly Ly^T, X = Ly^-T Ly^-1
    {
        double Ly[TRI], Ay[TRI];
        chol_inv<d>(
            [&](int rr, int c) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < d; ++m) s = fma(Finv[m * d + rr], Finv[m * d + c], s);
                return s;
            },
            Ly, Ay);
#pragma unroll
        for (int rr = 0; rr < d; ++rr)
#pragma unroll
            for (int c = rr; c < d; ++c) {
                double s = 0.0;
#pragma unroll
                for (int m = c; m < d; ++m) s = fma(Ay[tri_idx(m, rr)], Ay[tri_idx(m, c)], s);
                const int e = rr * d + c;
                if (lane == (e & 31)) {
                    x_io[rid * DD + rr * d + c] = s;
                    x_io[rid * DD + c * d + rr] = s;
                }
            }
    }
    if (lane == 0) {
        value[rid] = static_cast<double>(-cost);
        if (iters) iters[rid] = it;
        if (reason) reason[rid] = why;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A1 on SPD(d): ONE CTA PER RESTART, speculative backtracking across its kSpec warps (same idea as
// sphere_rcg_cta_kernel).  The solver state (inverse factor, whitened gradient / direction, eigenvectors of the
// direction) lives in shared memory and is written by warp 0 only; every warp reads it, carries the scalar CG state in
// registers and takes identical decisions.  Warp w evaluates trial step alpha c^(base + w): its lanes run the
// per-training-point one-sided Jacobi (distance + log map) for THAT trial, so a cost evaluation -- by far the most
// expensive piece, ~25k instructions for d = 8 -- is paid once per kSpec trials instead of once per trial.  The first
// trial that passes the Armijo test wins and its warp assembles the gradient; accepted steps, evaluation counts and
// iterates are those of the sequential search.
// ---------------------------------------------------------------------------------------------------------------
template <int d, typename T, int NCH, int kSpec>
__global__ void __launch_bounds__(kSpec * 32)
    spd_rcg_cta_kernel(GpParams gp, RcgParams opt, double* __restrict__ x_io, int64_t r, double* __restrict__ value,
                       int32_t* __restrict__ iters, int32_t* __restrict__ reason, int32_t* __restrict__ flags) {
    constexpr int TRI = tri_size(d);
    constexpr int FS = factor_stride(d);
    constexpr int DD = d * d;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    T* Ls = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * n * TRI));
    T* alpha = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad));
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * n * n));
    double* Finv = reinterpret_cast<double*>(smem_raw + cv.take(sizeof(double) * 2 * DD));  // inverse factor (fp64)
    double* Q0 = Finv + DD;                                                                  // V^T Finv
    constexpr int kSharedT = 3 * DD + 4 * TRI + d;
    T* Qs = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kSharedT));  // Q0 in T (what the lanes read)
    T* Vs = Qs + DD;            // eigenvectors of the whitened direction
    T* tmp = Vs + DD;           // scratch d x d
    T* Om = tmp + DD;           // whitened cost gradient (upper triangle)
    T* Hh = Om + TRI;           // whitened search direction
    T* OmV = Hh + TRI;          // V^T Om V
    T* gnew_s = OmV + TRI;      // new gradient
    T* lamH = gnew_s + TRI;     // eigenvalues of Hh
    T* f_sh = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * 2 * kSpec));              // trial costs, double-buffered
    T* wbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kSpec * (d + npad)));   // per warp: Es | ksh

    for (int e = threadIdx.x; e < n * TRI; e += blockDim.x)
        Ls[(e % TRI) * n + (e / TRI)] = static_cast<T>(gp.x_train[static_cast<int64_t>(e / TRI) * FS + (e % TRI)]);
    for (int e = threadIdx.x; e < n; e += blockDim.x) alpha[e] = static_cast<T>(gp.alpha[e]);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Minv[e] = static_cast<T>(gp.minv[e]);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* Es = wbase + warp * (d + npad);   // exp(-a lam / 2) of MY trial
    T* ksh = Es + d;
    const T s_out = static_cast<T>(gp.outputscale), beta = static_cast<T>(gp.beta);
    const T mingrad = static_cast<T>(opt.mingradnorm), minstep = static_cast<T>(opt.minstepsize);
    const T contraction = static_cast<T>(opt.contraction), suff = static_cast<T>(opt.suff_decr);

    T k_l[NCH], mk_l[NCH];
    T W[NCH][TRI];
    EiScalars<T> sc;

    // ---- cost at the trial point with inverse factor diag(Es) * Qs (Es: this warp's trial) ---------------------
    auto cost_trial = [&]() -> T {
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T kk = T(0);
            if (i < n) {
                T G[d][d];
#pragma unroll
                for (int rr = 0; rr < d; ++rr) {
                    const T er = Es[rr];
#pragma unroll
                    for (int c = 0; c < d; ++c) {
                        T s = T(0);
#pragma unroll
                        for (int m = c; m < d; ++m) s = fma(Qs[rr * d + m], Ls[tri_idx(m, c) * n + i], s);
                        G[rr][c] = er * s;
                    }
                }
                T lam[d];
                jacobi_onesided_compact<d, T>(G, lam);
                T dsq = T(1e-15);  // spd_utils_torch.py:120
                T f[d];
#pragma unroll
                for (int k = 0; k < d; ++k) {
                    const T l = M<T>::log_(lam[k]);
                    dsq = fma(l, l, dsq);
                    f[k] = l / lam[k];
                }
                kk = s_out * M<T>::exp_(-beta * dsq);
#pragma unroll
                for (int rr = 0; rr < d; ++rr)
#pragma unroll
                    for (int c = rr; c < d; ++c) {
                        T s = T(0);
#pragma unroll
                        for (int k = 0; k < d; ++k) s = fma(f[k] * G[rr][k], G[c][k], s);
                        W[ch][ui(d, rr, c)] = s;
                    }
                ksh[i] = kk;
            } else {
#pragma unroll
                for (int e = 0; e < TRI; ++e) W[ch][e] = T(0);
            }
            k_l[ch] = kk;
        }
        __syncwarp();
        T ka = T(0), kmk = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T mk = T(0);
            if (i < n) {
                T m0 = T(0), m1 = T(0), m2 = T(0), m3 = T(0);
                int j = 0;
                for (; j + 3 < n; j += 4) {
                    m0 = fma(Minv[j * n + i], ksh[j], m0);
                    m1 = fma(Minv[(j + 1) * n + i], ksh[j + 1], m1);
                    m2 = fma(Minv[(j + 2) * n + i], ksh[j + 2], m2);
                    m3 = fma(Minv[(j + 3) * n + i], ksh[j + 3], m3);
                }
                for (; j < n; ++j) m0 = fma(Minv[j * n + i], ksh[j], m0);
                mk = (m0 + m1) + (m2 + m3);
                ka = fma(k_l[ch], alpha[i], ka);
                kmk = fma(k_l[ch], mk, kmk);
            }
            mk_l[ch] = mk;
        }
        ka = warp_sum(ka);
        kmk = warp_sum(kmk);
        sc = ei_scalars<T>(ka, kmk, gp);
        const T cst = -sc.ei;
        return (cst == cst) ? cst : M<T>::inf();
    };

    // ---- whitened cost gradient at MY last trial point -> dst (upper triangle, shared); one warp only ----------
    auto assemble_grad = [&](T* dst) {
        T p[TRI];
#pragma unroll
        for (int e = 0; e < TRI; ++e) p[e] = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            if (i < n) {
                const T w = -sc.cdf * alpha[i] - sc.pdf_over_sigma * mk_l[ch];
                const T coef = T(2) * beta * w * k_l[ch];
#pragma unroll
                for (int e = 0; e < TRI; ++e) p[e] = fma(coef, W[ch][e], p[e]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < TRI; ++e) {
            const T s = warp_sum(p[e]);
            if (lane == (e & 31)) dst[e] = -s;  // cost = -EI
        }
        __syncwarp();
    };

    auto sym_inner = [&](const T* a, const T* b) -> T {  // <A, B>_F of two symmetric matrices (upper storage)
        T s = T(0);
        for (int e = lane; e < TRI; e += 32) {
            bool diag = false;
#pragma unroll
            for (int rr = 0; rr < d; ++rr) diag = diag || (e == ui(d, rr, rr));
            s = fma((diag ? T(1) : T(2)) * a[e], b[e], s);
        }
        return warp_sum(s);
    };
    auto set_trial = [&](T a) {   // my warp's trial scaling
        __syncwarp();
        for (int k = lane; k < d; k += 32) Es[k] = M<T>::exp_(T(-0.5) * a * lamH[k]);
        __syncwarp();
    };

    for (int64_t rid = blockIdx.x; rid < r; rid += gridDim.x) {
        __syncthreads();   // the previous restart of this CTA is completely finished
        // ---- initial point: Cholesky of X0, Finv = L^-1 (every thread, redundantly; warp 0 writes) -------------
        const double* xin = x_io + rid * DD;
        double L0[TRI], A0[TRI];
        const bool ok0 = chol_inv<d>([&](int rr, int c) { return xin[rr * d + c]; }, L0, A0);
        if (!ok0) {   // uniform across the CTA
            if (threadIdx.x == 0) {
                if (flags) atomicOr(flags, 1);
                value[rid] = M<double>::nan();
                if (iters) iters[rid] = 0;
                if (reason) reason[rid] = -1;
            }
            continue;
        }
        if (warp == 0) {
#pragma unroll
            for (int rr = 0; rr < d; ++rr)
#pragma unroll
                for (int c = 0; c < d; ++c) {
                    const int e = rr * d + c;
                    if (lane == (e & 31)) {
                        const double v = (c <= rr) ? A0[tri_idx(rr, c)] : 0.0;
                        Finv[e] = v;
                        Qs[e] = static_cast<T>(v);
                    }
                }
        }
        for (int k = lane; k < d; k += 32) Es[k] = T(1);
        __syncthreads();

        T cost = T(0), gPg = T(0), gradnorm = T(0);
        int it = 0, why = 0;
        unsigned fbuf = 0;
        T stepsize = M<T>::nan();
        T oldalpha = T(-1);
        T astep = T(0);        // accepted step of the last iteration
        bool first = true;
        while (true) {
            T df0 = T(0), norm_d = T(0), a = T(0);
            if (!first) {
                if (it + 1 >= opt.maxiter) { why = 1; break; }
                if (gradnorm < mingrad) { why = 2; break; }
                if (stepsize < minstep) { why = 3; break; }
                df0 = sym_inner(Om, Hh);
                if (df0 >= T(0)) {   // not a descent direction: restart from steepest descent
                    __syncthreads();
                    if (warp == 0)
                        for (int e = lane; e < TRI; e += 32) Hh[e] = -Om[e];
                    __syncthreads();
                    df0 = -gPg;
                }
                norm_d = M<T>::sqrt_(sym_inner(Hh, Hh));

                // eigen-decomposition of the whitened direction.  d >= 5: warp 0, cooperatively in shared memory (tmp =
                // scratch); small d: every lane redundantly in registers (cheaper than the round trips; measured at d = 3)
                if constexpr (d >= 5) {
                    __syncthreads();
                    if (warp == 0) jacobi_symmetric_warp<d, T>(Hh, tmp, Vs, lamH, lane);
                    __syncthreads();
                } else {
                    T S[d][d], lam[d], V[d][d];
#pragma unroll
                    for (int rr = 0; rr < d; ++rr)
#pragma unroll
                        for (int c = 0; c < d; ++c) S[rr][c] = (c >= rr) ? Hh[ui(d, rr, c)] : T(0);
                    jacobi_symmetric<d, T, true>(S, lam, V);
                    if (warp == 0) {
#pragma unroll
                        for (int rr = 0; rr < d; ++rr)
#pragma unroll
                            for (int c = 0; c < d; ++c) {
                                const int e = rr * d + c;
                                if (lane == (e & 31)) Vs[e] = V[rr][c];
                            }
#pragma unroll
                        for (int k = 0; k < d; ++k)
                            if (lane == k) lamH[k] = lam[k];
                    }
                    __syncthreads();
                }
                // Q0 = V^T Finv (fp64), Qs = (T) Q0;  tmp = Om V;  OmV = V^T tmp
                if (warp == 0) {
                    for (int e = lane; e < DD; e += 32) {
                        const int rr = e / d, c = e % d;
                        double s = 0.0;
                        T t = T(0);
                        for (int m = 0; m < d; ++m) {
                            s = fma(static_cast<double>(Vs[m * d + rr]), Finv[m * d + c], s);
                            const int lo = rr < m ? rr : m, hi = rr < m ? m : rr;
                            t = fma(Om[ui(d, lo, hi)], Vs[m * d + c], t);
                        }
                        Q0[e] = s;
                        Qs[e] = static_cast<T>(s);
                        tmp[e] = t;
                    }
                    __syncwarp();
                    for (int e = lane; e < DD; e += 32) {
                        const int rr = e / d, c = e % d;
                        if (c >= rr) {
                            T t = T(0);
                            for (int m = 0; m < d; ++m) t = fma(Vs[m * d + rr], tmp[m * d + c], t);
                            OmV[ui(d, rr, c)] = t;
                        }
                    }
                }
                __syncthreads();
                a = (oldalpha >= T(0)) ? oldalpha : static_cast<T>(opt.initial_stepsize) / norm_d;
            }

            // LineSearchAdaptive: trial k = 0 .. ls_maxiter uses alpha_0 c^k; first Armijo pass wins, the last trial is
            // kept when none passes.  Warp w evaluates trial base + w.
            // (The evaluation at the starting point goes through the same code as a one-trial search evaluated by every
            // warp: cost_trial() must keep a single call site, see spd_acq_kernel.)
            T newf = T(0);
            int evals = 0, sel = 0;
            {
                const int ktotal = first ? 1 : opt.ls_maxiter + 1;
                bool done = false;
                for (int base = 0; base < ktotal && !done; base += kSpec) {
                    T ak[kSpec];
#pragma unroll
                    for (int t = 0; t < kSpec; ++t) {
                        ak[t] = a;
                        a *= contraction;
                    }
                    T amine = ak[0];
#pragma unroll
                    for (int t = 1; t < kSpec; ++t) amine = (warp == t) ? ak[t] : amine;
                    T fmine = M<T>::inf();
                    if (first || base + warp < ktotal) {
                        if (!first) set_trial(amine);
                        fmine = cost_trial();
                    }
                    T* fb = f_sh + fbuf * kSpec;
                    fbuf ^= 1u;
                    if (lane == 0) fb[warp] = fmine;
                    __syncthreads();
#pragma unroll
                    for (int t = 0; t < kSpec; ++t) {
                        if (!done && base + t < ktotal) {
                            const T ft = fb[t];
                            evals = base + t + 1;
                            sel = t;
                            newf = ft;
                            a = ak[t];
                            done = first || !(ft > cost + suff * ak[t] * df0);
                        }
                    }
                    if (!done && base + kSpec < ktotal) a = ak[kSpec - 1] * contraction;
                }
            }

            if (first) {   // this pass was the evaluation at the starting point
                first = false;
                cost = newf;
                if (warp == 0) assemble_grad(Om);
                __syncthreads();
                gPg = sym_inner(Om, Om);
                gradnorm = M<T>::sqrt_(gPg);
                __syncthreads();
                if (warp == 0)
                    for (int e = lane; e < TRI; e += 32) Hh[e] = -Om[e];
                __syncthreads();
                continue;
            }

            const bool stay = newf > cost;
            if (stay) {
                a = T(0);
                newf = cost;
            }
            stepsize = a * norm_d;
            oldalpha = (evals == 2) ? a : T(2) * a;
            astep = a;

            // new gradient (in the coordinates of the accepted point) into gnew_s: the winning warp holds its log maps
            if (stay) {
                if (warp == 0)
                    for (int e = lane; e < TRI; e += 32) gnew_s[e] = OmV[e];
            } else if (warp == sel) {
                assemble_grad(gnew_s);
            }
            set_trial(a);          // every warp: Es of the accepted step
            __syncthreads();
            // transported old gradient E OmV E and direction diag(lam exp(-a lam)); Hestenes-Stiefel
            T ip = T(0), den = T(0), ngg = T(0);
            for (int e = lane; e < TRI; e += 32) {
                int rr = 0, c = 0;
#pragma unroll
                for (int q = 0; q < d; ++q)
                    if (e >= ui(d, q, q)) {
                        rr = q;
                        c = q + (e - ui(d, q, q));
                    }
                const T og = Es[rr] * OmV[e] * Es[c];
                const T oe = (rr == c) ? lamH[rr] * Es[rr] * Es[rr] : T(0);
                const T gnew = gnew_s[e];
                const T df = gnew - og;
                const T wgt = (rr == c) ? T(1) : T(2);
                ip = fma(wgt * gnew, df, ip);
                den = fma(wgt * df, oe, den);
                ngg = fma(wgt * gnew, gnew, ngg);
            }
            ip = warp_sum(ip);
            den = warp_sum(den);
            ngg = warp_sum(ngg);
            const T q = ip / den;
            const T bcg = (q > T(0) && q < M<T>::inf()) ? q : T(0);  // max(0, q); NaN -> 0 like Python's max(0, nan)
            __syncthreads();       // everyone has read Hh / Om / OmV of this iteration
            if (warp == 0) {
                for (int e = lane; e < TRI; e += 32) {
                    int rr = 0, c = 0;
#pragma unroll
                    for (int qq = 0; qq < d; ++qq)
                        if (e >= ui(d, qq, qq)) {
                            rr = qq;
                            c = qq + (e - ui(d, qq, qq));
                        }
                    const T oe = (rr == c) ? lamH[rr] * Es[rr] * Es[rr] : T(0);
                    const T gnew = gnew_s[e];
                    Hh[e] = fma(bcg, oe, -gnew);
                    Om[e] = gnew;
                }
                // accept: Finv <- E Q0
                for (int e = lane; e < DD; e += 32)
                    Finv[e] = exp(-0.5 * static_cast<double>(astep) * static_cast<double>(lamH[e / d])) * Q0[e];
            }
            __syncthreads();
            cost = newf;
            gPg = ngg;
            gradnorm = M<T>::sqrt_(ngg);
            ++it;
        }

        // materialise X = Finv^-1 Finv^-T = (Finv^T Finv)^-1: Y = Finv^T Finv, Y = Ly Ly^T, X = Ly^-T Ly^-1
        if (warp == 0) {
            double Ly[TRI], Ay[TRI];
            chol_inv<d>(
                [&](int rr, int c) {
                    double s = 0.0;
#pragma unroll
                    for (int m = 0; m < d; ++m) s = fma(Finv[m * d + rr], Finv[m * d + c], s);
                    return s;
                },
                Ly, Ay);
#pragma unroll
            for (int rr = 0; rr < d; ++rr)
#pragma unroll
                for (int c = rr; c < d; ++c) {
                    double s = 0.0;
#pragma unroll
                    for (int m = c; m < d; ++m) s = fma(Ay[tri_idx(m, rr)], Ay[tri_idx(m, c)], s);
                    const int e = rr * d + c;
                    if (lane == (e & 31)) {
                        x_io[rid * DD + rr * d + c] = s;
                        x_io[rid * DD + c * d + rr] = s;
                    }
                }
            if (lane == 0) {
                value[rid] = static_cast<double>(-cost);
                if (iters) iters[rid] = it;
                if (reason) reason[rid] = why;
            }
        }
    }
}

template <int d, typename T, int NCH, int kSpec>
int launch_spd_rcg(const GpParams& gp, const RcgParams& opt, double* x, int64_t r, double* value, int32_t* iters,
                   int32_t* reason, cudaStream_t stream, bool force) {   // returns 1 when the width does not fit
    constexpr int TRI = tri_size(d);
    constexpr int DD = d * d;
    const int n = gp.n;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    cv.take(sizeof(T) * n * TRI);
    cv.take(sizeof(T) * npad);
    cv.take(sizeof(T) * n * n);
    cv.take(sizeof(double) * 2 * DD);
    cv.take(sizeof(T) * (3 * DD + 4 * TRI + d));
    cv.take(sizeof(T) * 2 * kSpec);
    cv.take(sizeof(T) * kSpec * (d + npad));
    const size_t smem = cv.off;
    GABO_REQUIRE(smem <= 227 * 1024, GABO_E_UNSUPPORTED,
                 "spd acquisition kernel: n_train=%d, d=%d need %zu bytes of shared memory (> 227 KB)", n, d, smem);
    auto kern = spd_rcg_cta_kernel<d, T, NCH, kSpec>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSpec * 32, smem);
    if (occ < 1) occ = 1;
    const int64_t slots = static_cast<int64_t>(sm_count()) * occ;
    // speculation only pays when every restart has its CTA resident at once; otherwise fall back to a narrower width
    if (kSpec > 1 && !force && r > slots) return 1;
    const unsigned grid = static_cast<unsigned>(imin(r, slots));
    kern<<<grid, kSpec * 32, smem, stream>>>(gp, opt, x, r, value, iters, reason, nullptr);
    return check_launch("spd_rcg_cta_kernel");
}

template <int d, typename T, int NCH>
int launch_spd_t(const GpParams& gp, const RcgParams& opt, int mode, double* x, int64_t r, double* value, double* grad,
                 int32_t* iters, int32_t* reason, cudaStream_t stream) {
    constexpr int TRI = tri_size(d);
    constexpr int DD = d * d;
    const int n = gp.n;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    cv.take(sizeof(T) * n * TRI);
    cv.take(sizeof(T) * npad);
    cv.take(sizeof(T) * n * n);
    cv.take(sizeof(double) * kAcqWarps * 2 * DD);
    cv.take(sizeof(T) * kAcqWarps * (3 * DD + 3 * TRI + 2 * d + npad));
    const size_t smem = cv.off;
    GABO_REQUIRE(smem <= 227 * 1024, GABO_E_UNSUPPORTED,
                 "spd acquisition kernel: n_train=%d, d=%d need %zu bytes of shared memory (> 227 KB)", n, d, smem);
    if (mode == 1) {
        // speculation width from the restart count: idle schedulers (592 on the chip) take speculative trial steps
        const char* e = getenv("GABO_ACQ_SPEC");   // developer switch: force a width
        const int forced = e ? atoi(e) : 0;
        // widths: 4 only for d <= 5 (measured: SPD(3) 512 restarts 0.75 ms with 4 warps per restart vs 1.15 ms with 2;
        // at d = 8 four warps do not fit one wave and gained < 4 %)
        if constexpr (d <= 5) {
            if (forced == 4) return launch_spd_rcg<d, T, NCH, 4>(gp, opt, x, r, value, iters, reason, stream, true);
        }
        if (forced == 2) return launch_spd_rcg<d, T, NCH, 2>(gp, opt, x, r, value, iters, reason, stream, true);
        if (forced != 1) {
            if constexpr (d <= 5) {
                const int rc4 = launch_spd_rcg<d, T, NCH, 4>(gp, opt, x, r, value, iters, reason, stream, false);
                if (rc4 != 1) return rc4;
            }
            const int rc = launch_spd_rcg<d, T, NCH, 2>(gp, opt, x, r, value, iters, reason, stream, false);
            if (rc != 1) return rc;
        }
        return launch_spd_rcg<d, T, NCH, 1>(gp, opt, x, r, value, iters, reason, stream, true);
    }
    auto kern = spd_acq_kernel<d, T, NCH>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const unsigned grid = static_cast<unsigned>((r + kAcqWarps - 1) / kAcqWarps);
    kern<<<grid, kAcqWarps * 32, smem, stream>>>(gp, opt, mode, x, r, value, grad, iters, reason, nullptr);
    return check_launch("spd_acq_kernel");
}

}  // namespace

template <int d>
int launch_acq_spd(const gabo_gp_desc* g, double* x, int64_t r, const gabo_rcg_opts* o, double* value, double* grad,
                   int32_t* iters, int32_t* reason, cudaStream_t stream) {
    GpParams gp{g->n_train, g->dim, g->mean, g->outputscale, g->beta, g->best_f, g->kxx, g->x_train, g->alpha, g->minv};
    RcgParams opt{};
    int mode = 0;
    if (o) {
        mode = 1;
        opt = RcgParams{o->maxiter, o->ls_maxiter, o->mingradnorm, o->minstepsize, o->contraction, o->suff_decr,
                        o->initial_stepsize};
    }
    const bool f64 = g->compute == GABO_F64;
    if (gp.n <= 32) {
        return f64 ? launch_spd_t<d, double, 1>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
                   : launch_spd_t<d, float, 1>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
    }
    return f64 ? launch_spd_t<d, double, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
               : launch_spd_t<d, float, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
}

}  // namespace gabo
