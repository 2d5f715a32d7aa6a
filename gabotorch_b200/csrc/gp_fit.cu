// GP marginal log-likelihood, its gradient and the posterior factors, one CTA per hyper-parameter set.
//
// Replaces, for the models of the reference (botorch SingleTaskGP = constant mean + ScaleKernel(geodesic kernel) +
// GaussianLikelihood; examples/bo_sphere/benchmark_examples/gabo_sphere.py:131-165), the arithmetic that
// gpytorch.mlls.ExactMarginalLogLikelihood + torch.autograd perform on every objective evaluation of
// botorch.fit_gpytorch_model (gabo_sphere.py:162), and the (K + noise I)^-1 / alpha factors the acquisition needs:
//     K_theta = s exp(-beta Dm) + noise I            Dm = d^2 (Gaussian kernels) or d (Laplace kernels), n x n
//     ll      = -1/2 (r^T alpha + log det K_theta + n log 2 pi),    r = y - m,  alpha = K_theta^-1 r
//     dll/dp  = 1/2 tr((alpha alpha^T - K_theta^-1) dK/dp),  p in {beta, s, noise};   dll/dm = 1^T alpha
// The distance matrix is computed once per fit by the Gram kernels; every evaluation after that is this one launch.
// fp64 throughout; K_theta, its Cholesky factor L (lower triangle) and L^-1 (stored transposed in the strict upper
// triangle) share one n x (n+1) shared-memory tile, n <= 128.
#include "common.cuh"

namespace gabo {
namespace {

constexpr int kGpThreads = 256;      // upper bound; the launch uses 64 / 128 / 256 threads for n <= 32 / 64 / 128 so that
constexpr int kMaxGpTrain = 128;     // the per-column barriers of small problems synchronise 2 warps instead of 8

// CTA-wide barrier, or a warp barrier when the CTA is a single warp (n <= 32 in the fit kernel: the evaluation is a chain
// of ~200 short dependent steps, and a __syncthreads per step costs several times a __syncwarp)
template <bool kWarp>
__device__ __forceinline__ void bar() {
    if (kWarp) __syncwarp();
    else __syncthreads();
}

template <bool kWarp = false>
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    if (kWarp) return v;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < static_cast<int>(blockDim.x) / 32; ++i) t += red[i];
    return t;
}

// One evaluation by the whole CTA.  FROM_GRAM: dmat already holds the base-kernel matrix exp(-beta Dm) (as written by the
// fused Gram kernels) -- the factorisation-only form behind gabo_gp_factor.  Returns false when K_theta is not positive
// definite (uniform across the CTA).  On success ll and, when `want_grad`, g[0..3] = dll/d(beta, s, noise, mean) are
// valid in EVERY thread; alpha is left in vec[], and when out_kinv != nullptr K_theta^-1 is written there.
template <bool FROM_GRAM, bool kWarp = false>
__device__ __forceinline__ bool gp_eval_block(const double* __restrict__ dmat, int n, const double* __restrict__ y,
                                              double beta, double s, double noise, double mean, double* sm,
                                              bool want_grad, double* __restrict__ out_kinv, double& ll, double (&g)[4]) {
    const int ld = n + 1;                                   // odd row stride in 8-byte words: conflict-free columns
    double* A = sm;                                         // n x ld
    double* vec = A + n * ld;                               // n : residual -> z -> alpha
    double* dinv = vec + n;                                 // n : 1 / L_kk
    double* red = dinv + n;                                 // kGpThreads / 32
    const int tid = threadIdx.x, nthr = blockDim.x;
    bar<kWarp>();                                        // the previous evaluation's readers are done with sm
    for (int e = tid; e < n * n; e += nthr) {
        const int i = e / n, j = e % n;
        if (j <= i) A[i * ld + j] = fma(s, FROM_GRAM ? dmat[e] : exp(-beta * dmat[e]), (i == j) ? noise : 0.0);
    }
    for (int i = tid; i < n; i += nthr) vec[i] = y[i] - mean;
    bar<kWarp>();
    // --- Cholesky, right-looking, in place (lower triangle).  Two barriers per column: every thread reads the pivot
    // (published by the previous trailing update) and forms 1 / L_kk itself; L_kk is never stored (only 1 / L_kk is used
    // afterwards) and log det = -2 sum log(1 / L_kk) is summed in parallel after the loop. ---
    bool bad = false;
    for (int k = 0; k < n; ++k) {
        const double piv = A[k * ld + k];
        if (!(piv > 0.0)) {
            bad = true;
            break;                                          // uniform: every thread reads the same pivot
        }
        const double inv = rsqrt(piv);
        if (tid == 0) dinv[k] = inv;
        for (int i = k + 1 + tid; i < n; i += nthr) A[i * ld + k] *= inv;
        bar<kWarp>();
        // trailing update of the lower triangle: rows i > k, columns k < j <= i
        const int m = n - k - 1;
        for (int e = tid; e < m * m; e += nthr) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) A[i * ld + j] = fma(-A[i * ld + k], A[j * ld + k], A[i * ld + j]);
        }
        bar<kWarp>();
    }
    bar<kWarp>();
    if (bad) return false;
    double logdet = 0.0;
    for (int i = tid; i < n; i += nthr) logdet -= 2.0 * log(dinv[i]);
    logdet = block_sum<kWarp>(logdet, red);
    // --- L z = r (column-oriented forward substitution), quad = z^T z ---
    for (int k = 0; k < n; ++k) {
        const double zk = vec[k] * dinv[k];
        bar<kWarp>();
        if (tid == 0) vec[k] = zk;
        for (int i = k + 1 + tid; i < n; i += nthr) vec[i] = fma(-A[i * ld + k], zk, vec[i]);
        bar<kWarp>();
    }
    double quad = 0.0;
    for (int i = tid; i < n; i += nthr) quad = fma(vec[i], vec[i], quad);
    quad = block_sum<kWarp>(quad, red);
    // --- L^T alpha = z (backward substitution) ---
    for (int k = n - 1; k >= 0; --k) {
        const double ak = vec[k] * dinv[k];
        bar<kWarp>();
        if (tid == 0) vec[k] = ak;
        for (int i = tid; i < k; i += nthr) vec[i] = fma(-A[k * ld + i], ak, vec[i]);
        bar<kWarp>();
    }
    ll = -0.5 * (quad + logdet + n * 1.8378770664093453);   // log(2 pi)
    if (!want_grad && !out_kinv) return true;
    // --- X = L^-1, one column per thread, stored transposed in the strict upper triangle: A[j][i] = X_ij, i > j ---
    for (int j = tid; j < n; j += nthr) {
        for (int i = j + 1; i < n; ++i) {
            double acc = A[i * ld + j] * dinv[j], acc2 = 0.0;   // L_ij X_jj; two chains hide the DFMA latency
            int k = j + 1;
            for (; k + 1 < i; k += 2) {
                acc = fma(A[i * ld + k], A[j * ld + k], acc);
                acc2 = fma(A[i * ld + k + 1], A[j * ld + k + 1], acc2);
            }
            if (k < i) acc = fma(A[i * ld + k], A[j * ld + k], acc);
            A[j * ld + i] = -(acc + acc2) * dinv[i];
        }
    }
    bar<kWarp>();
    // --- K^-1 = X^T X entry by entry; gradient sums over the lower triangle ---
    double gb = 0.0, gs = 0.0, gn = 0.0, gm = 0.0;
    const int tri = n * (n + 1) / 2;
    for (int e = tid; e < tri; e += nthr) {
        // e -> (i, j), j <= i
        int i = static_cast<int>((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        while (i * (i + 1) / 2 > e) --i;
        const int j = e - i * (i + 1) / 2;
        // sum_{k >= i} X_ki X_kj with X_ii = dinv[i]
        double kin = dinv[i] * ((i == j) ? dinv[i] : A[j * ld + i]), kin2 = 0.0;
        int k = i + 1;
        for (; k + 1 < n; k += 2) {
            kin = fma(A[i * ld + k], A[j * ld + k], kin);
            kin2 = fma(A[i * ld + k + 1], A[j * ld + k + 1], kin2);
        }
        if (k < n) kin = fma(A[i * ld + k], A[j * ld + k], kin);
        kin += kin2;
        if (out_kinv) {
            out_kinv[i * n + j] = kin;
            out_kinv[j * n + i] = kin;
        }
        const double w = fma(vec[i], vec[j], -kin);
        const double dm = dmat[i * n + j];
        const double base = FROM_GRAM ? dm : exp(-beta * dm);
        const double mult = (i == j) ? 0.5 : 1.0;           // 1/2 tr(.) over both triangles
        gs = fma(mult * w, base, gs);
        gb = fma(mult * w, -s * dm * base, gb);
        if (i == j) gn = fma(0.5, w, gn);
    }
    for (int i = tid; i < n; i += nthr) gm += vec[i];
    if (want_grad) {
        g[0] = block_sum<kWarp>(gb, red);
        g[1] = block_sum<kWarp>(gs, red);
        g[2] = block_sum<kWarp>(gn, red);
        g[3] = block_sum<kWarp>(gm, red);
    }
    return true;
}

template <bool FROM_GRAM>
__global__ void __launch_bounds__(kGpThreads)
gp_mll_kernel(const double* __restrict__ dmat, int n, const double* __restrict__ y, const double* __restrict__ theta,
              double4 theta_val, double* __restrict__ out_ll, double* __restrict__ out_grad,
              double* __restrict__ out_alpha, double* __restrict__ out_kinv, int* __restrict__ flags) {
    extern __shared__ double sm[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t b = blockIdx.x;
    const double beta = FROM_GRAM ? theta_val.x : theta[b * 4 + 0], s = FROM_GRAM ? theta_val.y : theta[b * 4 + 1];
    const double noise = FROM_GRAM ? theta_val.z : theta[b * 4 + 2], mean = FROM_GRAM ? theta_val.w : theta[b * 4 + 3];
    double ll, g[4];
    const bool ok = gp_eval_block<FROM_GRAM>(dmat, n, y, beta, s, noise, mean, sm, out_grad != nullptr,
                                             out_kinv ? out_kinv + b * n * n : nullptr, ll, g);
    if (!ok) {                                              // not positive definite: NaN outputs + flag
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        if (tid == 0) {
            if (out_ll) out_ll[b] = nanv;
            flags[b] = 1;
        }
        if (out_grad && tid < 4) out_grad[b * 4 + tid] = nanv;
        if (out_alpha) for (int i = tid; i < n; i += nthr) out_alpha[b * n + i] = nanv;
        if (out_kinv) for (int e = tid; e < n * n; e += nthr) out_kinv[b * n * n + e] = nanv;
        return;
    }
    if (tid == 0) {
        if (out_ll) out_ll[b] = ll;
        flags[b] = 0;
        if (out_grad) {
            out_grad[b * 4 + 0] = g[0];
            out_grad[b * 4 + 1] = g[1];
            out_grad[b * 4 + 2] = g[2];
            out_grad[b * 4 + 3] = g[3];
        }
    }
    const double* alpha = sm + n * (n + 1);
    if (out_alpha) for (int i = tid; i < n; i += nthr) out_alpha[b * n + i] = alpha[i];
}

// ------------------------------------------------------------------------------------------------------------------
// Register-resident evaluation for n <= 32 by ONE warp (the fit kernel's evaluator): lane i owns row i of K_theta in
// registers, the right-looking Cholesky exchanges the pivot column with warp shuffles (no barriers, no shared-memory
// round trips: the block-wide form above spends ~95 us per evaluation at n = 32, almost all of it in barrier / integer
// index latency), then every lane solves K c = e_lane for ITS column of K^-1 from the factor broadcast out of shared
// memory (forward + backward substitution in registers, the same instruction stream in all lanes).  From the columns:
// alpha_j = c_j . r, quad = r . alpha, and the gradient traces 1/2 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij.
// Rows n .. 31 are padded with the identity (log 1 = 0, alpha = 0), so one fully unrolled 32 x 32 code path serves all n.
//   sm: Ls[32 * 33] factor rows, Bs[32 * 33] base kernel exp(-beta Dm) (lower triangle), Ds[32] = 1 / L_kk,
//       Cs[2 * 32] broadcast buffer (pivot column, double-buffered; later the residual and alpha).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kW = 32, kWld = 33;

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__device__ __noinline__ bool gp_eval_warp32(const double* __restrict__ dmat, int n, const double* __restrict__ y,
                                               double beta, double s, double noise, double mean, double* sm, double& ll,
                                               double (&g)[4]) {
    double* Ls = sm;
    double* Bs = sm + kW * kWld;
    double* Ds = Bs + kW * kWld;
    const int i = threadIdx.x & 31;
    double a[kW];
    __syncwarp();
    // fill row i of K_theta through shared memory with a ROLLED loop (32 unrolled fp64 exp() bodies would be scheduled
    // side by side and spill), then pull it into registers
#pragma unroll 1
    for (int j = 0; j < kW; ++j) {
        double v = (i == j) ? 1.0 : 0.0;                     // identity padding
        if (i < n && j <= i) {
            const double base = exp(-beta * __ldg(dmat + i * n + j));
            Bs[i * kWld + j] = base;
            v = fma(s, base, (i == j) ? noise : 0.0);
        }
        Ls[i * kWld + j] = v;
    }
#pragma unroll
    for (int j = 0; j < kW; ++j) a[j] = Ls[i * kWld + j];
    bool bad = false;
    double inv_own = 1.0;
    double* Cs = Ds + kW;                                    // the current column of L, broadcast through shared memory
#pragma unroll
    for (int k = 0; k < kW; ++k) {
        const double piv = shfl_d(a[k], k);
        bad = bad || !(piv > 0.0);
        const double inv = rsqrt(piv);
        if (i == k) inv_own = inv;
        const double lik = a[k] * inv;                       // L_ik for i >= k (the slots j > i of a row are never read)
        a[k] = lik;
        if (k + 1 < kW) {
            // column k goes through shared memory: one store per lane, then 16-byte broadcast loads (two L_jk each) --
            // a third of the instructions of one shuffle pair per entry, and the unrolled code stays inside the
            // instruction cache (ncu on the shuffle form: stall_no_instruction was the top stall)
            Cs[(k & 1) * kW + i] = lik;
            __syncwarp();
            const double* col = Cs + (k & 1) * kW;
#pragma unroll
            for (int j = k + 1; j < kW; ++j) a[j] = fma(-lik, col[j], a[j]);
        }
    }
    if (bad) return false;                                   // uniform: every lane saw the same pivots
    const double logdet = -2.0 * warp_sum(log(inv_own));
#pragma unroll
    for (int j = 0; j < kW; ++j) Ls[i * kWld + j] = a[j];
    Ds[i] = inv_own;
    __syncwarp();
    // column `i` of K^-1: forward then backward substitution on e_i
    double b[kW];
#pragma unroll
    for (int j = 0; j < kW; ++j) b[j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < kW; ++k) {
        const double yk = b[k] * Ds[k];
        b[k] = yk;
#pragma unroll
        for (int r = k + 1; r < kW; ++r) b[r] = fma(-Ls[r * kWld + k], yk, b[r]);
        __syncwarp();      // also a compiler fence: keeps the (loop-invariant) factor loads of later steps from being
    }                      // hoisted here, which is what made the fully unrolled solves spill
#pragma unroll
    for (int k = kW - 1; k >= 0; --k) {
        const double ck = b[k] * Ds[k];
        b[k] = ck;
#pragma unroll
        for (int r = 0; r < k; ++r) b[r] = fma(-Ls[k * kWld + r], ck, b[r]);
        __syncwarp();
    }
    // alpha_i = sum_r Kinv_ri res_r (K^-1 is symmetric), quad = res . alpha
    const double res = (i < n) ? y[i] - mean : 0.0;
    double* Rs = Cs;                                         // residual, then alpha, broadcast the same way
    __syncwarp();
    Rs[i] = res;
    __syncwarp();
    double al = 0.0;
#pragma unroll
    for (int r = 0; r < kW; ++r) al = fma(b[r], Rs[r], al);
    const double quad = warp_sum(res * al);
    Rs[kW + i] = al;
    __syncwarp();
    ll = -0.5 * (quad + logdet + n * 1.8378770664093453);
    // gradient traces over the full matrix (column i in this lane)
    double gb = 0.0, gs = 0.0, gn = 0.0;
#pragma unroll
    for (int r = 0; r < kW; ++r) {
        const double ar = Rs[kW + r];
        if (i < n && r < n) {
            const double w = 0.5 * fma(ar, al, -b[r]);
            const double base = (r >= i) ? Bs[r * kWld + i] : Bs[i * kWld + r];
            const double dm = __ldg(dmat + r * n + i);
            gs = fma(w, base, gs);
            gb = fma(w, -s * dm * base, gb);
            if (r == i) gn += w;
        }
    }
    g[0] = warp_sum(gb);
    g[1] = warp_sum(gs);
    g[2] = warp_sum(gn);
    g[3] = warp_sum(al);
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// The whole hyper-parameter fit in ONE launch: one CTA per start runs BFGS with a backtracking (Armijo) line search on
//     f(raw) = -(ll(theta(raw)) + log-priors(theta)) / n,    theta = (beta_min + softplus(raw_0), softplus(raw_1),
//                                                                      noise_min + softplus(raw_2), raw_3)
// -- botorch's objective (gpytorch softplus constraints, Gamma priors on the transformed values, division by n) that
// fit_gpytorch_model hands to scipy's L-BFGS-B.  With four parameters the dense 4 x 4 inverse-Hessian update IS L-BFGS with
// full memory.  Every evaluation is the CTA-wide gp_eval_block above, so a fit costs ~60 evaluations of ~10 us instead of
// ~36 launches + read-backs (7.7 ms measured in round 1).  All threads carry the same scalars (block_sum broadcasts).
// ------------------------------------------------------------------------------------------------------------------
struct FitOpts {
    double beta_min, noise_min;
    double prior_c[3], prior_r[3];     // Gamma(concentration, rate) on (beta, outputscale, noise); concentration <= 0: none
    int fixed[4];                      // != 0: the raw parameter is held at its start value
    int maxiter;
    double pgtol, ftol;                // scipy L-BFGS-B: max |g_i| <= pgtol, (f_k - f_k+1) <= ftol max(|f_k|, |f_k+1|, 1)
};

__device__ __forceinline__ double softplus_d(double r) { return log1p(exp(-fabs(r))) + fmax(r, 0.0); }
__device__ __forceinline__ double sigmoid_d(double r) { return r >= 0 ? 1.0 / (1.0 + exp(-r)) : exp(r) / (1.0 + exp(r)); }

// f and (optionally) its raw gradient; returns false when K is not PD.
template <bool kWarp>
__device__ __forceinline__ bool fit_objective(const double* dmat, int n, const double* y, const FitOpts& o,
                                              const double (&raw)[4], double* sm, bool want_grad, double& f,
                                              double (&gr)[4]) {
    const double th[4] = {o.beta_min + softplus_d(raw[0]), softplus_d(raw[1]), o.noise_min + softplus_d(raw[2]), raw[3]};
    double ll, g[4];
    if (kWarp) {
        if (!gp_eval_warp32(dmat, n, y, th[0], th[1], th[2], th[3], sm, ll, g)) return false;
    } else {
        if (!gp_eval_block<false, false>(dmat, n, y, th[0], th[1], th[2], th[3], sm, want_grad, nullptr, ll, g)) return false;
    }
    double lp = 0.0, dlp[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (o.prior_c[i] > 0.0) {
            const double c = o.prior_c[i], r = o.prior_r[i], v = th[i];
            lp += c * log(r) + (c - 1.0) * log(v) - r * v - lgamma(c);
            dlp[i] = (c - 1.0) / v - r;
        }
    }
    f = -(ll + lp) / n;
    if (!isfinite(f)) return false;
    if (want_grad) {
#pragma unroll
        for (int i = 0; i < 3; ++i) gr[i] = o.fixed[i] ? 0.0 : -(g[i] + dlp[i]) * sigmoid_d(raw[i]) / n;
        gr[3] = o.fixed[3] ? 0.0 : -g[3] / n;
    }
    return true;
}

template <bool kWarp>
__global__ void __launch_bounds__(kWarp ? 32 : kGpThreads, 1)
gp_fit_kernel(const double* __restrict__ dmat, int n, const double* __restrict__ y, const double* __restrict__ raw0,
              FitOpts o, double* __restrict__ out_raw, double* __restrict__ out_f, int* __restrict__ out_info) {
    extern __shared__ double sm[];
    const int64_t b = blockIdx.x;
    double x[4], g[4], H[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = raw0[b * 4 + i];
    double f;
    int evals = 1, iters = 0, status = 0;             // status: 0 converged (pgtol), 1 ftol, 2 maxiter, 3 line search, 4 not PD
    bool ok = fit_objective<kWarp>(dmat, n, y, o, x, sm, true, f, g);
    if (!ok) status = 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) H[i][j] = (i == j) ? 1.0 : 0.0;
    bool first = true;
    while (ok) {
        double gmax = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) gmax = fmax(gmax, fabs(g[i]));
        if (gmax <= o.pgtol) { status = 0; break; }
        if (iters >= o.maxiter) { status = 2; break; }
        double p[4], gp = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) s = fma(-H[i][j], g[j], s);
            p[i] = o.fixed[i] ? 0.0 : s;
            gp = fma(g[i], p[i], gp);
        }
        if (!(gp < 0.0)) {                              // not a descent direction: reset to steepest descent
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) H[i][j] = (i == j) ? 1.0 : 0.0;
                p[i] = -g[i];
            }
            gp = -(g[0] * g[0] + g[1] * g[1] + g[2] * g[2] + g[3] * g[3]);
        }
        double t = 1.0;
        if (first) {                                    // scipy's first step: min(1, 1 / |g|)
            const double gn = sqrt(-gp);
            t = fmin(1.0, 1.0 / gn);
        }
        // backtracking (Armijo) line search; the gradient is evaluated together with the value at every trial point: the
        // first trial (t = 1) is accepted in most quasi-Newton iterations, which then cost ONE evaluation
        double xn[4], gn4[4], fg = f;
        bool accepted = false;
        for (int ls = 0; ls < 30; ++ls) {
#pragma unroll
            for (int i = 0; i < 4; ++i) xn[i] = fma(t, p[i], x[i]);
            const bool good = fit_objective<kWarp>(dmat, n, y, o, xn, sm, true, fg, gn4);
            ++evals;
            if (good && fg <= f + 1e-4 * t * gp) { accepted = true; break; }
            t *= 0.5;
        }
        if (!accepted) { status = 3; break; }
        // BFGS update of the inverse Hessian with s = xn - x, yv = gn - g
        double sv[4], yv[4], sy = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sv[i] = xn[i] - x[i];
            yv[i] = gn4[i] - g[i];
            sy = fma(sv[i], yv[i], sy);
        }
        if (sy > 1e-12) {
            const double rho = 1.0 / sy;
            double Hy[4], yHy = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) s = fma(H[i][j], yv[j], s);
                Hy[i] = s;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) yHy = fma(yv[i], Hy[i], yHy);
            if (first) {                                // scale the initial Hessian (Nocedal & Wright 6.20)
                const double sc = sy / fmax(yv[0] * yv[0] + yv[1] * yv[1] + yv[2] * yv[2] + yv[3] * yv[3], 1e-300);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) H[i][j] = (i == j) ? sc : 0.0;
                    Hy[i] = sc * yv[i];
                }
                yHy = sc * (yv[0] * yv[0] + yv[1] * yv[1] + yv[2] * yv[2] + yv[3] * yv[3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    H[i][j] += -rho * (Hy[i] * sv[j] + sv[i] * Hy[j]) + rho * (1.0 + rho * yHy) * sv[i] * sv[j];
        }
        first = false;
        const double fold = f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i] = xn[i];
            g[i] = gn4[i];
        }
        f = fg;
        ++iters;
        if (fold - f <= o.ftol * fmax(fmax(fabs(fold), fabs(f)), 1.0)) { status = 1; break; }
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) out_raw[b * 4 + i] = x[i];
        out_f[b] = ok ? f : __longlong_as_double(0x7ff0000000000000LL);
        out_info[b * 3 + 0] = status;
        out_info[b * 3 + 1] = iters;
        out_info[b * 3 + 2] = evals;
    }
}

}  // namespace
}  // namespace gabo

using namespace gabo;

static unsigned gp_threads(int64_t n) { return n <= 32 ? 64u : (n <= 64 ? 128u : static_cast<unsigned>(kGpThreads)); }

// Function attributes are per device / context: set on every launch that needs more than the 48 KB default (as the
// acquisition launchers do); returns false (error recorded) when the driver refuses.
template <bool FACTOR_ONLY>
static bool configure_smem(size_t smem) {
    if (smem <= 48u * 1024u) return true;
    const cudaError_t e = cudaFuncSetAttribute(gp_mll_kernel<FACTOR_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(smem));
    if (e != cudaSuccess) {
        set_error("gp_mll_kernel: cudaFuncSetAttribute(%zu bytes of shared memory): %s", smem, cudaGetErrorString(e));
        return false;
    }
    return true;
}

extern "C" int gabo_gp_mll(const double* dmat, int64_t n, const double* y, const double* theta, int64_t batch,
                           double* out_ll, double* out_grad, double* out_alpha, double* out_kinv, int* flags,
                           void* stream) {
    GABO_REQUIRE(n >= 0 && batch >= 0, GABO_E_ARG, "gabo_gp_mll: negative size");
    GABO_REQUIRE(n >= 1 && n <= kMaxGpTrain, GABO_E_ARG, "gabo_gp_mll: n=%lld outside [1, %d]",
                 static_cast<long long>(n), kMaxGpTrain);
    if (batch == 0) return GABO_OK;
    GABO_REQUIRE(dmat && y && theta && out_ll && flags, GABO_E_ARG, "gabo_gp_mll: null pointer");
    const size_t smem = sizeof(double) * (static_cast<size_t>(n) * (n + 1) + 2 * n + kGpThreads / 32);
    if (!configure_smem<false>(smem)) return GABO_E_CUDA;
    gp_mll_kernel<false><<<static_cast<unsigned>(batch), gp_threads(n), smem, static_cast<cudaStream_t>(stream)>>>(
        dmat, static_cast<int>(n), y, theta, make_double4(0, 0, 0, 0), out_ll, out_grad, out_alpha, out_kinv, flags);
    return check_launch("gp_mll_kernel");
}

extern "C" int gabo_gp_factor(const double* kmat, int64_t n, const double* y, double outputscale, double noise,
                              double mean, double* out_alpha, double* out_kinv, int* flag, void* stream) {
    GABO_REQUIRE(n >= 1 && n <= kMaxGpTrain, GABO_E_ARG, "gabo_gp_factor: n=%lld outside [1, %d]",
                 static_cast<long long>(n), kMaxGpTrain);
    GABO_REQUIRE(kmat && y && out_alpha && out_kinv && flag, GABO_E_ARG, "gabo_gp_factor: null pointer");
    const size_t smem = sizeof(double) * (static_cast<size_t>(n) * (n + 1) + 2 * n + kGpThreads / 32);
    if (!configure_smem<true>(smem)) return GABO_E_CUDA;
    gp_mll_kernel<true><<<1, gp_threads(n), smem, static_cast<cudaStream_t>(stream)>>>(
        kmat, static_cast<int>(n), y, nullptr, make_double4(0.0, outputscale, noise, mean), nullptr, nullptr, out_alpha,
        out_kinv, flag);
    return check_launch("gp_mll_kernel");
}

extern "C" int gabo_gp_fit(const double* dmat, int64_t n, const double* y, const double* raw0, int64_t batch,
                           double beta_min, double noise_min, const double* priors, const int* fixed, int maxiter,
                           double pgtol, double ftol, double* out_raw, double* out_f, int* out_info, void* stream) {
    GABO_REQUIRE(n >= 1 && n <= kMaxGpTrain, GABO_E_ARG, "gabo_gp_fit: n=%lld outside [1, %d]",
                 static_cast<long long>(n), kMaxGpTrain);
    GABO_REQUIRE(batch >= 0, GABO_E_ARG, "gabo_gp_fit: negative batch");
    if (batch == 0) return GABO_OK;
    GABO_REQUIRE(dmat && y && raw0 && priors && fixed && out_raw && out_f && out_info, GABO_E_ARG,
                 "gabo_gp_fit: null pointer");
    GABO_REQUIRE(maxiter >= 0, GABO_E_ARG, "gabo_gp_fit: negative maxiter");
    FitOpts o;
    o.beta_min = beta_min;
    o.noise_min = noise_min;
    for (int i = 0; i < 3; ++i) {
        o.prior_c[i] = priors[2 * i];
        o.prior_r[i] = priors[2 * i + 1];
    }
    for (int i = 0; i < 4; ++i) o.fixed[i] = fixed[i];
    o.maxiter = maxiter;
    o.pgtol = pgtol;
    o.ftol = ftol;
    const size_t smem = sizeof(double) * (static_cast<size_t>(n) * (n + 1) + 2 * n + kGpThreads / 32);
    if (smem > 48u * 1024u) {
        const cudaError_t e = cudaFuncSetAttribute(gp_fit_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(smem));
        GABO_REQUIRE(e == cudaSuccess, GABO_E_CUDA, "gp_fit_kernel: cudaFuncSetAttribute(%zu bytes): %s", smem,
                     cudaGetErrorString(e));
    }
    if (n <= 32) {   // a single warp per start, matrix rows in registers
        const size_t smem_w = sizeof(double) * (2 * kW * kWld + 3 * kW);
        gp_fit_kernel<true><<<static_cast<unsigned>(batch), 32, smem_w, static_cast<cudaStream_t>(stream)>>>(
            dmat, static_cast<int>(n), y, raw0, o, out_raw, out_f, out_info);
    } else {
        gp_fit_kernel<false><<<static_cast<unsigned>(batch), gp_threads(n), smem, static_cast<cudaStream_t>(stream)>>>(
            dmat, static_cast<int>(n), y, raw0, o, out_raw, out_f, out_info);
    }
    return check_launch("gp_fit_kernel");
}
