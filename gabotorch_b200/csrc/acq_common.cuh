// Shared pieces of the acquisition kernels (A1 / A4 of SURVEY.md section 8): GP description on the device, the analytic
// Expected Improvement, the per-warp conjugate-gradient bookkeeping.
//
// What the reference does per restart (manifold_optimization/manifold_optimize.py:207-221): a pymanopt solve whose every
// cost / gradient call runs botorch's ExpectedImprovement through gpytorch's exact-GP posterior with torch autograd
// (pymanopt_addons/tools/autodiff/_pytorch.py:83-101).  Here ONE WARP owns one restart for the whole solve:
//   - lanes are spread over the GP training points (distance, kernel value, one row of the K^-1 mat-vec each),
//   - the posterior mean / variance are warp reductions, EI and its derivative are closed-form scalars,
//   - the Riemannian gradient is assembled from the same per-point quantities:
//         grad_x EI = 2 beta sum_i w_i k_i Log_x(X_i),   w_i = -Phi(u) alpha_i - phi(u) (M k)_i / sigma,
//   - the CG recurrences (Hestenes-Stiefel, pymanopt's adaptive backtracking line search) are warp-uniform scalars.
#pragma once
#include "common.cuh"

namespace gabo {

struct GpParams {
    int n;        // training points
    int dim;      // sphere: ambient D; spd: d
    double mean, outputscale, beta, best_f, kxx;
    const double* x_train;
    const double* alpha;
    const double* minv;
};

struct RcgParams {
    int maxiter, ls_maxiter;
    double mingradnorm, minstepsize, contraction, suff_decr, initial_stepsize;
};

struct RtrParams {
    int maxiter, mininner, maxinner;
    double mingradnorm, kappa, theta, rho_prime, rho_regularization, delta_bar, delta0, fd_eps;
};

struct CtrParams {
    RtrParams tr;
    int n_cons;        // 0 (plain TrustRegions), 1 or 2 eigenvalue constraints
    int strict;        // StrictConstrainedTrustRegions
    int kind[2];       // GABO_CONS_MAX_EIG | GABO_CONS_MIN_EIG
    double bound[2];
    double delta_cons; // Delta_cons of ConstrainedTrustRegions.solve (1e-6)
};

template <typename T>
struct M;
template <>
struct M<float> {
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float log_(float x) { return logf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
    static __device__ __forceinline__ float erfc_(float x) { return erfcf(x); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
    static __device__ __forceinline__ float clamp_eps() { return 0.0f; }  // 1e-15 is below fp32 resolution at 1
    // theta^2 = acos(c)^2 for c in [-1, 1], branch-free: polynomial for the near side, pi - sqrt for the far side
    // (same arithmetic as the fused Gram kernel; ~16 instructions instead of ~35 for acosf).
    static __device__ __forceinline__ float theta2_(float c) {
        const float r2 = 4.0f * asin2_sqrt(0.5f * (1.0f - fabsf(c)));
        const float far = (3.14159274101257324f - sqrt_approx(r2)) + (-8.74227765734758577e-8f);
        return (c < 0.0f) ? far * far : r2;
    }
    static __device__ __forceinline__ float theta_(float th2) { return sqrt_approx(th2); }
    static __device__ __forceinline__ float exp_neg_(float x) { return ex2_approx(x * 1.44269504088896341f); }  // x <= 0
    static __device__ __forceinline__ float rsqrt_(float x) {
        const float y = rsqrt_approx(x);
        return y * fmaf(-0.5f * x * y, y, 1.5f);
    }
};
template <>
struct M<double> {
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double acos_(double x) { return acos(x); }
    static __device__ __forceinline__ double erfc_(double x) { return erfc(x); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
    static __device__ __forceinline__ double clamp_eps() { return 1e-15; }  // sphere_utils_torch.py:53
    static __device__ __forceinline__ double theta2_(double c) {
        const double t = acos(c);
        return t * t;
    }
    static __device__ __forceinline__ double theta_(double th2) { return sqrt(th2); }
    static __device__ __forceinline__ double exp_neg_(double x) { return exp(x); }
    static __device__ __forceinline__ double rsqrt_(double x) { return 1.0 / sqrt(x); }
};

// botorch analytic EI, maximize=False:  sigma = sqrt(clamp_min(var, 1e-9)),  u = (best_f - mu) / sigma,
// EI = sigma (phi(u) + u Phi(u)).  Also returns the two factors of the gradient weights.
template <typename T>
struct EiScalars {
    T ei;
    T cdf;            // Phi(u)            : -dEI/dmu
    T pdf_over_sigma; // phi(u) / sigma    : dEI/dsigma / sigma, zero when the variance floor is active
};

template <typename T>
__device__ __forceinline__ EiScalars<T> ei_scalars(T k_alpha, T k_m_k, const GpParams& gp) {
    EiScalars<T> r;
    const T mu = static_cast<T>(gp.mean) + k_alpha;
    const T var_raw = static_cast<T>(gp.outputscale * gp.kxx) - k_m_k;
    const bool clamped = !(var_raw >= T(1e-9));
    const T var = clamped ? T(1e-9) : var_raw;
    const T rs = M<T>::rsqrt_(var);   // one reciprocal square root serves sigma, u and pdf / sigma (no divisions)
    const T sigma = var * rs;
    const T u = (static_cast<T>(gp.best_f) - mu) * rs;
    const T pdf = M<T>::exp_neg_(T(-0.5) * u * u) * T(0.3989422804014326779);
    const T cdf = T(0.5) * M<T>::erfc_(-u * T(0.7071067811865475244));
    r.ei = sigma * (pdf + u * cdf);
    r.cdf = cdf;
    r.pdf_over_sigma = clamped ? T(0) : pdf * rs;
    return r;
}

// Dynamic shared-memory carving helper (all offsets 16-byte aligned).
struct SmemCarver {
    size_t off = 0;
    __host__ __device__ size_t take(size_t bytes) {
        const size_t o = off;
        off = (off + bytes + 15) & ~static_cast<size_t>(15);
        return o;
    }
};

constexpr int kAcqWarps = 4;  // restarts per CTA

int launch_acq_sphere(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rcg_opts* opts, double* value,
                      double* grad, int32_t* iters, int32_t* reason, cudaStream_t stream);

int launch_rtr_sphere(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rtr_opts* opts, double* value,
                      int32_t* iters, int32_t* reason, cudaStream_t stream);

template <int d>
int launch_acq_spd(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rcg_opts* opts, double* value,
                   double* grad, int32_t* iters, int32_t* reason, cudaStream_t stream);

template <int d>
int launch_rtr_spd(const gabo_gp_desc* gp, double* x, int64_t r, const CtrParams& opt, double* value, int32_t* iters,
                   int32_t* reason, cudaStream_t stream);

}  // namespace gabo
