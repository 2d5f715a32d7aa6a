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

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < static_cast<int>(blockDim.x) / 32; ++i) t += red[i];
    return t;
}

// FROM_GRAM: dmat already holds the base-kernel matrix exp(-beta Dm) (as written by the fused Gram kernels) and the
// hyper-parameters come by value -- the factorisation-only form behind gabo_gp_factor.
template <bool FROM_GRAM>
__global__ void __launch_bounds__(kGpThreads)
gp_mll_kernel(const double* __restrict__ dmat, int n, const double* __restrict__ y, const double* __restrict__ theta,
              double4 theta_val, double* __restrict__ out_ll, double* __restrict__ out_grad,
              double* __restrict__ out_alpha, double* __restrict__ out_kinv, int* __restrict__ flags) {
    extern __shared__ double sm[];
    const int ld = n + 1;                                   // odd row stride in 8-byte words: conflict-free columns
    double* A = sm;                                         // n x ld
    double* vec = A + n * ld;                               // n : residual -> z -> alpha
    double* dinv = vec + n;                                 // n : 1 / L_kk
    double* red = dinv + n;                                 // kGpThreads / 32
    __shared__ int bad;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t b = blockIdx.x;
    const double beta = FROM_GRAM ? theta_val.x : theta[b * 4 + 0], s = FROM_GRAM ? theta_val.y : theta[b * 4 + 1];
    const double noise = FROM_GRAM ? theta_val.z : theta[b * 4 + 2], mean = FROM_GRAM ? theta_val.w : theta[b * 4 + 3];
    if (tid == 0) bad = 0;
    for (int e = tid; e < n * n; e += nthr) {
        const int i = e / n, j = e % n;
        if (j <= i) A[i * ld + j] = fma(s, FROM_GRAM ? dmat[e] : exp(-beta * dmat[e]), (i == j) ? noise : 0.0);
    }
    for (int i = tid; i < n; i += nthr) vec[i] = y[i] - mean;
    __syncthreads();
    // --- Cholesky, right-looking, in place (lower triangle) ---
    double logdet = 0.0;
    for (int k = 0; k < n; ++k) {
        const double piv = A[k * ld + k];
        if (!(piv > 0.0)) {
            if (tid == 0) bad = 1;
            break;                                          // uniform: every thread reads the same pivot
        }
        const double lkk = sqrt(piv), inv = 1.0 / lkk;
        logdet += log(piv);                                 // = 2 log L_kk
        __syncthreads();
        if (tid == 0) {
            A[k * ld + k] = lkk;
            dinv[k] = inv;
        }
        for (int i = k + 1 + tid; i < n; i += nthr) A[i * ld + k] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: rows i > k, columns k < j <= i
        const int m = n - k - 1;
        for (int e = tid; e < m * m; e += nthr) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) A[i * ld + j] = fma(-A[i * ld + k], A[j * ld + k], A[i * ld + j]);
        }
        __syncthreads();
    }
    __syncthreads();
    if (bad) {                                              // not positive definite: NaN outputs + flag
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
    // --- L z = r (column-oriented forward substitution), quad = z^T z ---
    for (int k = 0; k < n; ++k) {
        const double zk = vec[k] * dinv[k];
        __syncthreads();
        if (tid == 0) vec[k] = zk;
        for (int i = k + 1 + tid; i < n; i += nthr) vec[i] = fma(-A[i * ld + k], zk, vec[i]);
        __syncthreads();
    }
    double quad = 0.0;
    for (int i = tid; i < n; i += nthr) quad = fma(vec[i], vec[i], quad);
    quad = block_sum(quad, red);
    // --- L^T alpha = z (backward substitution) ---
    for (int k = n - 1; k >= 0; --k) {
        const double ak = vec[k] * dinv[k];
        __syncthreads();
        if (tid == 0) vec[k] = ak;
        for (int i = tid; i < k; i += nthr) vec[i] = fma(-A[k * ld + i], ak, vec[i]);
        __syncthreads();
    }
    const double ll = -0.5 * (quad + logdet + n * 1.8378770664093453);   // log(2 pi)
    if (tid == 0) {
        if (out_ll) out_ll[b] = ll;
        flags[b] = 0;
    }
    if (out_alpha) for (int i = tid; i < n; i += nthr) out_alpha[b * n + i] = vec[i];
    if (!out_grad && !out_kinv) return;
    // --- X = L^-1, one column per thread, stored transposed in the strict upper triangle: A[j][i] = X_ij, i > j ---
    for (int j = tid; j < n; j += nthr) {
        for (int i = j + 1; i < n; ++i) {
            double acc = A[i * ld + j] * dinv[j];           // L_ij X_jj
            for (int k = j + 1; k < i; ++k) acc = fma(A[i * ld + k], A[j * ld + k], acc);
            A[j * ld + i] = -acc * dinv[i];
        }
    }
    __syncthreads();
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
        double kin = dinv[i] * ((i == j) ? dinv[i] : A[j * ld + i]);
        for (int k = i + 1; k < n; ++k) kin = fma(A[i * ld + k], A[j * ld + k], kin);
        if (out_kinv) {
            out_kinv[b * n * n + i * n + j] = kin;
            out_kinv[b * n * n + j * n + i] = kin;
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
    if (out_grad) {
        gb = block_sum(gb, red);
        gs = block_sum(gs, red);
        gn = block_sum(gn, red);
        gm = block_sum(gm, red);
        if (tid == 0) {
            out_grad[b * 4 + 0] = gb;
            out_grad[b * 4 + 1] = gs;
            out_grad[b * 4 + 2] = gn;
            out_grad[b * 4 + 3] = gm;
        }
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
