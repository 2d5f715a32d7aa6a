// Input gradient of the SPD affine-invariant Gram: what the reference obtains by running
// torch.autograd through affine_invariant_distance_torch (Riemannian_utils/spd_utils_torch.py:53-120; the per-pair
// symeig is called with eigenvectors=True "for derivation", :110-111).
//
// For X = L L^T, A = L^-1 and S(X, Y) = logm(A Y A^T):   d^2(X, Y) = |S|_F^2   and the Euclidean gradient with respect
// to the (symmetric) matrix X is    grad_X d^2 = -2 X^-1 Log_X(Y) X^-1 = -2 A^T S A.
// Given upstream weights w_ij = dLoss / d(d_ij^2) this kernel returns, per row point i,
//     out_i = -2 A_i^T ( sum_j w_ij S(X_i, X2_j) ) A_i                      (d x d, fp64)
// whose Mandel vector is the gradient with respect to the Mandel input (the Mandel map is an isometry).
// One WARP per row point: lanes stride over the columns, each running the same in-register one-sided Jacobi as the
// Gram kernel on G = A_i L_j (log map = sum_k log(l_k)/l_k g_k g_k^T), partial sums in fp64 registers, a fixed-order
// butterfly at the end (deterministic), then the d x d sandwich in fp64.
#include "spd_common.cuh"

namespace gabo {
namespace {

constexpr int kWarpsPerCta = 4;

template <int d, typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    spd_ai_logsum_kernel(const double* __restrict__ fac1, int64_t n1, const double* __restrict__ fac2, int64_t n2,
                         const double* __restrict__ w, int64_t ld_w, int transpose_w, double* __restrict__ out) {
    constexpr int TRI = tri_size(d);
    constexpr int FS = factor_stride(d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + warp;
    if (i >= n1) return;
    const double* Ai = fac1 + i * FS + TRI;   // A = L^-1, packed lower-triangular
    double acc[TRI];
#pragma unroll
    for (int e = 0; e < TRI; ++e) acc[e] = 0.0;
    for (int64_t j = lane; j < n2; j += 32) {
        const double wij = transpose_w ? w[j * ld_w + i] : w[i * ld_w + j];
        if (wij == 0.0) continue;
        const double* Lj = fac2 + j * FS;
        T G[d][d];
        tri_product<d, T>([&](int e) { return __ldg(Ai + e); }, [&](int e) { return __ldg(Lj + e); }, G);
        T lam[d];
        jacobi_onesided_compact<d, T>(G, lam);
        T f[d];
#pragma unroll
        for (int k = 0; k < d; ++k) f[k] = static_cast<T>(log(static_cast<double>(lam[k]))) / lam[k];
#pragma unroll
        for (int r = 0; r < d; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k < d; ++k) s = fma(f[k] * G[r][k], G[c][k], s);
                acc[tri_idx(r, c)] = fma(wij, static_cast<double>(s), acc[tri_idx(r, c)]);
            }
    }
#pragma unroll
    for (int e = 0; e < TRI; ++e) acc[e] = warp_sum(acc[e]);
    // out = -2 A^T T A:  B = T A (d x d, T symmetric from its lower triangle), out[p][q] = -2 sum_r A[r][p] B[r][q]
    for (int e = lane; e < d * d; e += 32) {
        const int p = e / d, q = e % d;
        double s = 0.0;
        for (int r = p; r < d; ++r) {            // A[r][p] != 0 only for r >= p
            double b = 0.0;
            for (int c = q; c < d; ++c) {        // A[c][q] != 0 only for c >= q
                const double t = (r >= c) ? acc[tri_idx(r, c)] : acc[tri_idx(c, r)];
                b = fma(t, __ldg(Ai + tri_idx(c, q)), b);
            }
            s = fma(__ldg(Ai + tri_idx(r, p)), b, s);
        }
        out[i * d * d + e] = -2.0 * s;
    }
}

template <int d>
int launch(const double* fac1, int64_t n1, const double* fac2, int64_t n2, const double* w, int64_t ld_w, int transpose_w,
           int compute, double* out, cudaStream_t s) {
    const unsigned grid = static_cast<unsigned>((n1 + kWarpsPerCta - 1) / kWarpsPerCta);
    if (compute == GABO_F64)
        spd_ai_logsum_kernel<d, double><<<grid, kWarpsPerCta * 32, 0, s>>>(fac1, n1, fac2, n2, w, ld_w, transpose_w, out);
    else
        spd_ai_logsum_kernel<d, float><<<grid, kWarpsPerCta * 32, 0, s>>>(fac1, n1, fac2, n2, w, ld_w, transpose_w, out);
    return check_launch("spd_ai_logsum_kernel");
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_spd_ai_gram_backward(const double* fac1, int64_t n1, const double* fac2, int64_t n2, int d,
                                         const double* w, int64_t ld_w, int transpose_w, int compute, double* out,
                                         void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n1 >= 0 && n2 >= 0, GABO_E_ARG, "gabo_spd_ai_gram_backward: negative size");
    if (n1 == 0) return GABO_OK;
    GABO_REQUIRE(fac1 && out && (n2 == 0 || (fac2 && w)), GABO_E_ARG, "gabo_spd_ai_gram_backward: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_ai_gram_backward: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    GABO_REQUIRE(compute == GABO_F32 || compute == GABO_F64, GABO_E_ARG, "gabo_spd_ai_gram_backward: bad compute dtype");
    GABO_REQUIRE(ld_w >= (transpose_w ? n1 : n2), GABO_E_ARG, "gabo_spd_ai_gram_backward: ld_w too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (d) {
#define GABO_CASE(DD) \
    case DD:          \
        return launch<DD>(fac1, n1, fac2, n2, w, ld_w, transpose_w, compute, out, s);
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
