// Mandel conversions, batched logm and the Frobenius / log-Euclidean Gram (G3 and the "next" kernels of SURVEY 8f).
//
//   gabo_mandel_unpack / gabo_mandel_pack : vector_to_symmetric_matrix_mandel_torch / symmetric_matrix_to_vector_mandel_torch
//                                           (Riemannian_utils/spd_utils_torch.py:159-194 / :197-226), a Python double loop
//                                           in the reference; here one thread per matrix entry, coalesced both ways.
//   gabo_spd_logm                         : logm_torch (:13-30) for a batch: Cholesky X = L L^T, one-sided Jacobi on L
//                                           (L V = U Sigma, so X = U Sigma^2 U^T), logm X = sum_k log(s_k^2)/s_k^2 g_k g_k^T.
//   gabo_frobenius_gram                   : frobenius_distance_torch (:124-156) + exp(-d^2 / l^2) (kernels_spd.py:230-241,
//                                           :283-313); the reference adds 1e-15 to EVERY entry of the difference (:156).
#include "spd_common.cuh"

namespace gabo {
namespace {

// position -> (row, col) of a Mandel entry, diagonal by diagonal
__device__ __forceinline__ void mandel_rc(int d, int pos, int& r, int& c) {
    int k = 0;
    int len = d;
    while (pos >= len) {
        pos -= len;
        --len;
        ++k;
    }
    r = pos;
    c = pos + k;
}

__global__ void mandel_unpack_kernel(const double* __restrict__ vec, int64_t n, int d, double* __restrict__ mat) {
    const int dd = d * d;
    const int dv = d * (d + 1) / 2;
    const int64_t total = n * dd;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = e / dd;
        const int rc = static_cast<int>(e - i * dd);
        int r = rc / d, c = rc % d;
        if (r > c) {
            const int t = r;
            r = c;
            c = t;
        }
        const double v = vec[i * dv + mandel_pos(d, r, c)];
        mat[e] = (r == c) ? v : v / 1.4142135623730951;  // spd_utils_torch.py:186-187 divides by 2.0**0.5
    }
}

__global__ void mandel_pack_kernel(const double* __restrict__ mat, int64_t n, int d, double* __restrict__ vec) {
    const int dd = d * d;
    const int dv = d * (d + 1) / 2;
    const int64_t total = n * dv;
    for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = e / dv;
        const int pos = static_cast<int>(e - i * dv);
        int r, c;
        mandel_rc(d, pos, r, c);
        const double* m = mat + i * dd;
        if (r == c) {
            vec[e] = m[r * d + r];
        } else {
            // spd_utils_torch.py:219: 0.5 * (2**0.5 * upper + 2**0.5 * lower)
            const double s2 = 1.4142135623730951;
            // explicit roundings (no FMA contraction): bit-identical to the reference's torch expression
            vec[e] = __dmul_rn(0.5, __dadd_rn(__dmul_rn(s2, m[r * d + c]), __dmul_rn(s2, m[c * d + r])));
        }
    }
}

template <int d>
__global__ void spd_logm_kernel(const double* __restrict__ mat, int64_t n, double* __restrict__ out) {
    constexpr int TRI = tri_size(d);
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* m = mat + i * d * d;
    double L[TRI], A[TRI];
    // symeig(upper=True) reads the upper triangle (spd_utils_torch.py:25)
    const bool ok = chol_inv<d>([&](int r, int c) { return m[c * d + r]; }, L, A);
    double G[d][d];
    tri_expand<d, double>([&](int e) { return L[e]; }, G);
    double lam[d];
    jacobi_onesided<d, double>(G, lam);
    double f[d];
#pragma unroll
    for (int k = 0; k < d; ++k) f[k] = log(lam[k]) / lam[k];
    double C[d][d];
    weighted_outer<d, double>(G, f, C);
    double* o = out + i * d * d;
    const double nanv = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) o[r * d + c] = C[r][c] + nanv;
}

template <typename OutT, int KIND>
__global__ void __launch_bounds__(128)
    frobenius_gram_kernel(const double* __restrict__ m1, int64_t n1, const double* __restrict__ m2, int64_t n2, int dd,
                          double param, OutT* __restrict__ out, int64_t ld_out) {
    extern __shared__ __align__(16) double tile[];  // 16 rows of m1
    constexpr int kRows = 16;
    const int64_t tiles_j = (n2 + 127) / 128;
    const int64_t tiles_i = (n1 + kRows - 1) / kRows;
    for (int64_t t = blockIdx.x; t < tiles_i * tiles_j; t += gridDim.x) {
        const int64_t jb = t / tiles_i, ib = t % tiles_i;
        const int64_t i0 = ib * kRows;
        const int rows = static_cast<int>(imin(kRows, n1 - i0));
        __syncthreads();
        for (int e = threadIdx.x; e < rows * dd; e += blockDim.x) tile[e] = m1[i0 * dd + e];
        __syncthreads();
        const int64_t j = jb * 128 + threadIdx.x;
        if (j >= n2) continue;
        const double* b = m2 + j * dd;
        for (int i = 0; i < rows; ++i) {
            double s = 0.0;
            for (int e = 0; e < dd; ++e) {
                const double df = tile[i * dd + e] - __ldg(b + e) + 1e-15;  // spd_utils_torch.py:156
                s = fma(df, df, s);
            }
            double v;
            if (KIND == GABO_KIND_DIST) v = sqrt(s);
            else if (KIND == GABO_KIND_GAUSS) v = exp(-param * s);  // d*d == s up to one rounding
            else v = exp(-param * sqrt(s));
            st_cs(out + (i0 + i) * ld_out + j, static_cast<OutT>(v));
        }
    }
}

template <typename OutT>
int launch_frob(const double* m1, int64_t n1, const double* m2, int64_t n2, int d, double param, int kind, void* out,
                int64_t ld_out, cudaStream_t s) {
    const int dd = d * d;
    const int64_t tiles = ((n1 + 15) / 16) * ((n2 + 127) / 128);
    const unsigned grid = static_cast<unsigned>(imin(tiles, static_cast<int64_t>(sm_count()) * 8));
    const size_t smem = sizeof(double) * 16 * dd;
    OutT* o = static_cast<OutT*>(out);
    if (smem > 48 * 1024) {
        const int b = static_cast<int>(smem);
        cudaFuncSetAttribute(frobenius_gram_kernel<OutT, GABO_KIND_GAUSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
        cudaFuncSetAttribute(frobenius_gram_kernel<OutT, GABO_KIND_LAPLACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
        cudaFuncSetAttribute(frobenius_gram_kernel<OutT, GABO_KIND_DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, b);
    }
    switch (kind) {
        case GABO_KIND_GAUSS:
            frobenius_gram_kernel<OutT, GABO_KIND_GAUSS><<<grid, 128, smem, s>>>(m1, n1, m2, n2, dd, param, o, ld_out);
            break;
        case GABO_KIND_LAPLACE:
            frobenius_gram_kernel<OutT, GABO_KIND_LAPLACE><<<grid, 128, smem, s>>>(m1, n1, m2, n2, dd, param, o, ld_out);
            break;
        default:
            frobenius_gram_kernel<OutT, GABO_KIND_DIST><<<grid, 128, smem, s>>>(m1, n1, m2, n2, dd, param, o, ld_out);
            break;
    }
    return check_launch("frobenius_gram_kernel");
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_mandel_unpack(const double* vec, int64_t n, int d, double* mat, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0 && d >= 1 && d <= 64, GABO_E_ARG, "gabo_mandel_unpack: bad size (n=%lld, d=%d)",
                 static_cast<long long>(n), d);
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(vec && mat, GABO_E_ARG, "gabo_mandel_unpack: null pointer");
    const int64_t total = n * d * d;
    const unsigned grid = static_cast<unsigned>(imin((total + 255) / 256, static_cast<int64_t>(sm_count()) * 16));
    mandel_unpack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(vec, n, d, mat);
    return check_launch("mandel_unpack_kernel");
}

extern "C" int gabo_mandel_pack(const double* mat, int64_t n, int d, double* vec, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0 && d >= 1 && d <= 64, GABO_E_ARG, "gabo_mandel_pack: bad size (n=%lld, d=%d)",
                 static_cast<long long>(n), d);
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(vec && mat, GABO_E_ARG, "gabo_mandel_pack: null pointer");
    const int64_t total = n * (d * (d + 1) / 2);
    const unsigned grid = static_cast<unsigned>(imin((total + 255) / 256, static_cast<int64_t>(sm_count()) * 16));
    mandel_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mat, n, d, vec);
    return check_launch("mandel_pack_kernel");
}

extern "C" int gabo_spd_logm(const double* mat, int64_t n, int d, double* out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_spd_logm: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(mat && out, GABO_E_ARG, "gabo_spd_logm: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_logm: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((n + 63) / 64);
    switch (d) {
#define GABO_CASE(DD)                                         \
    case DD:                                                  \
        spd_logm_kernel<DD><<<grid, 64, 0, s>>>(mat, n, out); \
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
    return check_launch("spd_logm_kernel");
}

extern "C" int gabo_frobenius_gram(const double* m1, int64_t n1, const double* m2, int64_t n2, int d, double param,
                                   int kind, void* out, int out_dtype, int64_t ld_out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n1 >= 0 && n2 >= 0, GABO_E_ARG, "gabo_frobenius_gram: negative size");
    if (n1 == 0 || n2 == 0) return GABO_OK;
    GABO_REQUIRE(m1 && m2 && out, GABO_E_ARG, "gabo_frobenius_gram: null pointer");
    GABO_REQUIRE(d >= 1 && d <= 32, GABO_E_ARG, "gabo_frobenius_gram: d=%d outside [1, 32]", d);
    GABO_REQUIRE(kind >= GABO_KIND_GAUSS && kind <= GABO_KIND_DIST, GABO_E_ARG, "gabo_frobenius_gram: bad kind");
    GABO_REQUIRE(ld_out >= n2, GABO_E_ARG, "gabo_frobenius_gram: ld_out < n2");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (out_dtype == GABO_F32) return launch_frob<float>(m1, n1, m2, n2, d, param, kind, out, ld_out, s);
    if (out_dtype == GABO_F64) return launch_frob<double>(m1, n1, m2, n2, d, param, kind, out, ld_out, s);
    set_error("gabo_frobenius_gram: bad out_dtype");
    return GABO_E_ARG;
}
