// Backward kernels of the distance / projection operators (every kernel of the reference is
// differentiable with respect to its inputs and its manifold-valued parameters under torch.autograd --
// kernel_utils/kernels_sphere.py:71-134, kernels_spd.py:72-313, kernels_nested_spd.py:104-246).
//
//   gabo_weighted_points_sum : out_i = sum_j W_ij b_j  -- the reduction every pairwise-distance backward ends in:
//                              mode PLAIN   W = g                               (squared Euclidean / Frobenius distances)
//                              mode SPHERE  W = -g / sin d, 0 where the reference's clamp is active
//                                           (dd/dx_i of acos(clamp(<x_i, y_j>)), sphere_utils_torch.py:49-55)
//                              g (and d) are n1 x n2 row-major, or read transposed for the gradient of the SECOND operand.
//   gabo_spd_logm_backward   : adjoint of the Frechet derivative of the matrix logarithm at an SPD matrix
//                              (Daleckii-Krein: V [(V^T G V) o L] V^T, L_ab = (log l_a - log l_b) / (l_a - l_b)).
//   gabo_nested_spd_project_backward : Y = W^T X W  ->  dX_n = W G_n W^T,  dW = 2 sum_n X_n W G_n  (G symmetric).
#include "spd_common.cuh"

namespace gabo {
namespace {

constexpr int kWTile = 32;

// grid.x tiles of 32 output rows; block 256 threads.  W tiles go through shared memory so that g is read along its
// contiguous dimension in both orientations.
template <int MODE>
__global__ void __launch_bounds__(256)
    weighted_points_sum_kernel(const double* __restrict__ g, const double* __restrict__ dist, int64_t n1, int64_t n2,
                               int64_t ld, int transpose, const double* __restrict__ b, int k, double* __restrict__ out) {
    extern __shared__ double sm[];
    double* wt = sm;                       // [32][33]  W(i, j) for the current (row tile, j tile)
    double* bt = sm + kWTile * (kWTile + 1);  // [32][k]   b_j
    // logical sizes: rows = n1 (or n2 when transposed), cols = the other one
    const int64_t rows = transpose ? n2 : n1, cols = transpose ? n1 : n2;
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kWTile;
    const int tid = threadIdx.x;
    const int r = tid >> 3, c0 = tid & 7;                       // 32 rows x 8 column lanes
    constexpr int kMaxAcc = 16;                                  // k <= 128
    double acc[kMaxAcc];
#pragma unroll
    for (int q = 0; q < kMaxAcc; ++q) acc[q] = 0.0;
    const double lo = 4.4703483581542969e-08 * (1.0 + 1e-9);    // acos(1 - 1e-15): the clamp is active below this
    for (int64_t j0 = 0; j0 < cols; j0 += kWTile) {
        __syncthreads();
        // load the 32 x 32 weight tile: thread (a, bb) covers 4 entries; contiguous along the fast index of g
        for (int e = tid; e < kWTile * kWTile; e += 256) {
            const int fast = e & 31, slow = e >> 5;
            // memory index (mi, mj) of g: not transposed -> (i0 + slow, j0 + fast); transposed -> (j0 + slow, i0 + fast)
            const int64_t mi = transpose ? (j0 + slow) : (i0 + slow);
            const int64_t mj = transpose ? (i0 + fast) : (j0 + fast);
            double w = 0.0;
            if (mi < n1 && mj < n2) {
                w = g[mi * ld + mj];
                if (MODE == 1) {
                    const double dv = dist[mi * ld + mj];
                    const bool live = (dv > lo) && (dv < 3.14159265358979323846 - lo);
                    w = live ? -w / sin(dv) : 0.0;
                }
            }
            const int li = transpose ? fast : slow, lj = transpose ? slow : fast;   // logical (row, col) inside the tile
            wt[li * (kWTile + 1) + lj] = w;
        }
        for (int e = tid; e < kWTile * k; e += 256) {
            const int64_t j = j0 + e / k;
            bt[e] = (j < cols) ? b[j * k + (e % k)] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int jj = 0; jj < kWTile; ++jj) {
            const double w = wt[r * (kWTile + 1) + jj];
#pragma unroll
            for (int q = 0; q < kMaxAcc; ++q) {
                const int c = c0 + 8 * q;
                if (c < k) acc[q] = fma(w, bt[jj * k + c], acc[q]);
            }
        }
    }
    if (i0 + r < rows) {
#pragma unroll
        for (int q = 0; q < kMaxAcc; ++q) {
            const int c = c0 + 8 * q;
            if (c < k) out[(i0 + r) * k + c] = acc[q];
        }
    }
}

template <int d>
__global__ void spd_logm_backward_kernel(const double* __restrict__ mat, const double* __restrict__ gout, int64_t n,
                                         double* __restrict__ gin) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* m = mat + i * d * d;
    const double* go = gout + i * d * d;
    double S[d][d], V[d][d], lam[d];
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) S[r][c] = m[(r <= c ? r : c) * d + (r <= c ? c : r)];   // upper triangle, as symeig
    jacobi_symmetric<d, double, true>(S, lam, V);
    // T = V^T sym(G) V
    double T[d][d];
#pragma unroll
    for (int a = 0; a < d; ++a)
#pragma unroll
        for (int b2 = 0; b2 < d; ++b2) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < d; ++r) {
                double t = 0.0;
#pragma unroll
                for (int c = 0; c < d; ++c) t = fma(0.5 * (go[r * d + c] + go[c * d + r]), V[c][b2], t);
                s = fma(V[r][a], t, s);
            }
            T[a][b2] = s;
        }
    // Daleckii-Krein divided differences of log
#pragma unroll
    for (int a = 0; a < d; ++a)
#pragma unroll
        for (int b2 = 0; b2 < d; ++b2) {
            const double la = lam[a], lb = lam[b2];
            const double df = la - lb;
            double L;
            if (fabs(df) <= 1e-9 * fmax(fabs(la), fabs(lb))) L = 2.0 / (la + lb);   // limit 1 / lambda
            else L = (log(la) - log(lb)) / df;
            T[a][b2] *= L;
        }
    double* o = gin + i * d * d;
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < d; ++a) {
                double t = 0.0;
#pragma unroll
                for (int b2 = 0; b2 < d; ++b2) t = fma(T[a][b2], V[c][b2], t);
                s = fma(V[r][a], t, s);
            }
            o[r * d + c] = s;
        }
}

// One CTA per data matrix n: T = X_n W (D x d) in shared memory, then the contribution 2 T G_n to dW (atomicAdd, D d
// values) and, optionally, dX_n = W G_n W^T.  D <= 32, d <= 8, n is n_train-sized.
__global__ void __launch_bounds__(256)
    nested_project_backward_kernel(const double* __restrict__ x, const double* __restrict__ w,
                                   const double* __restrict__ gy, int64_t n, int D, int d, double* __restrict__ gx,
                                   double* __restrict__ gw) {
    __shared__ double ws[32 * 8], gs[8 * 8], ts[32 * 8];
    const int64_t i = blockIdx.x;
    const int tid = threadIdx.x;
    for (int e = tid; e < D * d; e += blockDim.x) ws[e] = w[e];
    for (int e = tid; e < d * d; e += blockDim.x) {
        const int r = e / d, c = e % d;
        gs[e] = 0.5 * (gy[i * d * d + r * d + c] + gy[i * d * d + c * d + r]);
    }
    __syncthreads();
    if (gw != nullptr) {
        const double* xm = x + i * D * D;
        for (int e = tid; e < D * d; e += blockDim.x) {         // T = sym(X) W
            const int r = e / d, c = e % d;
            double s = 0.0;
            for (int q = 0; q < D; ++q) s = fma(0.5 * (xm[r * D + q] + xm[q * D + r]), ws[q * d + c], s);
            ts[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < D * d; e += blockDim.x) {         // dW += 2 T G
            const int r = e / d, c = e % d;
            double s = 0.0;
            for (int q = 0; q < d; ++q) s = fma(ts[r * d + q], gs[q * d + c], s);
            atomicAdd(gw + e, 2.0 * s);
        }
    }
    if (gx != nullptr) {
        __syncthreads();
        for (int e = tid; e < D * d; e += blockDim.x) {         // T = W G
            const int r = e / d, c = e % d;
            double s = 0.0;
            for (int q = 0; q < d; ++q) s = fma(ws[r * d + q], gs[q * d + c], s);
            ts[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < D * D; e += blockDim.x) {         // dX = T W^T
            const int r = e / D, c = e % D;
            double s = 0.0;
            for (int q = 0; q < d; ++q) s = fma(ts[r * d + q], ws[c * d + q], s);
            gx[i * D * D + e] = s;
        }
    }
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_weighted_points_sum(const double* g, const double* dist, int64_t n1, int64_t n2, int64_t ld,
                                        int transpose, int mode, const double* b, int k, double* out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n1 >= 0 && n2 >= 0, GABO_E_ARG, "gabo_weighted_points_sum: negative size");
    const int64_t rows = transpose ? n2 : n1;
    if (rows == 0) return GABO_OK;
    GABO_REQUIRE(g && b && out, GABO_E_ARG, "gabo_weighted_points_sum: null pointer");
    GABO_REQUIRE(mode == 0 || (mode == 1 && dist), GABO_E_ARG, "gabo_weighted_points_sum: bad mode / missing distances");
    GABO_REQUIRE(k >= 1 && k <= 128, GABO_E_ARG, "gabo_weighted_points_sum: k=%d outside [1, 128]", k);
    GABO_REQUIRE(ld >= n2, GABO_E_ARG, "gabo_weighted_points_sum: ld < n2");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((rows + kWTile - 1) / kWTile);
    const size_t smem = sizeof(double) * (kWTile * (kWTile + 1) + kWTile * k);
    if (mode == 0) {
        weighted_points_sum_kernel<0><<<grid, 256, smem, s>>>(g, dist, n1, n2, ld, transpose, b, k, out);
    } else {
        weighted_points_sum_kernel<1><<<grid, 256, smem, s>>>(g, dist, n1, n2, ld, transpose, b, k, out);
    }
    return check_launch("weighted_points_sum_kernel");
}

extern "C" int gabo_spd_logm_backward(const double* mat, const double* grad_out, int64_t n, int d, double* grad_in,
                                      void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_spd_logm_backward: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(mat && grad_out && grad_in, GABO_E_ARG, "gabo_spd_logm_backward: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_logm_backward: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((n + 63) / 64);
    switch (d) {
#define GABO_CASE(DD)                                                                   \
    case DD:                                                                            \
        spd_logm_backward_kernel<DD><<<grid, 64, 0, s>>>(mat, grad_out, n, grad_in);    \
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
    return check_launch("spd_logm_backward_kernel");
}

extern "C" int gabo_nested_spd_project_backward(const double* x, const double* w, const double* grad_y, int64_t n, int D,
                                                int d, double* grad_x, double* grad_w, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_spd_project_backward: negative size");
    GABO_REQUIRE(D >= 1 && D <= 32 && d >= 1 && d <= 8 && d <= D, GABO_E_ARG,
                 "gabo_nested_spd_project_backward: sizes D=%d (<= 32), d=%d (<= 8)", D, d);
    GABO_REQUIRE(w && grad_y && (grad_x || grad_w), GABO_E_ARG, "gabo_nested_spd_project_backward: null pointer");
    GABO_REQUIRE(grad_w == nullptr || x != nullptr, GABO_E_ARG, "gabo_nested_spd_project_backward: grad_w needs x");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (grad_w) {
        const cudaError_t e = cudaMemsetAsync(grad_w, 0, sizeof(double) * D * d, s);
        GABO_REQUIRE(e == cudaSuccess, GABO_E_CUDA, "gabo_nested_spd_project_backward: memset failed");
    }
    if (n == 0) return GABO_OK;
    nested_project_backward_kernel<<<static_cast<unsigned>(n), 256, 0, s>>>(x, w, grad_y, n, D, d, grad_x, grad_w);
    return check_launch("nested_project_backward_kernel");
}
