// Batched manifold operations (M1, M2 of SURVEY.md section 8): one thread per point, fp64, elementwise over the batch.
//
// Sphere (pymanopt 0.2.x Sphere; the reference's own numpy statements are Riemannian_utils/sphere_utils.py:14-123):
//   proj / egrad2rgrad, retr, exp, log, projection transport, true parallel transport, dist.
// SPD under the affine-invariant metric (pymanopt 0.2.x PositiveDefinite; Riemannian_utils/spd_utils.py:104-213):
//   exp (= retr), log, egrad2rgrad, identity transport, true parallel transport E U E^T with E = (Y X^-1)^(1/2),
//   dist, norm, inner.  Everything goes through the Cholesky whitening X = L L^T, A = L^-1, so that only SYMMETRIC
//   eigen-problems are solved (Jacobi in registers), never the non-symmetric X^-1 Y of the numpy reference.
// The reference calls these one point at a time from Python inside the solvers (manifold_optimize.py:207-221).
#include "spd_common.cuh"

namespace gabo {
namespace {

// ------------------------------------------------------------------------------------------------------------
// sphere
// ------------------------------------------------------------------------------------------------------------
__global__ void sphere_op_kernel(int op, const double* __restrict__ a, const double* __restrict__ b,
                                 const double* __restrict__ c, int64_t n, int dim, double* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* x = a + i * dim;
    const double* y = b + i * dim;
    double* o = out + i * dim;
    switch (op) {
        case GABO_OP_PROJ:
        case GABO_OP_TRANSP: {  // out = y - <x,y> x
            double s = 0.0;
            for (int k = 0; k < dim; ++k) s = fma(x[k], y[k], s);
            for (int k = 0; k < dim; ++k) o[k] = y[k] - s * x[k];
            break;
        }
        case GABO_OP_RETR: {
            double s = 0.0;
            for (int k = 0; k < dim; ++k) {
                const double t = x[k] + y[k];
                s = fma(t, t, s);
            }
            const double inv = 1.0 / sqrt(s);
            for (int k = 0; k < dim; ++k) o[k] = (x[k] + y[k]) * inv;
            break;
        }
        case GABO_OP_EXP: {
            double s = 0.0;
            for (int k = 0; k < dim; ++k) s = fma(y[k], y[k], s);
            const double nu = sqrt(s);
            if (nu > 1e-3) {  // pymanopt Sphere.exp; same closed form as sphere_utils.py:33-36
                const double cs = cos(nu), sc = sin(nu) / nu;
                for (int k = 0; k < dim; ++k) o[k] = x[k] * cs + y[k] * sc;
            } else {  // retraction for tiny steps
                double q = 0.0;
                for (int k = 0; k < dim; ++k) {
                    const double t = x[k] + y[k];
                    q = fma(t, t, q);
                }
                const double inv = 1.0 / sqrt(q);
                for (int k = 0; k < dim; ++k) o[k] = (x[k] + y[k]) * inv;
            }
            break;
        }
        case GABO_OP_LOG: {  // proj(x, y - x) rescaled to length dist(x, y); sphere_utils.py:60-63
            double s = 0.0, cxy = 0.0;
            for (int k = 0; k < dim; ++k) {
                s = fma(x[k], y[k] - x[k], s);
                cxy = fma(x[k], y[k], cxy);
            }
            double pn = 0.0;
            for (int k = 0; k < dim; ++k) {
                const double p = (y[k] - x[k]) - s * x[k];
                pn = fma(p, p, pn);
            }
            pn = sqrt(pn);
            const double dist = acos(fmin(fmax(cxy, -1.0), 1.0));
            const double scale = (dist > 1e-6) ? dist / (pn > 0.0 ? pn : 1.0) : 1.0;
            for (int k = 0; k < dim; ++k) o[k] = ((y[k] - x[k]) - s * x[k]) * scale;
            break;
        }
        case GABO_OP_PTRANSP: {  // great-circle parallel transport x -> y of u: sphere_utils.py:116-121 applied to u
            const double* u = c + i * dim;
            double s = 0.0, cxy = 0.0;
            for (int k = 0; k < dim; ++k) {
                s = fma(x[k], y[k] - x[k], s);
                cxy = fma(x[k], y[k], cxy);
            }
            double pn = 0.0, pu = 0.0;
            for (int k = 0; k < dim; ++k) {
                const double p = (y[k] - x[k]) - s * x[k];
                pn = fma(p, p, pn);
                pu = fma(p, u[k], pu);
            }
            pn = sqrt(pn);
            const double dist = acos(fmin(fmax(cxy, -1.0), 1.0));
            const double nrm = (dist > 1e-6) ? dist : pn;  // |Log_x(y)|
            if (nrm < 1e-16 || pn == 0.0) {
                for (int k = 0; k < dim; ++k) o[k] = u[k];
            } else {
                const double ev = pu / pn;  // <e, u>, e = Log_x(y) / |Log_x(y)|
                const double sn = sin(nrm), cs = cos(nrm);
                for (int k = 0; k < dim; ++k) {
                    const double e = ((y[k] - x[k]) - s * x[k]) / pn;
                    o[k] = -x[k] * sn * ev + e * cs * ev + u[k] - e * ev;
                }
            }
            break;
        }
        default:
            break;
    }
}

__global__ void sphere_dist_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, int dim,
                                   double* __restrict__ out) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c = 0.0;
    for (int k = 0; k < dim; ++k) c = fma(x[i * dim + k], y[i * dim + k], c);
    out[i] = acos(fmin(fmax(c, -1.0), 1.0));  // pymanopt Sphere.dist; sphere_utils.py:68-90
}

// ------------------------------------------------------------------------------------------------------------
// SPD
// ------------------------------------------------------------------------------------------------------------
template <int d>
__device__ __forceinline__ void load_sym(const double* m, double (&M)[d][d]) {
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) M[r][c] = 0.5 * (m[r * d + c] + m[c * d + r]);
}

template <int d>
__device__ __forceinline__ void store_mat(double* o, const double (&M)[d][d], bool ok) {
    const double nanv = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) o[r * d + c] = M[r][c] + nanv;
}

// W = A M A^T for packed lower-triangular A and a full matrix M
template <int d>
__device__ __forceinline__ void whiten(const double (&A)[tri_size(d)], const double (&M)[d][d], double (&W)[d][d]) {
    double Af[d][d], T[d][d];
    tri_expand<d, double>([&](int e) { return A[e]; }, Af);
    matmul<d, double>(Af, M, T);
    matmul_nt<d, double>(T, Af, W);
}

template <int d>
__global__ void __launch_bounds__(64) spd_op_kernel(int op, const double* __restrict__ a, const double* __restrict__ b,
                                                    const double* __restrict__ c, int64_t n, double* __restrict__ out) {
    constexpr int TRI = tri_size(d);
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* xa = a + i * d * d;
    const double* xb = b + i * d * d;
    double* o = out + i * d * d;
    if (op == GABO_OP_TRANSP) {  // identity transport (pymanopt 0.2.x PositiveDefinite.transp)
#pragma unroll
        for (int e = 0; e < d * d; ++e) o[e] = xb[e];
        return;
    }
    if (op == GABO_OP_EGRAD2RGRAD || op == GABO_OP_PROJ) {
        double G[d][d];
        load_sym<d>(xb, G);
        if (op == GABO_OP_PROJ) {  // proj = sym
            store_mat<d>(o, G, true);
            return;
        }
        double X[d][d], T[d][d], R[d][d];
#pragma unroll
        for (int r = 0; r < d; ++r)
#pragma unroll
            for (int cc = 0; cc < d; ++cc) X[r][cc] = xa[r * d + cc];
        matmul<d, double>(X, G, T);
        matmul<d, double>(T, X, R);
        store_mat<d>(o, R, true);
        return;
    }
    double L[TRI], A[TRI];
    bool ok = chol_inv<d>([&](int r, int cc) { return xa[r * d + cc]; }, L, A);
    double Lf[d][d];
    tri_expand<d, double>([&](int e) { return L[e]; }, Lf);
    if (op == GABO_OP_EXP || op == GABO_OP_RETR) {
        double U[d][d], S[d][d], lam[d], V[d][d];
        load_sym<d>(xb, U);
        whiten<d>(A, U, S);
        jacobi_symmetric<d, double, true>(S, lam, V);
        double B[d][d], f[d], R[d][d];
        matmul<d, double>(Lf, V, B);
#pragma unroll
        for (int k = 0; k < d; ++k) f[k] = exp(lam[k]);
        weighted_outer<d, double>(B, f, R);
        store_mat<d>(o, R, ok);
        return;
    }
    // LOG and PTRANSP need the second point: G = A L_y, one-sided Jacobi
    double Ly[TRI], Ay[TRI];
    ok = chol_inv<d>([&](int r, int cc) { return xb[r * d + cc]; }, Ly, Ay) && ok;
    double G[d][d], lam[d];
    tri_product<d, double>([&](int e) { return A[e]; }, [&](int e) { return Ly[e]; }, G);
    jacobi_onesided<d, double>(G, lam);
    double f[d], C[d][d], T[d][d], R[d][d];
    if (op == GABO_OP_LOG) {
#pragma unroll
        for (int k = 0; k < d; ++k) f[k] = log(lam[k]) / lam[k];
        weighted_outer<d, double>(G, f, C);
        matmul<d, double>(Lf, C, T);
        matmul_nt<d, double>(T, Lf, R);
        store_mat<d>(o, R, ok);
        return;
    }
    // PTRANSP: E = L W^(1/2) A, out = E U E^T
#pragma unroll
    for (int k = 0; k < d; ++k) f[k] = rsqrt(lam[k]);
    weighted_outer<d, double>(G, f, C);
    double Af[d][d], E[d][d], U[d][d];
    tri_expand<d, double>([&](int e) { return A[e]; }, Af);
    matmul<d, double>(Lf, C, T);
    matmul<d, double>(T, Af, E);
    const double* xc = c + i * d * d;
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int cc = 0; cc < d; ++cc) U[r][cc] = xc[r * d + cc];
    matmul<d, double>(E, U, T);
    matmul_nt<d, double>(T, E, R);
    store_mat<d>(o, R, ok);
}

template <int d>
__global__ void __launch_bounds__(64)
    spd_scalar_kernel(int what, const double* __restrict__ x, const double* __restrict__ b, const double* __restrict__ c,
                      int64_t n, double* __restrict__ out) {
    constexpr int TRI = tri_size(d);
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* xa = x + i * d * d;
    const double* xb = b + i * d * d;
    double L[TRI], A[TRI];
    bool ok = chol_inv<d>([&](int r, int cc) { return xa[r * d + cc]; }, L, A);
    double res;
    if (what == 0) {  // dist
        double Ly[TRI], Ay[TRI];
        ok = chol_inv<d>([&](int r, int cc) { return xb[r * d + cc]; }, Ly, Ay) && ok;
        double G[d][d], lam[d];
        tri_product<d, double>([&](int e) { return A[e]; }, [&](int e) { return Ly[e]; }, G);
        jacobi_onesided<d, double>(G, lam);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < d; ++k) {
            const double l = log(lam[k]);
            s = fma(l, l, s);
        }
        res = sqrt(s);
    } else {
        double U[d][d], WU[d][d];
        load_sym<d>(xb, U);
        whiten<d>(A, U, WU);
        double s = 0.0;
        if (what == 1) {
#pragma unroll
            for (int r = 0; r < d; ++r)
#pragma unroll
                for (int cc = 0; cc < d; ++cc) s = fma(WU[r][cc], WU[r][cc], s);
            res = sqrt(s);
        } else {
            double V[d][d], WV[d][d];
            load_sym<d>(c + i * d * d, V);
            whiten<d>(A, V, WV);
#pragma unroll
            for (int r = 0; r < d; ++r)
#pragma unroll
                for (int cc = 0; cc < d; ++cc) s = fma(WU[r][cc], WV[r][cc], s);
            res = s;
        }
    }
    out[i] = ok ? res : __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_sphere_op(int op, const double* a, const double* b, const double* c, int64_t n, int dim,
                              double* out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_sphere_op: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(a && b && out, GABO_E_ARG, "gabo_sphere_op: null pointer");
    GABO_REQUIRE(dim >= 1 && dim <= 4096, GABO_E_ARG, "gabo_sphere_op: bad dim %d", dim);
    GABO_REQUIRE(op >= GABO_OP_PROJ && op <= GABO_OP_PTRANSP, GABO_E_ARG, "gabo_sphere_op: bad op %d", op);
    GABO_REQUIRE(op != GABO_OP_PTRANSP || c, GABO_E_ARG, "gabo_sphere_op: PTRANSP needs the vector in c");
    const unsigned grid = static_cast<unsigned>((n + 127) / 128);
    sphere_op_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(op, a, b, c, n, dim, out);
    return check_launch("sphere_op_kernel");
}

extern "C" int gabo_sphere_dist(const double* x, const double* y, int64_t n, int dim, double* out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_sphere_dist: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x && y && out, GABO_E_ARG, "gabo_sphere_dist: null pointer");
    GABO_REQUIRE(dim >= 1 && dim <= 4096, GABO_E_ARG, "gabo_sphere_dist: bad dim %d", dim);
    const unsigned grid = static_cast<unsigned>((n + 127) / 128);
    sphere_dist_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n, dim, out);
    return check_launch("sphere_dist_kernel");
}

extern "C" int gabo_spd_op(int op, const double* a, const double* b, const double* c, int64_t n, int d, double* out,
                           void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_spd_op: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(a && b && out, GABO_E_ARG, "gabo_spd_op: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_op: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    GABO_REQUIRE(op >= GABO_OP_PROJ && op <= GABO_OP_EGRAD2RGRAD, GABO_E_ARG, "gabo_spd_op: bad op %d", op);
    GABO_REQUIRE(op != GABO_OP_PTRANSP || c, GABO_E_ARG, "gabo_spd_op: PTRANSP needs the vector in c");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((n + 63) / 64);
    switch (d) {
#define GABO_CASE(DD)                                                \
    case DD:                                                         \
        spd_op_kernel<DD><<<grid, 64, 0, s>>>(op, a, b, c, n, out);  \
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
    return check_launch("spd_op_kernel");
}

extern "C" int gabo_spd_scalar(int what, const double* x, const double* b, const double* c, int64_t n, int d,
                               double* out, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_spd_scalar: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x && b && out, GABO_E_ARG, "gabo_spd_scalar: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_scalar: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    GABO_REQUIRE(what >= 0 && what <= 2, GABO_E_ARG, "gabo_spd_scalar: bad selector %d", what);
    GABO_REQUIRE(what != 2 || c, GABO_E_ARG, "gabo_spd_scalar: inner needs the second vector in c");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((n + 63) / 64);
    switch (d) {
#define GABO_CASE(DD)                                                     \
    case DD:                                                              \
        spd_scalar_kernel<DD><<<grid, 64, 0, s>>>(what, x, b, c, n, out); \
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
    return check_launch("spd_scalar_kernel");
}


// ---------------------------------------------------------------------------------------------------------------
// Nested-sphere projection chain of HD-GaBO on spheres: S^{D-1} -> S^{D-2} -> ... -> S^{dl-1}.
// Replaces projection_from_sphere_to_subsphere (BoManifolds/nested_mappings/nested_spheres_utils.py:120-147: per level
// a rotation matrix from rotation_from_sphere_points_torch, sphere_utils_torch.py:58-93, two dense matrix products,
// a projection onto the nested sphere and the identification with the next subsphere, :13-118).
// One thread per point, every level in registers/local memory.  The rotation R that moves the level's axis v to the
// north pole e is never formed: with c = <v, e>, s = sqrt(1 - c^2) and u = (v - c e) / |v - c e|,
//     R x = x + (s a + (c - 1) b) e + ((c - 1) a - s b) u,      a = <u, x>,  b = <e, x> = x_k,
// which is O(k) per point instead of O(k^2).  R^T is only used by the reference to go back to the nested sphere and
// forth again (R R^T = I), so the subsphere coordinates are the first k-1 rotated ones.
// ---------------------------------------------------------------------------------------------------------------
namespace gabo {
namespace {

constexpr int kMaxNestedDim = 64;

__global__ void nested_sphere_project_kernel(const double* __restrict__ x, int64_t n, int D, int dl,
                                             const double* __restrict__ axes, const double* __restrict__ dists,
                                             double* __restrict__ y) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[kMaxNestedDim];
    for (int k = 0; k < D; ++k) p[k] = x[i * D + k];
    const double* v = axes;
    for (int k = D, lvl = 0; k > dl; --k, ++lvl) {
        // rotation parameters of this level (identical for every thread; O(k))
        double c = fmin(fmax(v[k - 1], -1.0 + 1e-15), 1.0 - 1e-15);      // <v, e>, sphere_utils_torch.py:80-83
        double un = 0.0;
        for (int q = 0; q < k; ++q) {
            const double t = v[q] - ((q == k - 1) ? c : 0.0);
            un = fma(t, t, un);
        }
        const double uinv = 1.0 / sqrt(un);
        const double s = sin(acos(c));
        double a = 0.0;
        for (int q = 0; q < k; ++q) a = fma((v[q] - ((q == k - 1) ? c : 0.0)) * uinv, p[q], a);
        const double b = p[k - 1];
        const double ce = s * a + (c - 1.0) * b, cu = (c - 1.0) * a - s * b;
        for (int q = 0; q < k; ++q) p[q] = fma(cu, (v[q] - ((q == k - 1) ? c : 0.0)) * uinv, p[q]);
        p[k - 1] += ce;                                                   // p = R x
        // distance to the axis, projection onto the nested sphere (nested_spheres_utils.py:50-59)
        const double r = dists[lvl];
        const double da = acos(fmin(fmax(p[k - 1], -1.0 + 1e-15), 1.0 - 1e-15));
        const double sr = sin(r), inv_sd = 1.0 / (sin(da) + 1e-6);
        // identification with the subsphere (:104-112): first k-1 coordinates / (sin r + 1e-6), renormalised
        const double inv_sr = 1.0 / (sr + 1e-6);
        double nn = 0.0;
        for (int q = 0; q < k - 1; ++q) {
            p[q] = (sr * p[q]) * inv_sd * inv_sr;
            nn = fma(p[q], p[q], nn);
        }
        const double inv_n = 1.0 / (sqrt(nn) + 1e-6);
        for (int q = 0; q < k - 1; ++q) p[q] *= inv_n;
        v += k;
    }
    for (int k = 0; k < dl; ++k) y[i * dl + k] = p[k];
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_nested_sphere_project(const double* x, int64_t n, int D, int d_latent, const double* axes,
                                          const double* dists, double* y, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_sphere_project: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x && y && (D == d_latent || (axes && dists)), GABO_E_ARG, "gabo_nested_sphere_project: null pointer");
    GABO_REQUIRE(D >= 2 && D <= kMaxNestedDim && d_latent >= 1 && d_latent <= D, GABO_E_ARG,
                 "gabo_nested_sphere_project: need 1 <= d_latent <= D <= %d, got D=%d d_latent=%d", kMaxNestedDim, D,
                 d_latent);
    nested_sphere_project_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, D, d_latent, axes, dists, y);
    return check_launch("nested_sphere_project_kernel");
}
