// Nested-manifold maps either side of the latent acquisition optimisation of HD-GaBO:
//   gabo_nested_sphere_chain        : every level of projection_from_sphere_to_subsphere
//                                     (BoManifolds/nested_mappings/nested_spheres_utils.py:120-147)
//   gabo_nested_sphere_to_nested    : projection_from_sphere_to_nested_sphere (:13-67)
//   gabo_nested_sphere_reconstruct  : every level of projection_from_subsphere_to_sphere (:149-213)
//   gabo_spd_sqrtm                  : sqrtm_torch for a batch (Riemannian_utils/spd_utils_torch.py:33-50)
//   gabo_nested_spd_reconstruct_*   : projection_from_nested_spd_to_spd (nested_mappings/nested_spd_utils.py:51-118)
// All fp64 (the reference runs these in fp64 on a handful of points; here they are batched, one thread or one CTA
// per point, with every rotation applied in O(k) without forming the k x k matrix).
#include <cstdlib>

#include "common.cuh"
#include "spd_common.cuh"

namespace gabo {
namespace {

constexpr int kMaxNestedDim = 64;

// The reference's rotation_from_sphere_points_torch(x, y) (sphere_utils_torch.py:58-93) moves x to y:
//     R = I + sin(theta) (y (x) u - u (x) y) + (c - 1) (y (x) y + u (x) u),  c = clamp(<x, y>), theta = acos c,
//     u = (x - c y) / |x - c y|.
// Applied to p:  R p = p + (s a + (c-1) b) y + ((c-1) a - s b) u,   a = <u, p>, b = <y, p>;  R^T flips the sign of s.
// Here one of x, y is always the north pole e = (0, ..., 0, 1), so c = clamp(v[k-1]) for the level's axis v.
struct LevelRotation {
    double c, s, uinv;
};

__device__ __forceinline__ LevelRotation level_rotation(const double* __restrict__ v, int k, bool to_north) {
    LevelRotation r;
    r.c = fmin(fmax(v[k - 1], -1.0 + 1e-15), 1.0 - 1e-15);            // sphere_utils_torch.py:80-83
    double un = 0.0;
    if (to_north) {                                                    // u ~ v - c e
        for (int q = 0; q < k; ++q) {
            const double t = v[q] - ((q == k - 1) ? r.c : 0.0);
            un = fma(t, t, un);
        }
    } else {                                                           // u ~ e - c v
        for (int q = 0; q < k; ++q) {
            const double t = ((q == k - 1) ? 1.0 : 0.0) - r.c * v[q];
            un = fma(t, t, un);
        }
    }
    r.uinv = 1.0 / sqrt(un);
    r.s = sin(acos(r.c));
    return r;
}

// p <- R p (sign = +1) or R^T p (sign = -1) for the rotation axis -> north pole.
__device__ __forceinline__ void rotate_to_north(double* p, const double* __restrict__ v, int k, const LevelRotation& r,
                                                double sign) {
    double a = 0.0;
    for (int q = 0; q < k; ++q) a = fma((v[q] - ((q == k - 1) ? r.c : 0.0)) * r.uinv, p[q], a);
    const double b = p[k - 1];
    const double s = sign * r.s;
    const double ce = s * a + (r.c - 1.0) * b, cu = (r.c - 1.0) * a - s * b;
    for (int q = 0; q < k; ++q) p[q] = fma(cu, (v[q] - ((q == k - 1) ? r.c : 0.0)) * r.uinv, p[q]);
    p[k - 1] += ce;
}

// p <- R p for the rotation north pole -> axis.
__device__ __forceinline__ void rotate_from_north(double* p, const double* __restrict__ v, int k,
                                                  const LevelRotation& r) {
    double a = 0.0, b = 0.0;
    for (int q = 0; q < k; ++q) {
        a = fma((((q == k - 1) ? 1.0 : 0.0) - r.c * v[q]) * r.uinv, p[q], a);
        b = fma(v[q], p[q], b);
    }
    const double cy = r.s * a + (r.c - 1.0) * b, cu = (r.c - 1.0) * a - r.s * b;
    for (int q = 0; q < k; ++q) {
        const double u = (((q == k - 1) ? 1.0 : 0.0) - r.c * v[q]) * r.uinv;
        p[q] = fma(cy, v[q], fma(cu, u, p[q]));
    }
}

// One level down, in place: p (k values on S^{k-1}) -> k-1 values on S^{k-2} (nested_spheres_utils.py:44-59, 97-112).
__device__ __forceinline__ void level_down(double* p, const double* __restrict__ v, int k, double r) {
    const LevelRotation rot = level_rotation(v, k, true);
    rotate_to_north(p, v, k, rot, 1.0);
    const double da = acos(fmin(fmax(p[k - 1], -1.0 + 1e-15), 1.0 - 1e-15));
    const double sr = sin(r), inv_sd = 1.0 / (sin(da) + 1e-6), inv_sr = 1.0 / (sr + 1e-6);
    double nn = 0.0;
    for (int q = 0; q < k - 1; ++q) {
        p[q] = (sr * p[q]) * inv_sd * inv_sr;
        nn = fma(p[q], p[q], nn);
    }
    const double inv_n = 1.0 / (sqrt(nn) + 1e-6);
    for (int q = 0; q < k - 1; ++q) p[q] *= inv_n;
}

__global__ void nested_sphere_chain_kernel(const double* __restrict__ x, int64_t n, int D, int dl,
                                           const double* __restrict__ axes, const double* __restrict__ dists,
                                           double* __restrict__ levels) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[kMaxNestedDim];
    for (int k = 0; k < D; ++k) p[k] = x[i * D + k];
    const double* v = axes;
    double* out = levels;
    for (int k = D, lvl = 0; k > dl; --k, ++lvl) {
        level_down(p, v, k, dists[lvl]);
        for (int q = 0; q < k - 1; ++q) out[i * (k - 1) + q] = p[q];
        out += n * (k - 1);
        v += k;
    }
}

__global__ void nested_sphere_to_nested_kernel(const double* __restrict__ x, int64_t n, int k,
                                               const double* __restrict__ v, double r, double* __restrict__ y) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[kMaxNestedDim];
    for (int q = 0; q < k; ++q) p[q] = x[i * k + q];
    const LevelRotation rot = level_rotation(v, k, true);
    rotate_to_north(p, v, k, rot, 1.0);
    const double da = acos(fmin(fmax(p[k - 1], -1.0 + 1e-15), 1.0 - 1e-15));
    const double sr = sin(r), inv_sd = 1.0 / (sin(da) + 1e-6);
    for (int q = 0; q < k - 1; ++q) p[q] = (sr * p[q]) * inv_sd;
    p[k - 1] = fma(sr, p[k - 1], sin(da - r)) * inv_sd;
    rotate_to_north(p, v, k, rot, -1.0);                               // back: R^T
    for (int q = 0; q < k; ++q) y[i * k + q] = p[q];
}

// Up the chain: y (dl values) -> dl+1 -> ... -> D.  axes / dists are given in the order of the projection
// (levels D, D-1, ..., dl+1); the reconstruction consumes them last to first (nested_spheres_utils.py:207-211).
__global__ void nested_sphere_reconstruct_kernel(const double* __restrict__ y, int64_t n, int dl, int D,
                                                 const double* __restrict__ axes, const double* __restrict__ dists,
                                                 double* __restrict__ levels) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[kMaxNestedDim];
    for (int q = 0; q < dl; ++q) p[q] = y[i * dl + q];
    // offset of the axis of level k (k = D, D-1, ...): sum of the longer ones before it
    double* out = levels;
    for (int k = dl + 1; k <= D; ++k) {
        const int lvl = D - k;
        int off = 0;
        for (int kk = D; kk > k; --kk) off += kk;
        const double* v = axes + off;
        const double r = dists[lvl];
        const double sr = sin(r), cr = cos(r);
        for (int q = 0; q < k - 1; ++q) p[q] *= sr;                    // nested_spheres_utils.py:176-177
        p[k - 1] = cr;
        const LevelRotation rot = level_rotation(v, k, false);
        rotate_from_north(p, v, k, rot);
        for (int q = 0; q < k; ++q) out[i * k + q] = p[q];
        out += n * k;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// sqrtm for a batch of small SPD matrices: same route as gabo_spd_logm (Cholesky X = L L^T, one-sided Jacobi on L:
// L V = U Sigma, X = U Sigma^2 U^T, sqrtm X = sum_k g_k g_k^T / sigma_k with g_k = sigma_k u_k the columns of L V).
// ---------------------------------------------------------------------------------------------------------------
template <int d>
__global__ void spd_sqrtm_kernel(const double* __restrict__ mat, int64_t n, double* __restrict__ out) {
    constexpr int TRI = tri_size(d);
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* m = mat + i * d * d;
    double L[TRI], A[TRI];
    const bool ok = chol_inv<d>([&](int r, int c) { return m[c * d + r]; }, L, A);   // symeig(upper=True)
    double G[d][d];
    tri_expand<d, double>([&](int e) { return L[e]; }, G);
    double lam[d];
    jacobi_onesided<d, double>(G, lam);
    double f[d];
#pragma unroll
    for (int k = 0; k < d; ++k) f[k] = 1.0 / sqrt(lam[k]);
    double C[d][d];
    weighted_outer<d, double>(G, f, C);
    double* o = out + i * d * d;
    const double nanv = ok ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int r = 0; r < d; ++r)
#pragma unroll
        for (int c = 0; c < d; ++c) o[r * d + c] = C[r][c] + nanv;
}

// ---------------------------------------------------------------------------------------------------------------
// SPD reconstruction X = R [Y B; B^T C] R^T, R = [W V], B = Y^(1/2) K C^(1/2)  (nested_spd_utils.py:51-118), as
//     X = W Y W^T + E + E^T + Z,   E = W Y^(1/2) N,   N = K C^(1/2) V^T (d x D),   Z = V C V^T (D x D).
// N and Z do not depend on the point: the setup kernel computes them once (C^(1/2) by a warp-cooperative two-sided
// Jacobi in shared memory, any size up to 32), the batch kernel streams the points.
// pack = [W (D x d) | N (d x D) | Z (D x D) | P (K x D^2)] doubles (P: see the batch kernel).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxReconDim = 32;

__global__ void __launch_bounds__(32)
nested_spd_reconstruct_setup_kernel(const double* __restrict__ w, const double* __restrict__ v,
                                    const double* __restrict__ cmat, const double* __restrict__ kmat, int D, int d,
                                    double* __restrict__ pack, int* __restrict__ flag) {
    const int m = D - d;
    const int lane = threadIdx.x;
    __shared__ double S[kMaxReconDim * kMaxReconDim];   // working copy of C -> diagonal
    __shared__ double Q[kMaxReconDim * kMaxReconDim];   // eigenvectors (columns), then C^(1/2)
    __shared__ double M[8 * kMaxReconDim];              // K C^(1/2)  (d x m)
    __shared__ double lam[kMaxReconDim];
    double fro = 0.0;
    for (int e = lane; e < m * m; e += 32) {
        const int r = e / m, c = e % m;
        const double val = (r <= c) ? cmat[r * m + c] : cmat[c * m + r];   // symeig(upper=True)
        S[e] = val;
        Q[e] = (r == c) ? 1.0 : 0.0;
        fro = fma(val, val, fro);
    }
    fro = warp_sum(fro);
    const double thr = 1e-15 * sqrt(fro);
    __syncwarp();
    for (int sweep = 0; sweep < 30; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < m - 1; ++p) {
            for (int q = p + 1; q < m; ++q) {
                const double apq = S[p * m + q];
                if (fabs(apq) <= thr) continue;                           // uniform across the warp
                rotated = true;
                const double tau = (S[q * m + q] - S[p * m + p]) / (2.0 * apq);
                const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
                const double cs = 1.0 / sqrt(fma(t, t, 1.0)), sn = t * cs;
                __syncwarp();
                if (lane < m) {                                            // columns p, q
                    const double sp = S[lane * m + p], sq = S[lane * m + q];
                    S[lane * m + p] = fma(cs, sp, -sn * sq);
                    S[lane * m + q] = fma(sn, sp, cs * sq);
                    const double vp = Q[lane * m + p], vq = Q[lane * m + q];
                    Q[lane * m + p] = fma(cs, vp, -sn * vq);
                    Q[lane * m + q] = fma(sn, vp, cs * vq);
                }
                __syncwarp();
                if (lane < m) {                                            // rows p, q
                    const double sp = S[p * m + lane], sq = S[q * m + lane];
                    S[p * m + lane] = fma(cs, sp, -sn * sq);
                    S[q * m + lane] = fma(sn, sp, cs * sq);
                }
                __syncwarp();
            }
        }
        if (!rotated) break;
    }
    bool bad = false;
    if (lane < m) {
        const double l = S[lane * m + lane];
        bad = !(l > 0.0);
        lam[lane] = sqrt(l);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) *flag = 1;              // bottom block not positive definite
    __syncwarp();
    // C^(1/2) = Q diag(sqrt lam) Q^T into S
    for (int e = lane; e < m * m; e += 32) {
        const int r = e / m, c = e % m;
        double acc = 0.0;
        for (int k = 0; k < m; ++k) acc = fma(Q[r * m + k] * lam[k], Q[c * m + k], acc);
        S[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < d * m; e += 32) {                               // M = K C^(1/2)
        const int r = e / m, c = e % m;
        double acc = 0.0;
        for (int k = 0; k < m; ++k) acc = fma(kmat[r * m + k], S[k * m + c], acc);
        M[e] = acc;
    }
    __syncwarp();
    double* pw = pack;
    double* pn = pack + D * d;
    double* pz = pn + d * D;
    for (int e = lane; e < D * d; e += 32) pw[e] = w[e];
    for (int e = lane; e < d * D; e += 32) {                               // N = M V^T
        const int r = e / D, c = e % D;
        double acc = 0.0;
        for (int k = 0; k < m; ++k) acc = fma(M[r * m + k], v[c * m + k], acc);
        pn[e] = acc;
    }
    // Z = V C V^T: first T = V C into Q (D x m fits: D <= 32, m < 32), then Z
    for (int e = lane; e < D * m; e += 32) {
        const int r = e / m, c = e % m;
        double acc = 0.0;
        for (int k = 0; k < m; ++k) acc = fma(v[r * m + k], cmat[k * m + c], acc);
        Q[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < D * D; e += 32) {
        const int r = e / D, c = e % D;
        double acc = 0.0;
        for (int k = 0; k < m; ++k) acc = fma(Q[r * m + k], v[c * m + k], acc);
        pz[e] = acc;
    }
    __syncwarp();
    // P (K x D^2): coefficients of vec(Y) and of the upper triangle of S = Y^(1/2) in vec(X)
    double* pp = pz + D * D;
    const int dd = d * d, DD = D * D, K = dd + d * (d + 1) / 2;
    for (int e = lane; e < K * DD; e += 32) {
        const int k = e / DD, rc = e % DD, r = rc / D, c = rc % D;
        double val;
        if (k < dd) {
            val = pw[r * d + k / d] * pw[c * d + k % d];                       // W Y W^T
        } else {
            int rem = k - dd, a = 0;
            while (rem >= d - a) {
                rem -= d - a;
                ++a;
            }
            const int q = a + rem;                                              // S_aq, a <= q
            val = pw[r * d + a] * pn[q * D + c] + pw[c * d + a] * pn[q * D + r];    // W S N + (W S N)^T
            if (q != a) val += pw[r * d + q] * pn[a * D + c] + pw[c * d + q] * pn[a * D + r];
        }
        pp[e] = val;
    }
}

// The reconstruction is LINEAR in (Y, S = Y^(1/2)):  vec(X) = P^T u + vec(Z),  u = [vec(Y) (d^2) | S_aq, a <= q],
// so a batch is one (n x K) x (K x D^2) contraction, K = d^2 + d(d+1)/2 (40 for SPD(5) -> SPD(20)).  P is built once by
// the setup kernel and stays in L1/L2; a CTA takes 32 points at a time and stages their u in shared memory.  A work
// item is 2 output columns x 8 points (16 fp64 accumulators, <= 85 registers so that three CTAs share an SM): per k
// one 128-bit P load, 4 broadcast 128-bit shared loads and 16 DFMA.  (ncu on the earlier versions: 1 column x 32 points
// was bound by the shared-memory pipe, 2 columns x 16 points by occupancy -- 144 registers, one CTA per SM.)
constexpr int kReconPts = 32;
constexpr int kReconHalf = 8;      // points per work item
constexpr int kReconThreads = 256;

__global__ void __launch_bounds__(kReconThreads, 3)
nested_spd_reconstruct_kernel(const double* __restrict__ y, const double* __restrict__ sq, int64_t n, int D, int d,
                              const double* __restrict__ pack, double* __restrict__ x) {
    extern __shared__ __align__(16) double u[];      // K x kReconPts
    const int dd = d * d, DD = D * D, K = dd + d * (d + 1) / 2;
    const double* __restrict__ Z = pack + 2 * D * d;
    const double* __restrict__ P = Z + DD;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int ncp = (DD + 1) / 2, items = ncp * (kReconPts / kReconHalf);
    const bool even = (DD & 1) == 0 && ((reinterpret_cast<uintptr_t>(pack) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
    const int64_t tiles = (n + kReconPts - 1) / kReconPts;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t i0 = t * kReconPts;
        __syncthreads();
        for (int e = tid; e < K * kReconPts; e += nthr) {
            const int pt = e / K, k = e % K;         // consecutive threads read consecutive entries of one point
            const int64_t i = i0 + pt;
            double v = 0.0;
            if (i < n) {
                if (k < dd) {
                    v = y[i * dd + k];
                } else {                              // k - dd enumerates (a, q), a <= q, row-major
                    int rem = k - dd, a = 0;
                    while (rem >= d - a) {
                        rem -= d - a;
                        ++a;
                    }
                    v = sq[i * dd + a * d + a + rem];
                }
            }
            u[k * kReconPts + pt] = v;
        }
        __syncthreads();
        for (int it = tid; it < items; it += nthr) {
            const int cp = it % ncp, h = it / ncp;
            const int c0 = 2 * cp, c1 = (c0 + 1 < DD) ? c0 + 1 : c0;
            double a0[kReconHalf], a1[kReconHalf];
            const double z0 = Z[c0], z1 = Z[c1];
#pragma unroll
            for (int pt = 0; pt < kReconHalf; ++pt) {
                a0[pt] = z0;
                a1[pt] = z1;
            }
#pragma unroll 2
            for (int k = 0; k < K; ++k) {
                double p0, p1;
                if (even) {                                       // 16-byte aligned pair (pack offsets are even)
                    const double2 pv = *reinterpret_cast<const double2*>(P + k * DD + c0);
                    p0 = pv.x;
                    p1 = pv.y;
                } else {
                    p0 = P[k * DD + c0];
                    p1 = P[k * DD + c1];
                }
                const double2* uk = reinterpret_cast<const double2*>(u + k * kReconPts + h * kReconHalf);
#pragma unroll
                for (int q = 0; q < kReconHalf / 2; ++q) {
                    const double2 uv = uk[q];
                    a0[2 * q] = fma(uv.x, p0, a0[2 * q]);
                    a0[2 * q + 1] = fma(uv.y, p0, a0[2 * q + 1]);
                    a1[2 * q] = fma(uv.x, p1, a1[2 * q]);
                    a1[2 * q + 1] = fma(uv.y, p1, a1[2 * q + 1]);
                }
            }
            const int64_t ib = i0 + h * kReconHalf;
#pragma unroll
            for (int pt = 0; pt < kReconHalf; ++pt) {
                if (ib + pt < n) {
                    if (even) {
                        *reinterpret_cast<double2*>(x + (ib + pt) * DD + c0) = make_double2(a0[pt], a1[pt]);
                    } else {
                        x[(ib + pt) * DD + c0] = a0[pt];
                        if (c1 != c0) x[(ib + pt) * DD + c1] = a1[pt];
                    }
                }
            }
        }
    }
}

// Symmetric form with the operator resident in shared memory (used whenever it fits).  ncu on the kernel above (round 1):
// 20 % of HBM, fp64 pipe 36 % busy, long_scoreboard on the P loads -- the 128 KB operator does not fit L1 next to the
// shared-memory tiles and is re-read from L2 for every 32-point tile; and the contraction itself (2 D^2 K FLOP per point,
// fp64) needs 58 TFLOP/s to keep up with HBM, more than the chip has.  X is symmetric, so only the column pairs of the
// UPPER triangle are contracted (D (D+1)/2 entries instead of D^2: half the FLOP) and mirrored on the store, and that half
// of P (K x ~D(D+1)/2 doubles: 70 KB for SPD(5) -> SPD(20)) is staged ONCE per persistent CTA in shared memory.
// Work item = one pair of row-adjacent upper-triangle entries x 8 points; per k: one 16-byte P load and four 16-byte
// broadcast u loads from shared memory feed 16 DFMA.
constexpr int kReconSymThreads = 224;
constexpr int kReconLd = kReconPts + 2;   // padded row stride of the tile inputs (even: 16-byte aligned rows)
constexpr int kReconFill = 8;             // tile entries per thread held in registers (K * 32 <= 8 * 224)

__global__ void __launch_bounds__(kReconSymThreads, 2)
nested_spd_reconstruct_sym_kernel(const double* __restrict__ y, const double* __restrict__ sq, int64_t n, int D, int d,
                                  const double* __restrict__ pack, double* __restrict__ x, int npairs) {
    extern __shared__ __align__(16) double smem_sym[];
    const int dd = d * d, DD = D * D, K = dd + d * (d + 1) / 2;
    double* Ps = smem_sym;                            // K x (2 npairs): pair p holds entries (r, c0), (r, c0 + 1)
    double* Zs = Ps + static_cast<size_t>(K) * 2 * npairs;   // 2 npairs
    double* u = Zs + 2 * npairs;                      // K x kReconLd
    int* tab = reinterpret_cast<int*>(u + K * kReconLd);     // npairs: r << 8 | c0 ; bit 16: second entry valid
    const double* __restrict__ Z = pack + 2 * D * d;
    const double* __restrict__ P = Z + DD;
    const int tid = threadIdx.x, nthr = blockDim.x;
    // pair table: row r contributes the pairs (r, r), (r, r+2), ... ; an odd run ends with a single
    if (tid == 0) {
        int p = 0;
        for (int r = 0; r < D; ++r)
            for (int c = r; c < D; c += 2) tab[p++] = (r << 8) | c | ((c + 1 < D) ? (1 << 16) : 0);
    }
    __syncthreads();
    for (int e = tid; e < (K + 1) * npairs; e += nthr) {
        const int k = e / npairs, p = e % npairs;
        const int t = tab[p], r = (t >> 8) & 0xff, c0 = t & 0xff;
        const bool two = (t >> 16) & 1;
        const double* src = (k < K) ? P + static_cast<size_t>(k) * DD : Z;
        double* dst = (k < K) ? Ps + static_cast<size_t>(k) * 2 * npairs : Zs;
        dst[2 * p] = src[r * D + c0];
        dst[2 * p + 1] = two ? src[r * D + c0 + 1] : 0.0;
    }
    const int items = npairs * (kReconPts / kReconHalf);
    const int64_t tiles = (n + kReconPts - 1) / kReconPts;
    // Tile inputs u[k][pt] (k-major, row stride kReconLd: the padding spreads the transposing stores over the banks).
    // Entry e = tid + j nthr of a tile is (pt, k) = (e / K, e % K) for EVERY tile, so its source offset and its slot are
    // computed once; the values of the NEXT tile are loaded into registers before the contraction of the current one
    // (ncu on the first version of this kernel: long_scoreboard on these loads + 19 % IMAD for the index arithmetic).
    int src[kReconFill], dst[kReconFill];
#pragma unroll
    for (int j = 0; j < kReconFill; ++j) {
        const int e = tid + j * nthr;
        src[j] = -1;
        dst[j] = 0;
        if (e < K * kReconPts) {
            const int pt = e / K, k = e % K;
            int off;
            if (k < dd) {
                off = k;
            } else {                                  // k - dd enumerates (a, q), a <= q, row-major
                int rem = k - dd, a = 0;
                while (rem >= d - a) {
                    rem -= d - a;
                    ++a;
                }
                off = a * d + a + rem;
            }
            src[j] = (pt * dd + off) * 2 + (k < dd ? 0 : 1);   // low bit: 0 = y, 1 = sq
            dst[j] = (k * kReconLd + pt) | (pt << 20);
        }
    }
    double val[kReconFill];
    auto prefetch = [&](int64_t i0) {
#pragma unroll
        for (int j = 0; j < kReconFill; ++j) {
            double v = 0.0;
            if (src[j] >= 0 && i0 + (dst[j] >> 20) < n) {
                const double* base = (src[j] & 1) ? sq : y;
                v = __ldg(base + i0 * dd + (src[j] >> 1));
            }
            val[j] = v;
        }
    };
    if (static_cast<int64_t>(blockIdx.x) < tiles) prefetch(static_cast<int64_t>(blockIdx.x) * kReconPts);
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int64_t i0 = t * kReconPts;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kReconFill; ++j)
            if (src[j] >= 0) u[dst[j] & 0xfffff] = val[j];
        __syncthreads();
        if (t + gridDim.x < tiles) prefetch((t + gridDim.x) * kReconPts);
        for (int it = tid; it < items; it += nthr) {
            const int p = it % npairs, h = it / npairs;
            double a0[kReconHalf], a1[kReconHalf];
            const double2 zv = *reinterpret_cast<const double2*>(Zs + 2 * p);
#pragma unroll
            for (int pt = 0; pt < kReconHalf; ++pt) {
                a0[pt] = zv.x;
                a1[pt] = zv.y;
            }
            // (a hand-pipelined form of this loop -- operands of step k + 1 loaded during the DFMAs of step k -- measured
            // slower: 0.149 ms against 0.125 ms at N = 65536; the compiler's own schedule of the 4x unrolled loop is kept)
            const double* pp = Ps + 2 * p;
            const double* up = u + h * kReconHalf;
#pragma unroll 4
            for (int k = 0; k < K; ++k, pp += 2 * npairs, up += kReconLd) {
                const double2 pv = *reinterpret_cast<const double2*>(pp);
                const double2* uk = reinterpret_cast<const double2*>(up);
#pragma unroll
                for (int q = 0; q < kReconHalf / 2; ++q) {
                    const double2 uv = uk[q];
                    a0[2 * q] = fma(uv.x, pv.x, a0[2 * q]);
                    a0[2 * q + 1] = fma(uv.y, pv.x, a0[2 * q + 1]);
                    a1[2 * q] = fma(uv.x, pv.y, a1[2 * q]);
                    a1[2 * q + 1] = fma(uv.y, pv.y, a1[2 * q + 1]);
                }
            }
            const int tb = tab[p], r = (tb >> 8) & 0xff, c0 = tb & 0xff;
            const bool two = (tb >> 16) & 1;
            const int64_t ib = i0 + h * kReconHalf;
#pragma unroll
            for (int pt = 0; pt < kReconHalf; ++pt) {
                if (ib + pt < n) {
                    double* xo = x + (ib + pt) * DD;
                    xo[r * D + c0] = a0[pt];
                    if (c0 != r) xo[c0 * D + r] = a0[pt];                 // mirrored entry
                    if (two) {
                        xo[r * D + c0 + 1] = a1[pt];
                        xo[(c0 + 1) * D + r] = a1[pt];
                    }
                }
            }
        }
    }
}


// ---- the contraction on the fp64 tensor cores --------------------------------------------------------------------------
// x_half[n x 2 npairs] = u[n x K] * P_half[K x 2 npairs] + Z_half as mma.sync.m8n8k4.f64 products (DMMA: 256 FMA per warp
// instruction, measured at the full fp64 rate of the SM, scripts/micro/dmma_rate.cu: 36.6 TFLOP/s against 32.6 for DFMA, with
// one operand load per 256 FMA instead of per 3).  M = 8 points, N = 8 entries = 4 column pairs of the upper triangle, K in
// steps of 4.  A warp OWNS kDmTilesN N-tiles for the whole kernel: their B fragments (the operator) are loaded once from
// global memory into kDmStepsK * kDmTilesN registers and never touch shared memory; the A fragments of an M-tile (the
// point inputs u, staged k-major in shared memory as in the kernels above, row stride = 4 mod 16 doubles: conflict-free
// 64-bit fragment loads) are read once per M-tile and used for all N-tiles.  C fragment: lane (l / 4) = point, lanes' two
// values = the two entries of column pair (l % 4): stored to (r, c0), (r, c0 + 1) and mirrored.
constexpr int kDmStepsK = 10;      // K <= 40 (d <= 5)
constexpr int kDmTilesN = 2;       // N-tiles per warp (14 warps x 2 for SPD(20): one 448-thread CTA per SM, no spills)
constexpr int kDmPts = 32;         // points per tile (4 M-tiles)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int KS>
__global__ void __launch_bounds__(448, 1)
nested_spd_reconstruct_dmma_kernel(const double* __restrict__ y, const double* __restrict__ sq, int64_t n, int D, int d,
                                   const double* __restrict__ pack, double* __restrict__ x, int npairs, int nbuf, int xstride) {
    extern __shared__ __align__(16) double u_dm[];    // 2 x {y, sqrt(y)} input tiles | nbuf output tiles (kDmPts x xstride) | barriers
    // Output path.  Scattered 8-byte global stores of the C fragments (and their mirrored twins) cost ~4 L2 sector transactions
    // per 32 bytes written and bound the kernel at 0.15 ms (N = 65536) -- the same bound the DFMA kernels above sit on.  So a
    // tile is assembled in shared memory and leaves as bulk stores (cp.async.bulk shared -> global, one per point: 8 D^2
    // contiguous bytes), two tiles in flight so that the stores of a tile overlap the products of the next one.  Roles of
    // the product: A = operator (M = 8 entries), B = point inputs (N = 8 points), so a lane's C values are ONE entry of TWO
    // points: with the per-point stride xstride = 2 (mod 8) doubles the fragment stores of the upper triangle are
    // bank-conflict free; the mirrored twins are stored from the same fragments.
    const int dd = d * d, DD = D * D, K = dd + d * (d + 1) / 2;
    // tile inputs: the y and sqrt(y) rows of a tile are two contiguous blocks of kDmPts d^2 doubles in global memory: they
    // arrive by TMA bulk loads (two tiles in flight) and the B fragments are read straight from the row-major images
    // (8-byte `__ldg` prefetches into registers + a transposing store cost ~1 us per tile that nothing hid)
    const int raw_tile = kDmPts * dd;                   // doubles per array per tile
    double* raw = u_dm;                                 // [2 buffers][y | sq][kDmPts x d^2]
    double* xs = u_dm + 4 * raw_tile;
    uint64_t* in_bar = reinterpret_cast<uint64_t*>(xs + static_cast<size_t>(nbuf) * kDmPts * xstride);
    const double* __restrict__ Z = pack + 2 * D * d;
    const double* __restrict__ P = Z + DD;
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    // this lane's entries: N-tile j of the warp covers half-vector entries 8 (warp kDmTilesN + j) + g (A fragment rows and
    // C fragment rows alike); entry e = pair e / 2, member e & 1 of the enumeration r = 0.., c0 = r, r + 2, ...
    int e_off[kDmTilesN];         // r * D + c of the entry, -1: padding
    int e_mir[kDmTilesN];         // c * D + r, -1 on the diagonal / padding
    double zc[kDmTilesN];
    double afrag[KS][kDmTilesN];
#pragma unroll
    for (int j = 0; j < kDmTilesN; ++j) {
        const int col = 8 * (warp * kDmTilesN + j) + g, pb = col >> 1;
        e_off[j] = e_mir[j] = -1;
        zc[j] = 0.0;
        if (pb < npairs) {
            int r = 0, left = pb;
            while (r < D && left >= (D - r + 1) / 2) {
                left -= (D - r + 1) / 2;
                ++r;
            }
            const int c = r + 2 * left + (col & 1);
            if (c < D) {
                e_off[j] = r * D + c;
                if (c != r) e_mir[j] = c * D + r;
                zc[j] = Z[r * D + c];
            }
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const int k = 4 * s + t4;
            afrag[s][j] = (e_off[j] >= 0 && k < K) ? P[static_cast<size_t>(k) * DD + e_off[j]] : 0.0;
        }
    }
    // offset of this lane's k = 4 s + t4 inside a point's inputs: k < d^2 -> y entry k, else the (a, q >= a) entry of sqrt(y)
    int koff[KS];
#pragma unroll
    for (int s = 0; s < KS; ++s) {
        const int k = 4 * s + t4;
        koff[s] = -1;
        if (k < dd) {
            koff[s] = k;
        } else if (k < K) {
            int rem = k - dd, a = 0;
            while (rem >= d - a) {
                rem -= d - a;
                ++a;
            }
            koff[s] = raw_tile + a * d + a + rem;
        }
    }
    const int64_t tiles = (n + kDmPts - 1) / kDmPts;
    const uint32_t tile_bytes = static_cast<uint32_t>(raw_tile * sizeof(double));
    if (tid == 0) {
        mbar_init(&in_bar[0], 1);
        mbar_init(&in_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue_in = [&](int64_t t, int b) {             // thread 0; full tiles only (the ragged last tile is copied by hand)
        if (t < tiles && (t + 1) * kDmPts <= n) {
            mbar_expect_tx(&in_bar[b], 2 * tile_bytes);
            tma_load_1d(raw + (2 * b) * raw_tile, y + t * raw_tile, tile_bytes, &in_bar[b]);
            tma_load_1d(raw + (2 * b + 1) * raw_tile, sq + t * raw_tile, tile_bytes, &in_bar[b]);
        }
    };
    if (tid == 0) {
        issue_in(blockIdx.x, 0);
        issue_in(static_cast<int64_t>(blockIdx.x) + gridDim.x, 1);
    }
    // ONE barrier per tile: it publishes the output tile (for the bulk stores), releases the input buffer for the load of
    // the tile after next, and tells that the stores of two tiles ago released their output buffer.
    int it = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int64_t i0 = t * kDmPts;
        const bool full = i0 + kDmPts <= n;
        const bool staged = nbuf > 0 && full;           // the ragged last tile takes the direct stores
        double* xt = xs + static_cast<size_t>(nbuf > 1 ? (it & 1) : 0) * kDmPts * xstride;
        const double* ucur = raw + (2 * (it & 1)) * raw_tile;
        if (full) {
            mbar_wait(&in_bar[it & 1], (it >> 1) & 1);
        } else {                                        // ragged last tile: plain cooperative copy, zero fill
            double* dstb = raw + (2 * (it & 1)) * raw_tile;
            const int64_t valid = (n - i0) * dd;
            for (int e = tid; e < raw_tile; e += nthr) {
                dstb[e] = e < valid ? y[i0 * dd + e] : 0.0;
                dstb[raw_tile + e] = e < valid ? sq[i0 * dd + e] : 0.0;
            }
            __syncthreads();
        }
        if (nbuf == 1 && staged) {                     // single output buffer: wait for the previous tile's stores here
            if (tid < kDmPts) tma_store_wait_read<0>();
            __syncthreads();
        }
#pragma unroll 1
        for (int m0 = 0; m0 < kDmPts; m0 += 16) {       // two M-tiles at a time: four independent DMMA chains per warp
            if (i0 + m0 >= n) break;
            double b[2][KS];
#pragma unroll
            for (int s = 0; s < KS; ++s) {
                b[0][s] = koff[s] >= 0 ? ucur[(m0 + g) * dd + koff[s]] : 0.0;
                b[1][s] = koff[s] >= 0 ? ucur[(m0 + 8 + g) * dd + koff[s]] : 0.0;
            }
            double c[2][kDmTilesN][2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < kDmTilesN; ++j) c[h][j][0] = c[h][j][1] = zc[j];
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int j = 0; j < kDmTilesN; ++j) {
                    dmma(c[0][j], afrag[s][j], b[0][s]);
                    dmma(c[1][j], afrag[s][j], b[1][s]);
                }
            // C fragment: row g = this lane's entry, columns 2 t4, 2 t4 + 1 = points m0 + 8 h + 2 t4 (+ 1)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p0 = m0 + 8 * h + 2 * t4;
                if (staged) {
                    double* o0 = xt + p0 * xstride;
#pragma unroll
                    for (int j = 0; j < kDmTilesN; ++j)
                        if (e_off[j] >= 0) {
                            o0[e_off[j]] = c[h][j][0];
                            o0[xstride + e_off[j]] = c[h][j][1];
                            if (e_mir[j] >= 0) {        // mirrored twin straight from the fragment: 4-way bank conflicts, but
                                o0[e_mir[j]] = c[h][j][0];   // a separate conflict-free mirror pass over the tile costs a second
                                o0[xstride + e_mir[j]] = c[h][j][1];   // barrier and measured slower (0.125 against 0.106 ms)
                            }
                        }
                } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int64_t pt = i0 + p0 + q;
                        if (pt < n) {
                            double* xo = x + pt * DD;
#pragma unroll
                            for (int j = 0; j < kDmTilesN; ++j)
                                if (e_off[j] >= 0) {
                                    xo[e_off[j]] = c[h][j][q];
                                    if (e_mir[j] >= 0) xo[e_mir[j]] = c[h][j][q];
                                }
                        }
                    }
                }
            }
        }
        if (staged) {
            // every earlier bulk store of this thread has finished reading shared memory (they were issued a tile ago): after
            // the barrier the OTHER output buffer is free for the next tile
            if (tid < kDmPts) tma_store_wait_read<0>();
            fence_proxy_async();                        // generic-proxy writes of the tile -> async-proxy reads of the stores
            __syncthreads();
            if (tid < kDmPts) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(x + (i0 + tid) * DD),
                             "r"(smem_u32(xt + tid * xstride)), "r"(static_cast<uint32_t>(DD * sizeof(double)))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            __syncthreads();
        }
        if (tid == 0) {                                 // everyone is past its reads of this input buffer: refill it
            fence_proxy_async();
            issue_in(t + 2 * static_cast<int64_t>(gridDim.x), it & 1);
        }
    }
    if (tid < kDmPts) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores complete before exit
}

}  // namespace
}  // namespace gabo

using namespace gabo;

static int check_nested_dims(const char* who, int D, int dl) {
    GABO_REQUIRE(D >= 2 && D <= kMaxNestedDim && dl >= 1 && dl <= D, GABO_E_ARG,
                 "%s: need 1 <= d_latent <= D <= %d, got D=%d d_latent=%d", who, kMaxNestedDim, D, dl);
    return GABO_OK;
}

extern "C" int gabo_nested_sphere_chain(const double* x, int64_t n, int D, int d_latent, const double* axes,
                                        const double* dists, double* levels, void* stream) {
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_sphere_chain: negative size");
    if (int rc = check_nested_dims("gabo_nested_sphere_chain", D, d_latent)) return rc;
    if (n == 0 || D == d_latent) return GABO_OK;
    GABO_REQUIRE(x && axes && dists && levels, GABO_E_ARG, "gabo_nested_sphere_chain: null pointer");
    nested_sphere_chain_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, n, D, d_latent, axes, dists, levels);
    return check_launch("nested_sphere_chain_kernel");
}

extern "C" int gabo_nested_sphere_to_nested(const double* x, int64_t n, int dim, const double* axis, double dist,
                                            double* y, void* stream) {
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_sphere_to_nested: negative size");
    if (int rc = check_nested_dims("gabo_nested_sphere_to_nested", dim, dim)) return rc;
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x && axis && y, GABO_E_ARG, "gabo_nested_sphere_to_nested: null pointer");
    nested_sphere_to_nested_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0,
                                     static_cast<cudaStream_t>(stream)>>>(x, n, dim, axis, dist, y);
    return check_launch("nested_sphere_to_nested_kernel");
}

extern "C" int gabo_nested_sphere_reconstruct(const double* y, int64_t n, int d_latent, int D, const double* axes,
                                              const double* dists, double* levels, void* stream) {
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_sphere_reconstruct: negative size");
    if (int rc = check_nested_dims("gabo_nested_sphere_reconstruct", D, d_latent)) return rc;
    if (n == 0 || D == d_latent) return GABO_OK;
    GABO_REQUIRE(y && axes && dists && levels, GABO_E_ARG, "gabo_nested_sphere_reconstruct: null pointer");
    nested_sphere_reconstruct_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0,
                                       static_cast<cudaStream_t>(stream)>>>(y, n, d_latent, D, axes, dists, levels);
    return check_launch("nested_sphere_reconstruct_kernel");
}

extern "C" int gabo_spd_sqrtm(const double* mat, int64_t n, int d, double* out, void* stream) {
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_spd_sqrtm: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(mat && out, GABO_E_ARG, "gabo_spd_sqrtm: null pointer");
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM, GABO_E_ARG, "gabo_spd_sqrtm: d=%d outside [1, %d]", d,
                 GABO_MAX_SPD_DIM);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>((n + 63) / 64);
    switch (d) {
#define GABO_CASE(DD)                                           \
    case DD:                                                    \
        spd_sqrtm_kernel<DD><<<grid, 64, 0, s>>>(mat, n, out);  \
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
    return check_launch("spd_sqrtm_kernel");
}

static int check_recon_dims(const char* who, int D, int d) {
    GABO_REQUIRE(d >= 1 && d <= GABO_MAX_SPD_DIM && D > d && D <= kMaxReconDim, GABO_E_ARG,
                 "%s: need 1 <= d <= %d and d < D <= %d, got D=%d d=%d", who, GABO_MAX_SPD_DIM, kMaxReconDim, D, d);
    return GABO_OK;
}

extern "C" int64_t gabo_nested_spd_reconstruct_pack_size(int D, int d) {
    if (d < 1 || D <= d) return 0;
    const int64_t K = static_cast<int64_t>(d) * d + static_cast<int64_t>(d) * (d + 1) / 2;
    return static_cast<int64_t>(2) * D * d + static_cast<int64_t>(D) * D + K * D * D;
}

extern "C" int gabo_nested_spd_reconstruct_setup(const double* w, const double* v, const double* c, const double* k,
                                                 int D, int d, double* pack, int* flag, void* stream) {
    if (int rc = check_recon_dims("gabo_nested_spd_reconstruct_setup", D, d)) return rc;
    GABO_REQUIRE(w && v && c && k && pack && flag, GABO_E_ARG, "gabo_nested_spd_reconstruct_setup: null pointer");
    nested_spd_reconstruct_setup_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(w, v, c, k, D, d, pack, flag);
    return check_launch("nested_spd_reconstruct_setup_kernel");
}

extern "C" int gabo_nested_spd_reconstruct(const double* y, const double* y_sqrt, int64_t n, int D, int d,
                                           const double* pack, double* x, void* stream) {
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_spd_reconstruct: negative size");
    if (int rc = check_recon_dims("gabo_nested_spd_reconstruct", D, d)) return rc;
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(y && y_sqrt && pack && x, GABO_E_ARG, "gabo_nested_spd_reconstruct: null pointer");
    const int K = d * d + d * (d + 1) / 2;
    const int64_t tiles = (n + kReconPts - 1) / kReconPts;
    int npairs = 0;
    for (int r = 0; r < D; ++r) npairs += (D - r + 1) / 2;
    {   // fp64 tensor cores (DMMA): the operator in registers, K <= 40 and at most 14 warps x 2 N-tiles of 4 column pairs
        const int nt = (2 * npairs + 7) / 8;
        const int warps = (nt + kDmTilesN - 1) / kDmTilesN;
        static const bool no_dmma = std::getenv("GABO_RECONSTRUCT_KERNEL") != nullptr;   // developer switch: older kernels
        if (!no_dmma && K <= 4 * kDmStepsK && warps <= 14 && aligned16(y) && aligned16(y_sqrt)) {
            const size_t smem_u = sizeof(double) * 4 * kDmPts * d * d + 2 * sizeof(uint64_t);   // input tiles + their barriers
            int xstride = D * D;                        // per-point stride of the staged tile: = 2 (mod 8) doubles
            while ((xstride & 7) != 2) ++xstride;
            const size_t tile_bytes = sizeof(double) * kDmPts * xstride;
            // staged output needs 16-byte aligned bulk copies of 8 D^2 bytes (D even)
            int nbuf = (aligned16(x) && (D & 1) == 0) ? 2 : 0;
            while (nbuf > 0 && smem_u + nbuf * tile_bytes > 227u * 1024u) --nbuf;
            const size_t smem_dm = smem_u + nbuf * tile_bytes;
            const unsigned grid_dm = static_cast<unsigned>(imin(tiles, static_cast<int64_t>(sm_count())));
            auto go = [&](auto kern) {
                const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                           static_cast<int>(smem_dm));
                if (e != cudaSuccess) return e;
                kern<<<grid_dm, 32 * warps, smem_dm, static_cast<cudaStream_t>(stream)>>>(y, y_sqrt, n, D, d, pack, x, npairs,
                                                                                          nbuf, xstride);
                return cudaSuccess;
            };
            cudaError_t e = cudaSuccess;
            switch ((K + 3) / 4) {                      // k-steps: d = 1 .. 5 -> 1, 2, 4, 7, 10
                case 1: e = go(nested_spd_reconstruct_dmma_kernel<1>); break;
                case 2: e = go(nested_spd_reconstruct_dmma_kernel<2>); break;
                case 4: e = go(nested_spd_reconstruct_dmma_kernel<4>); break;
                case 7: e = go(nested_spd_reconstruct_dmma_kernel<7>); break;
                default: e = go(nested_spd_reconstruct_dmma_kernel<10>); break;
            }
            GABO_REQUIRE(e == cudaSuccess, GABO_E_CUDA, "nested_spd_reconstruct_dmma_kernel: cudaFuncSetAttribute: %s",
                         cudaGetErrorString(e));
            return check_launch("nested_spd_reconstruct_dmma_kernel");
        }
    }
    {   // symmetric form with the operator in shared memory, when it fits twice per SM
        const size_t smem_sym = sizeof(double) * (static_cast<size_t>(K + 1) * 2 * npairs + static_cast<size_t>(K) * kReconLd) +
                                sizeof(int) * npairs + 16;
        if (smem_sym <= 100u * 1024u && tiles >= 4 && K * kReconPts <= kReconFill * kReconSymThreads) {
            const cudaError_t e = cudaFuncSetAttribute(nested_spd_reconstruct_sym_kernel,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       static_cast<int>(smem_sym));
            GABO_REQUIRE(e == cudaSuccess, GABO_E_CUDA, "nested_spd_reconstruct_sym_kernel: cudaFuncSetAttribute: %s",
                         cudaGetErrorString(e));
            const unsigned grid_sym = static_cast<unsigned>(imin(tiles, static_cast<int64_t>(sm_count()) * 2));
            nested_spd_reconstruct_sym_kernel<<<grid_sym, kReconSymThreads, smem_sym, static_cast<cudaStream_t>(stream)>>>(
                y, y_sqrt, n, D, d, pack, x, npairs);
            return check_launch("nested_spd_reconstruct_sym_kernel");
        }
    }
    const size_t smem = sizeof(double) * static_cast<size_t>(K) * kReconPts;
    const unsigned grid = static_cast<unsigned>(imin(tiles, static_cast<int64_t>(sm_count()) * 6));
    // block size: the work items (column pairs x point halves) split evenly over the passes of the item loop
    const int items = ((D * D + 1) / 2) * (kReconPts / kReconHalf);
    const int passes = (items + kReconThreads - 1) / kReconThreads;
    const unsigned threads = static_cast<unsigned>(32 * (((items + passes - 1) / passes + 31) / 32));
    // a quarter of the unified L1 as shared memory (three 10 KB tiles per SM), the rest caches the operator P
    cudaFuncSetAttribute(nested_spd_reconstruct_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 25);
    nested_spd_reconstruct_kernel<<<grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(y, y_sqrt, n, D, d, pack, x);
    return check_launch("nested_spd_reconstruct_kernel");
}
