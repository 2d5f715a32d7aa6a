// Batched symmetric eigendecomposition for matrices beyond the register kernels (8 < D <= 32), fp64.
//
//   gabo_sym_eig : M_n = U_n diag(lambda_n) U_n^T for a batch of symmetric D x D matrices.
//
// Used by the reconstruction-parameter fit of the nested SPD mapping (nested_optimization.py; reference
// nested_mappings/nested_spd_optimization.py:22-92, where every cost evaluation runs N torch.symeig calls of size D = 10 .. 20
// inside a Python loop: affine_invariant_distance_torch spd_utils_torch.py:108-112, logm_torch :25-30, sqrtm_torch :45-50).
//
// One warp per matrix.  The matrix and the accumulated rotations live in shared memory (leading dimension D + 1: a lane
// walking down a column and a lane walking along a row are both conflict-free).  Two-sided cyclic Jacobi with the round-robin
// (tournament) ordering: every step holds D/2 disjoint pairs, so the D/2 rotations of a step commute and are applied as three
// sweeps over the matrix -- lane i updates row i of A J and of V J for every pair, then lane j updates column j of J^T (A J) --
// with the (c, s) of the step computed by the first D/2 lanes and broadcast through shared memory.  Convergence: the largest
// |a_pq| / sqrt(|a_pp a_qq|) seen in a sweep below 1e-15 (the same relative criterion as the register kernels), at most 30
// sweeps.  The latency of a solve does not matter here (a few hundred matrices per cost evaluation, all warps in flight at
// once); what matters is that the whole batch is ONE launch with no host round trip.
#include "common.cuh"

namespace gabo {
namespace {

constexpr int kEigMaxDim = 32;
constexpr int kEigWarps = 4;   // matrices per CTA

__global__ void __launch_bounds__(32 * kEigWarps)
sym_eig_kernel(const double* __restrict__ mats, int64_t n, int D, double* __restrict__ evals, double* __restrict__ evecs,
               int32_t* __restrict__ flags) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * kEigWarps + warp;
    if (idx >= n) return;                      // whole warp leaves: no CTA-wide barrier below
    const int ld = D + 1;
    const int per_warp = 2 * D * ld + 2 * kEigMaxDim;
    double* A = smem + static_cast<size_t>(warp) * per_warp;
    double* V = A + D * ld;
    double* cs = V + D * ld;                   // c of pair k at cs[k], s at cs[16 + k]
    const double* src = mats + idx * D * D;
    bool finite = true;
    for (int e = lane; e < D * D; e += 32) {
        const int r = e / D, c = e - r * D;
        const double v = src[e];
        finite = finite && isfinite(v);
        A[r * ld + c] = v;
        V[r * ld + c] = (r == c) ? 1.0 : 0.0;
    }
    finite = __all_sync(0xffffffffu, finite);
    __syncwarp();
    // symmetrise (the callers build M = L^-1 X L^-T by two products: symmetric up to rounding)
    for (int e = lane; e < D * D; e += 32) {
        const int r = e / D, c = e - r * D;
        if (r < c) {
            const double m = 0.5 * (A[r * ld + c] + A[c * ld + r]);
            A[r * ld + c] = m;
            A[c * ld + r] = m;
        }
    }
    __syncwarp();
    const int m = (D + 1) & ~1;                // players of the tournament (one bye when D is odd)
    const int half = m >> 1;
    bool converged = !finite;
    for (int sweep = 0; sweep < 30 && !converged; ++sweep) {
        double worst = 0.0;
        for (int step = 0; step < m - 1; ++step) {
            // round-robin pairing: player m-1 is fixed, the others rotate
            auto player = [&](int pos) { return pos == m - 1 ? m - 1 : (pos + step) % (m - 1); };
            if (lane < half) {
                int p = player(lane), q = player(m - 1 - lane);
                double c = 1.0, s = 0.0;
                if (p < D && q < D) {
                    if (p > q) { const int t = p; p = q; q = t; }
                    const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                    const double scale = sqrt(fabs(app * aqq));
                    const double rel = fabs(apq) / (scale > 0.0 ? scale : 1.0);
                    worst = fmax(worst, rel);
                    if (rel > 1e-17 && apq != 0.0) {
                        const double tau = (aqq - app) / (2.0 * apq);
                        const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = rsqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                cs[lane] = c;
                cs[16 + lane] = s;
            }
            __syncwarp();
            // A <- A J and V <- V J: lane = row
            if (lane < D) {
                double* arow = A + lane * ld;
                double* vrow = V + lane * ld;
                for (int k = 0; k < half; ++k) {
                    int p = player(k), q = player(m - 1 - k);
                    if (p >= D || q >= D) continue;
                    if (p > q) { const int t = p; p = q; q = t; }
                    const double c = cs[k], s = cs[16 + k];
                    const double ap = arow[p], aq = arow[q];
                    arow[p] = c * ap - s * aq;
                    arow[q] = s * ap + c * aq;
                    const double vp = vrow[p], vq = vrow[q];
                    vrow[p] = c * vp - s * vq;
                    vrow[q] = s * vp + c * vq;
                }
            }
            __syncwarp();
            // A <- J^T A: lane = column
            if (lane < D) {
                for (int k = 0; k < half; ++k) {
                    int p = player(k), q = player(m - 1 - k);
                    if (p >= D || q >= D) continue;
                    if (p > q) { const int t = p; p = q; q = t; }
                    const double c = cs[k], s = cs[16 + k];
                    const double ap = A[p * ld + lane], aq = A[q * ld + lane];
                    A[p * ld + lane] = c * ap - s * aq;
                    A[q * ld + lane] = s * ap + c * aq;
                }
            }
            __syncwarp();
        }
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, 16));
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, 8));
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, 4));
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, 2));
        worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, 1));
        converged = worst < 1e-15;
    }
    if ((!converged || !finite) && lane == 0 && flags) atomicOr(flags, 1);
    if (lane < D) evals[idx * D + lane] = finite ? A[lane * ld + lane] : nan("");
    if (evecs) {
        double* dst = evecs + idx * D * D;
        for (int e = lane; e < D * D; e += 32) {
            const int r = e / D, c = e - r * D;
            dst[e] = V[r * ld + c];              // column c = eigenvector of evals[c]
        }
    }
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_sym_eig(const double* mats, int64_t n, int D, double* evals, double* evecs, int32_t* flags,
                            void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_sym_eig: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(mats && evals, GABO_E_ARG, "gabo_sym_eig: null pointer");
    GABO_REQUIRE(D >= 1 && D <= kEigMaxDim, GABO_E_ARG, "gabo_sym_eig: D=%d outside [1, %d]", D, kEigMaxDim);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(kEigWarps) * (2 * D * (D + 1) + 2 * kEigMaxDim) * sizeof(double);
    if (smem > 48 * 1024) {
        GABO_REQUIRE(cudaFuncSetAttribute(sym_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)) == cudaSuccess,
                     GABO_E_CUDA, "gabo_sym_eig: cannot reserve %zu bytes of shared memory", smem);
    }
    const unsigned grid = static_cast<unsigned>((n + kEigWarps - 1) / kEigWarps);
    sym_eig_kernel<<<grid, 32 * kEigWarps, smem, s>>>(mats, n, D, evals, evecs, flags);
    return check_launch("sym_eig_kernel");
}
