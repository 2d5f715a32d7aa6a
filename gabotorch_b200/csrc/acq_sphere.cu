// Acquisition value / multi-start Riemannian CG on the sphere (A1 + A4 with M1 primitives).
// One warp per restart, all CG steps inside one launch; see acq_common.cuh for the scheme.
//
// Manifold operations used by the solver are pymanopt's Sphere (reference call sites manifold_optimize.py:207-221,
// numpy statements Riemannian_utils/sphere_utils.py:14-123): retr(x,u) = (x+u)/|x+u|, transp(x,y,u) = u - <y,u> y,
// inner = Euclidean dot, Log_x(y) = (y - <x,y> x) * theta / |y - <x,y> x|.
#include <cstdlib>
#include <type_traits>

#include "acq_common.cuh"

namespace gabo {
namespace {

template <typename T>
struct WarpVecs {
    int D;
    int lane;
    __device__ __forceinline__ T dot(const T* a, const T* b) const {
        T s = T(0);
        for (int k = lane; k < D; k += 32) s = fma(a[k], b[k], s);
        return warp_sum(s);
    }
};

template <typename T, int NCH>
__global__ void __launch_bounds__(kAcqWarps * 32)
    sphere_acq_kernel(GpParams gp, RcgParams opt, int mode, double* __restrict__ x_io, int64_t r,
                      double* __restrict__ value, double* __restrict__ grad_out, int32_t* __restrict__ iters,
                      int32_t* __restrict__ reason) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n, D = gp.dim;
    const int Dp = D | 1;            // odd row stride: conflict-free lane-per-row reads
    const int npad = (n + 3) & ~3;
    const int Dv = (D + 3) & ~3;
    SmemCarver cv;
    T* Xs = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * n * Dp));
    T* alpha = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad));
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * n * n));
    T* wbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kAcqWarps * (5 * Dv + 3 * npad)));

    for (int e = threadIdx.x; e < n * D; e += blockDim.x) Xs[(e / D) * Dp + (e % D)] = static_cast<T>(gp.x_train[e]);
    for (int e = threadIdx.x; e < n; e += blockDim.x) alpha[e] = static_cast<T>(gp.alpha[e]);
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Minv[e] = static_cast<T>(gp.minv[e]);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rid = static_cast<int64_t>(blockIdx.x) * kAcqWarps + warp;
    if (rid >= r) return;

    T* xv = wbase + warp * (5 * Dv + 3 * npad);
    T* gv = xv + Dv;
    T* eta = gv + Dv;
    T* xn = eta + Dv;
    T* gn = xn + Dv;
    T* ksh = gn + Dv;
    T* gsh = ksh + npad;
    T* csh = gsh + npad;
    WarpVecs<T> wv{D, lane};

    const T s_out = static_cast<T>(gp.outputscale), beta = static_cast<T>(gp.beta);
    T c_l[NCH], th_l[NCH], k_l[NCH], mk_l[NCH];
    EiScalars<T> sc;

    // cost(p) = -EI(p); leaves the per-point quantities in the lane registers for grad()
    auto cost_at = [&](const T* p) -> T {
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T kk = T(0), cc = T(0), tt = T(0);
            if (i < n) {
                const T* xi = Xs + i * Dp;
                for (int k = 0; k < D; ++k) cc = fma(xi[k], p[k], cc);
                cc = fmin(fmax(cc, T(-1) + M<T>::clamp_eps()), T(1) - M<T>::clamp_eps());
                tt = M<T>::acos_(cc);
                kk = s_out * M<T>::exp_(-beta * tt * tt);
                ksh[i] = kk;
            }
            c_l[ch] = cc;
            th_l[ch] = tt;
            k_l[ch] = kk;
        }
        __syncwarp();
        T ka = T(0), kmk = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T mk = T(0);
            if (i < n) {
                // four independent accumulators: the n-long dependent FMA chain was the longest latency of a cost call
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

    // Riemannian gradient of the cost at p (the point of the last cost_at call) into out[]
    auto grad_at = [&](const T* p, T* out) {
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            if (i < n) {
                const T w = -sc.cdf * alpha[i] - sc.pdf_over_sigma * mk_l[ch];
                const T coef = T(2) * beta * w * k_l[ch];
                const T* xi = Xs + i * Dp;
                T pn = T(0);
                for (int k = 0; k < D; ++k) {
                    const T pk = fma(-c_l[ch], p[k], xi[k]);
                    pn = fma(pk, pk, pn);
                }
                pn = M<T>::sqrt_(pn);
                const T scale = (th_l[ch] > T(1e-6)) ? th_l[ch] / (pn > T(0) ? pn : T(1)) : T(1);
                gsh[i] = coef * scale;
                csh[i] = c_l[ch];
            }
        }
        __syncwarp();
        for (int k = lane; k < D; k += 32) {
            T acc = T(0), sgc = T(0);
            for (int i = 0; i < n; ++i) {
                acc = fma(gsh[i], Xs[i * Dp + k], acc);
                sgc = fma(gsh[i], csh[i], sgc);
            }
            out[k] = -(acc - sgc * p[k]);  // cost = -EI
        }
        __syncwarp();
    };

    const double* xin = x_io + rid * D;
    for (int k = lane; k < D; k += 32) xv[k] = static_cast<T>(xin[k]);
    __syncwarp();

    T cost = cost_at(xv);
    if (mode == 0) {
        if (lane == 0) value[rid] = static_cast<double>(-cost);
        if (grad_out) {
            grad_at(xv, gv);
            for (int k = lane; k < D; k += 32) grad_out[rid * D + k] = static_cast<double>(-gv[k]);
        }
        return;
    }

    grad_at(xv, gv);
    T gPg = wv.dot(gv, gv);
    T gradnorm = M<T>::sqrt_(gPg);
    for (int k = lane; k < D; k += 32) eta[k] = -gv[k];
    __syncwarp();
    int it = 0, why = 0;
    T stepsize = M<T>::nan();
    T oldalpha = T(-1);  // unset
    const T mingrad = static_cast<T>(opt.mingradnorm), minstep = static_cast<T>(opt.minstepsize);
    const T contraction = static_cast<T>(opt.contraction), suff = static_cast<T>(opt.suff_decr);

    auto retract = [&](T a) {  // xn = (x + a eta) / |x + a eta|
        T s = T(0);
        for (int k = lane; k < D; k += 32) {
            const T t = fma(a, eta[k], xv[k]);
            xn[k] = t;
            s = fma(t, t, s);
        }
        s = warp_sum(s);
        const T inv = T(1) / M<T>::sqrt_(s);
        for (int k = lane; k < D; k += 32) xn[k] *= inv;
        __syncwarp();
    };

    while (true) {
        if (it + 1 >= opt.maxiter) { why = 1; break; }
        if (gradnorm < mingrad) { why = 2; break; }
        if (stepsize < minstep) { why = 3; break; }
        T df0 = wv.dot(gv, eta);
        if (df0 >= T(0)) {  // not a descent direction: restart from steepest descent
            __syncwarp();
            for (int k = lane; k < D; k += 32) eta[k] = -gv[k];
            __syncwarp();
            df0 = -gPg;
        }
        const T norm_d = M<T>::sqrt_(wv.dot(eta, eta));
        T a = (oldalpha >= T(0)) ? oldalpha : static_cast<T>(opt.initial_stepsize) / norm_d;
        retract(a);
        T newf = cost_at(xn);
        int evals = 1;
        while (newf > cost + suff * a * df0 && evals <= opt.ls_maxiter) {
            a *= contraction;
            retract(a);
            newf = cost_at(xn);
            ++evals;
        }
        if (newf > cost) {  // no decrease: stay
            a = T(0);
            for (int k = lane; k < D; k += 32) {
                xn[k] = xv[k];
                gn[k] = gv[k];
            }
            __syncwarp();
            newf = cost;
        } else {
            grad_at(xn, gn);
        }
        stepsize = a * norm_d;
        oldalpha = (evals == 2) ? a : T(2) * a;
        // transport g and eta to xn (projection), Hestenes-Stiefel beta
        const T xg = wv.dot(xn, gv), xe = wv.dot(xn, eta);
        T ip = T(0), den = T(0), ngg = T(0);
        for (int k = lane; k < D; k += 32) {
            const T og = fma(-xg, xn[k], gv[k]);
            const T oe = fma(-xe, xn[k], eta[k]);
            const T df = gn[k] - og;
            ip = fma(gn[k], df, ip);
            den = fma(df, oe, den);
            ngg = fma(gn[k], gn[k], ngg);
            eta[k] = oe;
        }
        ip = warp_sum(ip);
        den = warp_sum(den);
        ngg = warp_sum(ngg);
        T bcg;
        if (den == T(0)) {
            bcg = T(1);  // pymanopt: ZeroDivisionError branch for float inner products
        } else {
            const T q = ip / den;
            bcg = (q > T(0)) ? q : T(0);
        }
        __syncwarp();
        for (int k = lane; k < D; k += 32) {
            eta[k] = fma(bcg, eta[k], -gn[k]);
            xv[k] = xn[k];
            gv[k] = gn[k];
        }
        __syncwarp();
        cost = newf;
        gPg = ngg;
        gradnorm = M<T>::sqrt_(ngg);
        ++it;
    }

    // write back: renormalised in fp64 so the candidate is on the sphere to fp64 accuracy
    double s = 0.0;
    for (int k = lane; k < D; k += 32) s = fma(static_cast<double>(xv[k]), static_cast<double>(xv[k]), s);
    s = warp_sum(s);
    const double inv = 1.0 / sqrt(s);
    for (int k = lane; k < D; k += 32) x_io[rid * D + k] = static_cast<double>(xv[k]) * inv;
    if (lane == 0) {
        value[rid] = static_cast<double>(-cost);
        if (iters) iters[rid] = it;
        if (reason) reason[rid] = why;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Register-resident variant for the common case (ambient dimension <= 16, e.g. S^5 of BASELINE config 3).
// ncu on the shared-memory version above: ~2100 warp-instructions per CG iteration issued at ~15 cycles each -- every
// vector operation was LDS -> op -> STS -> __syncwarp on 6 of 32 lanes, every inner product a 5-step shuffle tree.
// Here every lane carries the WHOLE iterate / gradient / direction in registers (DP values, zero-padded from D), so
// retraction, transport, inner products and the CG recurrences are a handful of FMAs computed redundantly by all
// lanes with no communication; lanes still own the training points (one per lane and chunk), whose coordinates and
// alpha also live in registers.  Warp shuffles remain only where the math sums over training points: the posterior
// mean / variance (2 sums per cost call) and the gradient (D + 1 sums).  Same formulas as the kernel above.
// ---------------------------------------------------------------------------------------------------------------
// Four consecutive shared-memory values with one (fp32) or two (fp64) vector loads.
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

// Per-warp evaluator: lanes own the training points (coordinates and alpha in registers), the query point is replicated.
// For n <= 32 (NCH == 1, the regime of the reference's BO loops: 5..35 points) the lane's row of K^-1 lives in
// registers too, so the K^-1 k product is 8 broadcast vector loads of k + 32 FMAs with no address arithmetic.
template <typename T, int DP, int NCH>
struct SphereEval {
    static constexpr bool kRowInRegs = (NCH == 1);
    static constexpr int kKsh = kRowInRegs ? 32 : 0;       // minimum size of the per-warp k buffer
    T Xi[NCH][DP], al[NCH];
    T Mrow[kRowInRegs ? 32 : 1];
    T c_l[NCH], th2_l[NCH], k_l[NCH], mk_l[NCH];   // per-point quantities of the last cost() call, kept for grad()
    EiScalars<T> sc;
    T s_out, beta;
    int n, npad, lane;
    const T* Minv;   // shared: npad x n, rows n..npad-1 zero
    T* ksh;          // shared: max(npad, kKsh) values, private to the warp

    static __host__ __device__ int ksh_size(int n_train) {
        const int np = (n_train + 3) & ~3;
        return np > kKsh ? np : kKsh;
    }

    static __device__ __forceinline__ T dot(const T (&a)[DP], const T (&b)[DP]) {
        T s = a[0] * b[0];
#pragma unroll
        for (int k = 1; k < DP; ++k) s = fma(a[k], b[k], s);
        return s;
    }

    __device__ __forceinline__ void load(const GpParams& gp, int lane_, const T* minv_s, T* ksh_w) {
        n = gp.n;
        npad = (n + 3) & ~3;
        lane = lane_;
        Minv = minv_s;
        ksh = ksh_w;
        s_out = static_cast<T>(gp.outputscale);
        beta = static_cast<T>(gp.beta);
        const int D = gp.dim;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            al[ch] = (i < n) ? static_cast<T>(gp.alpha[i]) : T(0);
#pragma unroll
            for (int k = 0; k < DP; ++k) Xi[ch][k] = (i < n && k < D) ? static_cast<T>(gp.x_train[i * D + k]) : T(0);
        }
        if (kRowInRegs) {
#pragma unroll
            for (int j = 0; j < (kRowInRegs ? 32 : 1); ++j) Mrow[j] = (lane < n && j < n) ? minv_s[j * n + lane] : T(0);
        }
    }

    // cost(p) = -EI(p); all lanes return the same value
    __device__ __forceinline__ T cost(const T (&p)[DP], const GpParams& gp) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T cc = dot(Xi[ch], p);
            cc = fmin(fmax(cc, T(-1) + M<T>::clamp_eps()), T(1) - M<T>::clamp_eps());
            const T t2 = M<T>::theta2_(cc);
            const T kk = (i < n) ? s_out * M<T>::exp_neg_(-beta * t2) : T(0);
            if (kRowInRegs || i < npad) ksh[i] = kk;
            c_l[ch] = cc;
            th2_l[ch] = t2;
            k_l[ch] = kk;
        }
        __syncwarp();
        T ka = T(0), kmk = T(0);
        if (kRowInRegs) {
            T m[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                T kv[4];
                ld4(ksh + j, kv);
#pragma unroll
                for (int q = 0; q < 4; ++q) m[q] = fma(Mrow[kRowInRegs ? j + q : 0], kv[q], m[q]);
            }
            const T mk = (m[0] + m[1]) + (m[2] + m[3]);
            mk_l[0] = mk;
            ka = k_l[0] * al[0];
            kmk = k_l[0] * mk;
        } else {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int i = lane + 32 * ch;
                T mk = T(0);
                if (i < n) {
                    T m0 = T(0), m1 = T(0), m2 = T(0), m3 = T(0);
                    for (int j = 0; j < npad; j += 4) {     // K^-1 is symmetric: row i read as column i, conflict-free
                        T kv[4];
                        ld4(ksh + j, kv);
                        m0 = fma(Minv[j * n + i], kv[0], m0);
                        m1 = fma(Minv[(j + 1) * n + i], kv[1], m1);
                        m2 = fma(Minv[(j + 2) * n + i], kv[2], m2);
                        m3 = fma(Minv[(j + 3) * n + i], kv[3], m3);
                    }
                    mk = (m0 + m1) + (m2 + m3);
                }
                mk_l[ch] = mk;
                ka = fma(k_l[ch], al[ch], ka);
                kmk = fma(k_l[ch], mk, kmk);
            }
        }
        __syncwarp();   // ksh is rewritten by the next call
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ka += __shfl_xor_sync(0xffffffffu, ka, o);
            kmk += __shfl_xor_sync(0xffffffffu, kmk, o);
        }
        sc = ei_scalars<T>(ka, kmk, gp);
        const T cst = -sc.ei;
        return (cst == cst) ? cst : M<T>::inf();
    }

    // Riemannian gradient of the cost at p (the point of the last cost() call); all lanes get the whole vector
    __device__ __forceinline__ void grad(const T (&p)[DP], T (&out)[DP]) {
        T acc[DP], sgc = T(0);
#pragma unroll
        for (int k = 0; k < DP; ++k) acc[k] = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const T w = -sc.cdf * al[ch] - sc.pdf_over_sigma * mk_l[ch];
            const T coef = T(2) * beta * w * k_l[ch];          // k_l = 0 for lanes without a training point
            T pn2 = T(0);
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const T pk = fma(-c_l[ch], p[k], Xi[ch][k]);
                pn2 = fma(pk, pk, pn2);
            }
            // Log_p(X_i) = (X_i - c p) * theta / |X_i - c p|;  unscaled when theta <= 1e-6 (pymanopt Sphere.log)
            const T th = M<T>::theta_(th2_l[ch]);
            const T scale = (th > T(1e-6) && pn2 > T(0)) ? th * M<T>::rsqrt_(pn2) : T(1);
            const T g = coef * scale;
#pragma unroll
            for (int k = 0; k < DP; ++k) acc[k] = fma(g, Xi[ch][k], acc[k]);
            sgc = fma(g, c_l[ch], sgc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // D + 1 interleaved butterflies
#pragma unroll
            for (int k = 0; k < DP; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            sgc += __shfl_xor_sync(0xffffffffu, sgc, o);
        }
#pragma unroll
        for (int k = 0; k < DP; ++k) out[k] = -(acc[k] - sgc * p[k]);  // cost = -EI
    }
};

// A4: EI (and its Riemannian gradient) at r points, one warp per point.
template <typename T, int DP, int NCH>
__global__ void __launch_bounds__(kAcqWarps * 32)
    sphere_ei_reg_kernel(GpParams gp, const double* __restrict__ x_in, int64_t r, double* __restrict__ value,
                         double* __restrict__ grad_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n, D = gp.dim;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad * n));
    const int ksz = SphereEval<T, DP, NCH>::ksh_size(n);
    T* kbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kAcqWarps * ksz));
    for (int e = threadIdx.x; e < npad * n; e += blockDim.x) Minv[e] = (e < n * n) ? static_cast<T>(gp.minv[e]) : T(0);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rid = static_cast<int64_t>(blockIdx.x) * kAcqWarps + warp;
    if (rid >= r) return;
    SphereEval<T, DP, NCH> ev;
    ev.load(gp, lane, Minv, kbase + warp * ksz);
    T xv[DP], gv[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xv[k] = (k < D) ? static_cast<T>(x_in[rid * D + k]) : T(0);
    const T cost = ev.cost(xv, gp);
    if (lane == 0) value[rid] = static_cast<double>(-cost);
    if (grad_out) {
        ev.grad(xv, gv);
#pragma unroll
        for (int k = 0; k < DP; ++k)
            if (k < D && lane == 0) grad_out[rid * D + k] = static_cast<double>(-gv[k]);
    }
}

// A1: multi-start conjugate gradient, ONE CTA PER RESTART with speculative backtracking across its warps.
// pymanopt's LineSearchAdaptive tries alpha, alpha c, alpha c^2, ... one cost call at a time and takes the FIRST step
// that passes the Armijo test.  The cost calls of different trial steps are independent, so warp w of the CTA evaluates
// trial (base + w) and the first acceptable one wins: the accepted step, the number of cost evaluations it accounts
// for and hence every later iterate are exactly those of the sequential search, at the latency of one cost call per
// kSpec trials.  (ncu on the warp-per-restart version: latency-bound at 1-2 warps per scheduler, and the launch time
// was set by the restarts whose line search backtracks to the limit, 11 sequential cost calls per iteration.)
// Every warp carries the whole CG state in registers and executes the same scalar recurrences on the same data, so
// control flow is identical across the CTA; only the trial costs and the new point + gradient go through shared memory.
// kSpec = warps per restart = trial steps evaluated concurrently (1 = plain sequential search, one warp per restart)
template <typename T, int DP, int NCH, int kSpec>
__global__ void __launch_bounds__(kSpec * 32)
    sphere_rcg_cta_kernel(GpParams gp, RcgParams opt, double* __restrict__ x_io, int64_t r, double* __restrict__ value,
                          int32_t* __restrict__ iters, int32_t* __restrict__ reason) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n, D = gp.dim;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad * n));
    const int ksz = SphereEval<T, DP, NCH>::ksh_size(n);
    T* kbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kSpec * ksz));
    T* f_sh = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * 2 * kSpec));      // trial costs, double-buffered
    T* xg_sh = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * 2 * 2 * DP));    // new point | gradient, double-buffered
    for (int e = threadIdx.x; e < npad * n; e += blockDim.x) Minv[e] = (e < n * n) ? static_cast<T>(gp.minv[e]) : T(0);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    SphereEval<T, DP, NCH> ev;
    ev.load(gp, lane, Minv, kbase + warp * ksz);
    using E = SphereEval<T, DP, NCH>;
    const T mingrad = static_cast<T>(opt.mingradnorm), minstep = static_cast<T>(opt.minstepsize);
    const T contraction = static_cast<T>(opt.contraction), suff = static_cast<T>(opt.suff_decr);

    for (int64_t rid = blockIdx.x; rid < r; rid += gridDim.x) {
        T xv[DP], gv[DP], eta[DP], xn[DP], gn[DP], xt[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) xv[k] = (k < D) ? static_cast<T>(x_io[rid * D + k]) : T(0);
        T cost = ev.cost(xv, gp);   // every warp computes the start redundantly: same data, same result
        ev.grad(xv, gv);
        T gPg = E::dot(gv, gv);
        T gradnorm = M<T>::sqrt_(gPg);
#pragma unroll
        for (int k = 0; k < DP; ++k) eta[k] = -gv[k];
        int it = 0, why = 0;
        unsigned fbuf = 0, gbuf = 0;
        T stepsize = M<T>::nan();
        T oldalpha = T(-1);  // unset

        while (true) {
            if (it + 1 >= opt.maxiter) { why = 1; break; }
            if (gradnorm < mingrad) { why = 2; break; }
            if (stepsize < minstep) { why = 3; break; }
            T df0 = E::dot(gv, eta);
            if (df0 >= T(0)) {  // not a descent direction: restart from steepest descent
#pragma unroll
                for (int k = 0; k < DP; ++k) eta[k] = -gv[k];
                df0 = -gPg;
            }
            const T norm_d = M<T>::sqrt_(E::dot(eta, eta));
            // trial k (k = 0 .. ls_maxiter) uses alpha_k = alpha_0 c^k; the first one passing the Armijo test wins, the
            // last one is kept when none passes (LineSearchAdaptive leaves its loop after ls_maxiter + 1 cost calls)
            T a = (oldalpha >= T(0)) ? oldalpha : static_cast<T>(opt.initial_stepsize) / norm_d;
            const int ktotal = opt.ls_maxiter + 1;
            int evals = 0, sel = 0;
            T newf = cost;
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
                {   // my trial: retraction (x + a eta) / |x + a eta| and its cost
                    T s = T(0);
#pragma unroll
                    for (int k = 0; k < DP; ++k) {
                        xt[k] = fma(amine, eta[k], xv[k]);
                        s = fma(xt[k], xt[k], s);
                    }
                    const T inv = T(1) / M<T>::sqrt_(s);
#pragma unroll
                    for (int k = 0; k < DP; ++k) xt[k] *= inv;
                }
                const T fmine = (base + warp < ktotal) ? ev.cost(xt, gp) : M<T>::inf();
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
                        a = ak[t];                   // on exit: the step of the accepted (or of the last) trial
                        done = !(ft > cost + suff * ak[t] * df0);
                    }
                }
                if (!done && base + kSpec < ktotal) a = ak[kSpec - 1] * contraction;
            }
            if (newf > cost) {  // no decrease: stay
                a = T(0);
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    xn[k] = xv[k];
                    gn[k] = gv[k];
                }
                newf = cost;
            } else {            // the winning warp holds the per-point state of the accepted point: it takes the gradient
                T* xb = xg_sh + gbuf * 2 * DP;
                gbuf ^= 1u;
                if (warp == sel) {
                    ev.grad(xt, gn);
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < DP; ++k) {
                            xb[k] = xt[k];
                            xb[DP + k] = gn[k];
                        }
                    }
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    xn[k] = xb[k];
                    gn[k] = xb[DP + k];
                }
            }
            stepsize = a * norm_d;
            oldalpha = (evals == 2) ? a : T(2) * a;
            // transport g and eta to xn (projection), Hestenes-Stiefel beta
            const T xg = E::dot(xn, gv), xe = E::dot(xn, eta);
            T ip = T(0), den = T(0), ngg = T(0);
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const T og = fma(-xg, xn[k], gv[k]);
                const T oe = fma(-xe, xn[k], eta[k]);
                const T df = gn[k] - og;
                ip = fma(gn[k], df, ip);
                den = fma(df, oe, den);
                ngg = fma(gn[k], gn[k], ngg);
                eta[k] = oe;
            }
            T bcg;
            if (den == T(0)) {
                bcg = T(1);  // pymanopt: ZeroDivisionError branch for float inner products
            } else {
                const T q = ip / den;
                bcg = (q > T(0)) ? q : T(0);
            }
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                eta[k] = fma(bcg, eta[k], -gn[k]);
                xv[k] = xn[k];
                gv[k] = gn[k];
            }
            cost = newf;
            gPg = ngg;
            gradnorm = M<T>::sqrt_(ngg);
            ++it;
        }

        // write back: renormalised in fp64 so the candidate is on the sphere to fp64 accuracy
        if (warp == 0 && lane == 0) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DP; ++k) s = fma(static_cast<double>(xv[k]), static_cast<double>(xv[k]), s);
            const double inv = 1.0 / sqrt(s);
#pragma unroll
            for (int k = 0; k < DP; ++k)
                if (k < D) x_io[rid * D + k] = static_cast<double>(xv[k]) * inv;
            value[rid] = static_cast<double>(-cost);
            if (iters) iters[rid] = it;
            if (reason) reason[rid] = why;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Trust-region acquisition solver, ONE WARP PER RESTART.  Restates the reference's own
// TrustRegions.solve / _truncated_conjugate_gradient (BoManifolds/manifold_optimization/robust_trust_regions.py:116-352,
// :410-520, use_rand=False, identity preconditioner, the d_Hd != 0 guard of :461-465) with the finite-difference
// Hessian of approximate_hessian.py:11-62 (H[a] = (T_{x1->x} grad f(x1) - grad f(x)) / c, x1 = R_x(c a),
// c = 2^-14 / |a|) -- the combination gen_candidates_manifold builds for approx_hessian=True
// (manifold_optimize.py:199-200).  Sphere operations as in pymanopt: R_x(u) = (x + u) / |x + u|, T_{x1->x} = P_x.
// Every vector of the solver (iterate, gradient, eta, H eta, residual, search direction, H delta) is replicated in
// the registers of all 32 lanes, as in the CG kernel; lanes differ only in the training points they own inside
// SphereEval, so the whole tCG recurrence is warp-uniform scalar code and one Hessian-vector product costs one
// cost + gradient evaluation.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct Eps;
template <>
struct Eps<float> {
    static __device__ __forceinline__ float value() { return 1.1920929e-7f; }
};
template <>
struct Eps<double> {
    static __device__ __forceinline__ double value() { return 2.220446049250313e-16; }   // np.spacing(1)
};

template <typename T, int DP, int NCH>
__global__ void __launch_bounds__(kAcqWarps * 32)
    sphere_rtr_kernel(GpParams gp, RtrParams opt, double* __restrict__ x_io, int64_t r, double* __restrict__ value,
                      int32_t* __restrict__ iters, int32_t* __restrict__ reason) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n, D = gp.dim;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad * n));
    const int ksz = SphereEval<T, DP, NCH>::ksh_size(n);
    T* kbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kAcqWarps * ksz));
    for (int e = threadIdx.x; e < npad * n; e += blockDim.x) Minv[e] = (e < n * n) ? static_cast<T>(gp.minv[e]) : T(0);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rid = static_cast<int64_t>(blockIdx.x) * kAcqWarps + warp;
    if (rid >= r) return;
    using E = SphereEval<T, DP, NCH>;
    E ev;
    ev.load(gp, lane, Minv, kbase + warp * ksz);
    const T mingrad = static_cast<T>(opt.mingradnorm), kappa = static_cast<T>(opt.kappa);
    const T theta = static_cast<T>(opt.theta), rho_prime = static_cast<T>(opt.rho_prime);
    const T delta_bar = static_cast<T>(opt.delta_bar), fd_eps = static_cast<T>(opt.fd_eps);
    const T rho_scale = Eps<T>::value() * static_cast<T>(opt.rho_regularization);

    T xv[DP], g[DP], eta[DP], heta[DP], rv[DP], dl[DP], hd[DP], xt[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xv[k] = (k < D) ? static_cast<T>(x_io[rid * D + k]) : T(0);
    T fx = ev.cost(xv, gp);
    ev.grad(xv, g);
    T norm_grad = M<T>::sqrt_(E::dot(g, g));
    T radius = static_cast<T>(opt.delta0);
    int it = 0, why = 0;

    while (true) {
        // ---- truncated CG on the model m(eta) = <g, eta> + 1/2 <eta, H eta> within |eta| <= radius ----
        T r_r = E::dot(g, g);
        const T norm_r0 = M<T>::sqrt_(r_r);
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            eta[k] = T(0);
            heta[k] = T(0);
            rv[k] = g[k];
            dl[k] = -g[k];
        }
        T e_pe = T(0), z_r = r_r, d_pd = r_r, e_pd = T(0), model_value = T(0);
        int stop = 4;   // MAX_INNER_ITER
        const T radius2 = radius * radius;
#pragma unroll 1
        for (int j = 0; j < opt.maxinner; ++j) {
            // H delta by finite differences of the gradient (approximate_hessian.py:30-62)
            const T norm_a = M<T>::sqrt_(E::dot(dl, dl));
            if (norm_a < T(1e-15)) {
#pragma unroll
                for (int k = 0; k < DP; ++k) hd[k] = T(0);
            } else {
                const T c = fd_eps / norm_a;
                T s = T(0);
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    xt[k] = fma(c, dl[k], xv[k]);
                    s = fma(xt[k], xt[k], s);
                }
                const T inv = T(1) / M<T>::sqrt_(s);
#pragma unroll
                for (int k = 0; k < DP; ++k) xt[k] *= inv;
                ev.cost(xt, gp);
                ev.grad(xt, hd);
                const T t = E::dot(xv, hd);                       // transport x1 -> x: projection onto T_x
#pragma unroll
                for (int k = 0; k < DP; ++k) hd[k] = fma(-t, xv[k], hd[k]) / c - g[k] / c;
            }
            const T d_hd = E::dot(dl, hd);
            T alpha = T(0), e_pe_new = e_pe;
            if (d_hd != T(0)) {
                alpha = z_r / d_hd;
                e_pe_new = e_pe + T(2) * alpha * e_pd + alpha * alpha * d_pd;
            }
            if (d_hd <= T(0) || e_pe_new >= radius2) {
                const T tau = (-e_pd + M<T>::sqrt_(e_pd * e_pd + d_pd * (radius2 - e_pe))) / d_pd;
#pragma unroll
                for (int k = 0; k < DP; ++k) {
                    eta[k] = fma(tau, dl[k], eta[k]);
                    heta[k] = fma(tau, hd[k], heta[k]);
                }
                stop = (d_hd <= T(0)) ? 0 : 1;                    // NEGATIVE_CURVATURE : EXCEEDED_TR
                break;
            }
            e_pe = e_pe_new;
            T m1 = T(0), m2 = T(0);
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const T ne = fma(alpha, dl[k], eta[k]), nh = fma(alpha, hd[k], heta[k]);
                m1 = fma(ne, g[k], m1);
                m2 = fma(ne, nh, m2);
            }
            const T new_model_value = m1 + T(0.5) * m2;
            if (new_model_value >= model_value) {
                stop = 5;                                         // MODEL_INCREASED
                break;
            }
            model_value = new_model_value;
            r_r = T(0);
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                eta[k] = fma(alpha, dl[k], eta[k]);
                heta[k] = fma(alpha, hd[k], heta[k]);
                rv[k] = fma(alpha, hd[k], rv[k]);
                r_r = fma(rv[k], rv[k], r_r);
            }
            const T norm_r = M<T>::sqrt_(r_r);
            const T pw = (theta == T(1)) ? norm_r0 : static_cast<T>(pow(static_cast<double>(norm_r0),
                                                                        static_cast<double>(theta)));
            if (j >= opt.mininner && norm_r <= norm_r0 * fmin(pw, kappa)) {
                stop = (kappa < pw) ? 2 : 3;                      // REACHED_TARGET_LINEAR : SUPERLINEAR
                break;
            }
            const T zold_rold = z_r;
            z_r = r_r;
            const T bcg = z_r / zold_rold;
#pragma unroll
            for (int k = 0; k < DP; ++k) dl[k] = fma(bcg, dl[k], -rv[k]);
            e_pd = bcg * (e_pd + alpha * d_pd);
            d_pd = z_r + bcg * bcg * d_pd;
        }
        // ---- proposal, ratio of actual to predicted decrease, radius update (robust_trust_regions.py:225-300) ----
        T s = T(0);
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            xt[k] = xv[k] + eta[k];
            s = fma(xt[k], xt[k], s);
        }
        const T inv = T(1) / M<T>::sqrt_(s);
#pragma unroll
        for (int k = 0; k < DP; ++k) xt[k] *= inv;
        const T fx_prop = ev.cost(xt, gp);
        const T rho_reg = fmax(T(1), fabs(fx)) * rho_scale;
        const T rhonum = (fx - fx_prop) + rho_reg;
        const T rhoden = (-E::dot(g, eta) - T(0.5) * E::dot(eta, heta)) + rho_reg;
        const bool model_decreased = rhoden >= T(0);
        const T rho = rhonum / rhoden;                            // NaN for 0/0, as np.nan in the reference
        if (rho < T(0.25) || !model_decreased || rho != rho) {
            radius = radius / T(4);
        } else if (rho > T(0.75) && (stop == 0 || stop == 1)) {
            radius = fmin(T(2) * radius, delta_bar);
        }
        if (model_decreased && rho > rho_prime) {
#pragma unroll
            for (int k = 0; k < DP; ++k) xv[k] = xt[k];
            fx = fx_prop;
            ev.grad(xv, g);                                       // the evaluator's state is that of cost(x_prop)
            norm_grad = M<T>::sqrt_(E::dot(g, g));
        }
        ++it;
        if (it >= opt.maxiter) { why = 1; break; }
        if (norm_grad < mingrad) { why = 2; break; }
    }
    if (lane == 0) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DP; ++k) s = fma(static_cast<double>(xv[k]), static_cast<double>(xv[k]), s);
        const double inv = 1.0 / sqrt(s);
#pragma unroll
        for (int k = 0; k < DP; ++k)
            if (k < D) x_io[rid * D + k] = static_cast<double>(xv[k]) * inv;
        value[rid] = static_cast<double>(-fx);
        if (iters) iters[rid] = it;
        if (reason) reason[rid] = why;
    }
}

int spec_width() {   // GABO_ACQ_SPEC = 1 | 2 | 4 (developer switch for the speculation width)
    static const int w = [] {
        const char* e = getenv("GABO_ACQ_SPEC");
        const int v = e ? atoi(e) : 0;
        return (v == 1 || v == 2 || v == 4) ? v : 0;
    }();
    return w;
}

template <typename T, int DP, int NCH, int kSpec>
int launch_rcg(const GpParams& gp, const RcgParams& opt, double* x, int64_t r, double* value, int32_t* iters,
               int32_t* reason, cudaStream_t stream) {
    const int n = gp.n, npad = (n + 3) & ~3;
    SmemCarver cv;
    cv.take(sizeof(T) * npad * n);
    cv.take(sizeof(T) * kSpec * SphereEval<T, DP, NCH>::ksh_size(n));
    cv.take(sizeof(T) * 2 * kSpec);
    cv.take(sizeof(T) * 2 * 2 * DP);
    const size_t smem = cv.off;
    auto kern = sphere_rcg_cta_kernel<T, DP, NCH, kSpec>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSpec * 32, smem);
    if (occ < 1) occ = 1;
    const unsigned grid = static_cast<unsigned>(imin(r, static_cast<int64_t>(sm_count()) * occ));
    kern<<<grid, kSpec * 32, smem, stream>>>(gp, opt, x, r, value, iters, reason);
    return check_launch("sphere_rcg_cta_kernel");
}

template <typename T, int DP, int NCH>
int launch_reg(const GpParams& gp, const RcgParams& opt, int mode, double* x, int64_t r, double* value, double* grad,
               int32_t* iters, int32_t* reason, cudaStream_t stream) {
    const int n = gp.n, npad = (n + 3) & ~3;
    if (mode == 0) {
        SmemCarver cv;
        cv.take(sizeof(T) * npad * n);
        cv.take(sizeof(T) * kAcqWarps * SphereEval<T, DP, NCH>::ksh_size(n));
        const size_t smem = cv.off;
        auto kern = sphere_ei_reg_kernel<T, DP, NCH>;
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        const unsigned grid = static_cast<unsigned>((r + kAcqWarps - 1) / kAcqWarps);
        kern<<<grid, kAcqWarps * 32, smem, stream>>>(gp, x, r, value, grad);
        return check_launch("sphere_ei_reg_kernel");
    }
    // Speculation width: the chip has 592 warp schedulers; with fewer restarts than that, idle schedulers are put to
    // work on speculative trial steps, beyond it the speculative work would compete with useful work.
    int spec = spec_width();
    if (spec == 0) spec = (r * 4 <= 592) ? 4 : ((r <= 1216) ? 2 : 1);   // measured on B200: R=1024 -> 2, R=4096 -> 1
    if (spec == 4) return launch_rcg<T, DP, NCH, 4>(gp, opt, x, r, value, iters, reason, stream);
    if (spec == 2) return launch_rcg<T, DP, NCH, 2>(gp, opt, x, r, value, iters, reason, stream);
    return launch_rcg<T, DP, NCH, 1>(gp, opt, x, r, value, iters, reason, stream);
}

template <typename T, int DP>
int launch_reg_n(const GpParams& gp, const RcgParams& opt, int mode, double* x, int64_t r, double* value, double* grad,
                 int32_t* iters, int32_t* reason, cudaStream_t stream) {
    if (gp.n <= 32) return launch_reg<T, DP, 1>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
    if (gp.n <= 64) return launch_reg<T, DP, 2>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
    if constexpr (DP <= 8) return launch_reg<T, DP, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
    return GABO_E_UNSUPPORTED;  // not reached: the dispatcher sends dim > 8 with n > 64 to the shared-memory kernel
}

template <typename T, int NCH>
int launch_t(const GpParams& gp, const RcgParams& opt, int mode, double* x, int64_t r, double* value, double* grad,
             int32_t* iters, int32_t* reason, cudaStream_t stream) {
    const int n = gp.n, D = gp.dim;
    const int Dp = D | 1, npad = (n + 3) & ~3, Dv = (D + 3) & ~3;
    SmemCarver cv;
    cv.take(sizeof(T) * n * Dp);
    cv.take(sizeof(T) * npad);
    cv.take(sizeof(T) * n * n);
    cv.take(sizeof(T) * kAcqWarps * (5 * Dv + 3 * npad));
    const size_t smem = cv.off;
    GABO_REQUIRE(smem <= 227 * 1024, GABO_E_UNSUPPORTED,
                 "sphere acquisition kernel: n_train=%d, dim=%d need %zu bytes of shared memory (> 227 KB)", n, D, smem);
    auto kern = sphere_acq_kernel<T, NCH>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const unsigned grid = static_cast<unsigned>((r + kAcqWarps - 1) / kAcqWarps);
    kern<<<grid, kAcqWarps * 32, smem, stream>>>(gp, opt, mode, x, r, value, grad, iters, reason);
    return check_launch("sphere_acq_kernel");
}

template <typename T, int DP, int NCH>
int launch_rtr(const GpParams& gp, const RtrParams& opt, double* x, int64_t r, double* value, int32_t* iters,
               int32_t* reason, cudaStream_t stream) {
    const int n = gp.n, npad = (n + 3) & ~3;
    SmemCarver cv;
    cv.take(sizeof(T) * npad * n);
    cv.take(sizeof(T) * kAcqWarps * SphereEval<T, DP, NCH>::ksh_size(n));
    const size_t smem = cv.off;
    auto kern = sphere_rtr_kernel<T, DP, NCH>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const unsigned grid = static_cast<unsigned>((r + kAcqWarps - 1) / kAcqWarps);
    kern<<<grid, kAcqWarps * 32, smem, stream>>>(gp, opt, x, r, value, iters, reason);
    return check_launch("sphere_rtr_kernel");
}

template <typename T, int DP>
int launch_rtr_n(const GpParams& gp, const RtrParams& opt, double* x, int64_t r, double* value, int32_t* iters,
                 int32_t* reason, cudaStream_t stream) {
    if (gp.n <= 32) return launch_rtr<T, DP, 1>(gp, opt, x, r, value, iters, reason, stream);
    if (gp.n <= 64) return launch_rtr<T, DP, 2>(gp, opt, x, r, value, iters, reason, stream);
    if constexpr (DP <= 8) return launch_rtr<T, DP, 4>(gp, opt, x, r, value, iters, reason, stream);
    return GABO_E_UNSUPPORTED;
}

}  // namespace

int launch_acq_sphere(const gabo_gp_desc* g, double* x, int64_t r, const gabo_rcg_opts* o, double* value, double* grad,
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
    if (gp.dim <= 8 || (gp.dim <= 16 && gp.n <= 64)) {  // register-resident kernel: dimension padded to 4, 8 or 16
        if (gp.dim <= 4)
            return f64 ? launch_reg_n<double, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
                       : launch_reg_n<float, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
        if (gp.dim <= 8)
            return f64 ? launch_reg_n<double, 8>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
                       : launch_reg_n<float, 8>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
        return f64 ? launch_reg_n<double, 16>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
                   : launch_reg_n<float, 16>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
    }
    if (gp.n <= 32) {
        return f64 ? launch_t<double, 1>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
                   : launch_t<float, 1>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
    }
    return f64 ? launch_t<double, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream)
               : launch_t<float, 4>(gp, opt, mode, x, r, value, grad, iters, reason, stream);
}

int launch_rtr_sphere(const gabo_gp_desc* g, double* x, int64_t r, const gabo_rtr_opts* o, double* value, int32_t* iters,
                      int32_t* reason, cudaStream_t stream) {
    GpParams gp{g->n_train, g->dim, g->mean, g->outputscale, g->beta, g->best_f, g->kxx, g->x_train, g->alpha, g->minv};
    RtrParams opt{};
    opt.maxiter = o->maxiter;
    opt.mininner = o->mininner;
    opt.maxinner = o->maxinner > 0 ? o->maxinner : g->dim - 1;                  // pymanopt Sphere.dim
    opt.mingradnorm = o->mingradnorm;
    opt.kappa = o->kappa;
    opt.theta = o->theta;
    opt.rho_prime = o->rho_prime;
    opt.rho_regularization = o->rho_regularization;
    opt.delta_bar = o->delta_bar > 0.0 ? o->delta_bar : 3.14159265358979323846; // pymanopt Sphere.typicaldist
    opt.delta0 = o->delta0 > 0.0 ? o->delta0 : opt.delta_bar / 8.0;
    opt.fd_eps = 1.0 / 16384.0;                                                 // approximate_hessian.py:43
    const bool f64 = g->compute == GABO_F64;
    GABO_REQUIRE(gp.dim <= 8 || (gp.dim <= 16 && gp.n <= 64), GABO_E_UNSUPPORTED,
                 "gabo_acq_rtr: the trust-region kernel keeps the iterate in registers: ambient dimension <= 8, or "
                 "<= 16 with n_train <= 64 (got dim=%d, n_train=%d)", gp.dim, gp.n);
    if (gp.dim <= 4)
        return f64 ? launch_rtr_n<double, 4>(gp, opt, x, r, value, iters, reason, stream)
                   : launch_rtr_n<float, 4>(gp, opt, x, r, value, iters, reason, stream);
    if (gp.dim <= 8)
        return f64 ? launch_rtr_n<double, 8>(gp, opt, x, r, value, iters, reason, stream)
                   : launch_rtr_n<float, 8>(gp, opt, x, r, value, iters, reason, stream);
    return f64 ? launch_rtr_n<double, 16>(gp, opt, x, r, value, iters, reason, stream)
               : launch_rtr_n<float, 16>(gp, opt, x, r, value, iters, reason, stream);
}

}  // namespace gabo
