// Acquisition value / multi-start Riemannian CG on the sphere (A1 + A4 with M1 primitives).
// One warp per restart, all CG steps inside one launch; see acq_common.cuh for the scheme.
//
// Manifold operations used by the solver are pymanopt's Sphere (reference call sites manifold_optimize.py:207-221,
// numpy statements Riemannian_utils/sphere_utils.py:14-123): retr(x,u) = (x+u)/|x+u|, transp(x,y,u) = u - <y,u> y,
// inner = Euclidean dot, Log_x(y) = (y - <x,y> x) * theta / |y - <x,y> x|.
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
template <typename T, int DP, int NCH>
__global__ void __launch_bounds__(kAcqWarps * 32)
    sphere_acq_reg_kernel(GpParams gp, RcgParams opt, int mode, double* __restrict__ x_io, int64_t r,
                          double* __restrict__ value, double* __restrict__ grad_out, int32_t* __restrict__ iters,
                          int32_t* __restrict__ reason) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = gp.n, D = gp.dim;
    const int npad = (n + 3) & ~3;
    SmemCarver cv;
    T* Minv = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * npad * n));       // rows n..npad-1 are zero
    T* kbase = reinterpret_cast<T*>(smem_raw + cv.take(sizeof(T) * kAcqWarps * npad));
    for (int e = threadIdx.x; e < npad * n; e += blockDim.x) Minv[e] = (e < n * n) ? static_cast<T>(gp.minv[e]) : T(0);
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t rid = static_cast<int64_t>(blockIdx.x) * kAcqWarps + warp;
    if (rid >= r) return;
    T* ksh = kbase + warp * npad;

    // training points owned by this lane
    T Xi[NCH][DP], al[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
        const int i = lane + 32 * ch;
        al[ch] = (i < n) ? static_cast<T>(gp.alpha[i]) : T(0);
#pragma unroll
        for (int k = 0; k < DP; ++k) Xi[ch][k] = (i < n && k < D) ? static_cast<T>(gp.x_train[i * D + k]) : T(0);
    }
    const T s_out = static_cast<T>(gp.outputscale), beta = static_cast<T>(gp.beta);
    T c_l[NCH], th_l[NCH], k_l[NCH], mk_l[NCH];
    EiScalars<T> sc;

    auto dot = [](const T (&a)[DP], const T (&b)[DP]) {
        T s = a[0] * b[0];
#pragma unroll
        for (int k = 1; k < DP; ++k) s = fma(a[k], b[k], s);
        return s;
    };

    auto cost_at = [&](const T (&p)[DP]) -> T {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const int i = lane + 32 * ch;
            T cc = dot(Xi[ch], p);
            cc = fmin(fmax(cc, T(-1) + M<T>::clamp_eps()), T(1) - M<T>::clamp_eps());
            const T tt = M<T>::acos_(cc);
            const T kk = (i < n) ? s_out * M<T>::exp_(-beta * tt * tt) : T(0);
            if (i < npad) ksh[i] = kk;
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
                T m0 = T(0), m1 = T(0), m2 = T(0), m3 = T(0);
                for (int j = 0; j < npad; j += 4) {     // K^-1 is symmetric: row i read as column i, conflict-free
                    m0 = fma(Minv[j * n + i], ksh[j], m0);
                    m1 = fma(Minv[(j + 1) * n + i], ksh[j + 1], m1);
                    m2 = fma(Minv[(j + 2) * n + i], ksh[j + 2], m2);
                    m3 = fma(Minv[(j + 3) * n + i], ksh[j + 3], m3);
                }
                mk = (m0 + m1) + (m2 + m3);
            }
            mk_l[ch] = mk;
            ka = fma(k_l[ch], al[ch], ka);
            kmk = fma(k_l[ch], mk, kmk);
        }
        __syncwarp();   // ksh is rewritten by the next call
        ka = warp_sum(ka);
        kmk = warp_sum(kmk);
        sc = ei_scalars<T>(ka, kmk, gp);
        const T cst = -sc.ei;
        return (cst == cst) ? cst : M<T>::inf();
    };

    // Riemannian gradient of the cost at p (the point of the last cost_at call)
    auto grad_at = [&](const T (&p)[DP], T (&out)[DP]) {
        T acc[DP], sgc = T(0);
#pragma unroll
        for (int k = 0; k < DP; ++k) acc[k] = T(0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            const T w = -sc.cdf * al[ch] - sc.pdf_over_sigma * mk_l[ch];
            const T coef = T(2) * beta * w * k_l[ch];          // k_l = 0 for lanes without a training point
            T pn = T(0);
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                const T pk = fma(-c_l[ch], p[k], Xi[ch][k]);
                pn = fma(pk, pk, pn);
            }
            pn = M<T>::sqrt_(pn);
            const T scale = (th_l[ch] > T(1e-6)) ? th_l[ch] / (pn > T(0) ? pn : T(1)) : T(1);
            const T g = coef * scale;
#pragma unroll
            for (int k = 0; k < DP; ++k) acc[k] = fma(g, Xi[ch][k], acc[k]);
            sgc = fma(g, c_l[ch], sgc);
        }
#pragma unroll
        for (int k = 0; k < DP; ++k) acc[k] = warp_sum(acc[k]);
        sgc = warp_sum(sgc);
#pragma unroll
        for (int k = 0; k < DP; ++k) out[k] = -(acc[k] - sgc * p[k]);  // cost = -EI
    };

    T xv[DP], gv[DP], eta[DP], xn[DP], gn[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) xv[k] = (k < D) ? static_cast<T>(x_io[rid * D + k]) : T(0);

    T cost = cost_at(xv);
    if (mode == 0) {
        if (lane == 0) value[rid] = static_cast<double>(-cost);
        if (grad_out) {
            grad_at(xv, gv);
#pragma unroll
            for (int k = 0; k < DP; ++k)
                if (k < D && lane == 0) grad_out[rid * D + k] = static_cast<double>(-gv[k]);
        }
        return;
    }

    grad_at(xv, gv);
    T gPg = dot(gv, gv);
    T gradnorm = M<T>::sqrt_(gPg);
#pragma unroll
    for (int k = 0; k < DP; ++k) eta[k] = -gv[k];
    int it = 0, why = 0;
    T stepsize = M<T>::nan();
    T oldalpha = T(-1);  // unset
    const T mingrad = static_cast<T>(opt.mingradnorm), minstep = static_cast<T>(opt.minstepsize);
    const T contraction = static_cast<T>(opt.contraction), suff = static_cast<T>(opt.suff_decr);

    auto retract = [&](T a) {  // xn = (x + a eta) / |x + a eta|
        T s = T(0);
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            xn[k] = fma(a, eta[k], xv[k]);
            s = fma(xn[k], xn[k], s);
        }
        const T inv = T(1) / M<T>::sqrt_(s);
#pragma unroll
        for (int k = 0; k < DP; ++k) xn[k] *= inv;
    };

    while (true) {
        if (it + 1 >= opt.maxiter) { why = 1; break; }
        if (gradnorm < mingrad) { why = 2; break; }
        if (stepsize < minstep) { why = 3; break; }
        T df0 = dot(gv, eta);
        if (df0 >= T(0)) {  // not a descent direction: restart from steepest descent
#pragma unroll
            for (int k = 0; k < DP; ++k) eta[k] = -gv[k];
            df0 = -gPg;
        }
        const T norm_d = M<T>::sqrt_(dot(eta, eta));
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
#pragma unroll
            for (int k = 0; k < DP; ++k) {
                xn[k] = xv[k];
                gn[k] = gv[k];
            }
            newf = cost;
        } else {
            grad_at(xn, gn);
        }
        stepsize = a * norm_d;
        oldalpha = (evals == 2) ? a : T(2) * a;
        // transport g and eta to xn (projection), Hestenes-Stiefel beta
        const T xg = dot(xn, gv), xe = dot(xn, eta);
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
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < DP; ++k) s = fma(static_cast<double>(xv[k]), static_cast<double>(xv[k]), s);
    const double inv = 1.0 / sqrt(s);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < DP; ++k)
            if (k < D) x_io[rid * D + k] = static_cast<double>(xv[k]) * inv;
        value[rid] = static_cast<double>(-cost);
        if (iters) iters[rid] = it;
        if (reason) reason[rid] = why;
    }
}

template <typename T, int DP, int NCH>
int launch_reg(const GpParams& gp, const RcgParams& opt, int mode, double* x, int64_t r, double* value, double* grad,
               int32_t* iters, int32_t* reason, cudaStream_t stream) {
    const int n = gp.n, npad = (n + 3) & ~3;
    SmemCarver cv;
    cv.take(sizeof(T) * npad * n);
    cv.take(sizeof(T) * kAcqWarps * npad);
    const size_t smem = cv.off;
    auto kern = sphere_acq_reg_kernel<T, DP, NCH>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    const unsigned grid = static_cast<unsigned>((r + kAcqWarps - 1) / kAcqWarps);
    kern<<<grid, kAcqWarps * 32, smem, stream>>>(gp, opt, mode, x, r, value, grad, iters, reason);
    return check_launch("sphere_acq_reg_kernel");
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

}  // namespace gabo
