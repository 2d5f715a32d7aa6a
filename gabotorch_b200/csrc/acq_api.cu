// C-ABI entry points of the acquisition path: gabo_ei_eval (A4), gabo_acq_rcg / gabo_acq_rtr (A1), gabo_argmax_records (A3).
#include "acq_common.cuh"

namespace gabo {
namespace {

int validate_gp(const gabo_gp_desc* gp, const char* who) {
    GABO_REQUIRE(gp, GABO_E_ARG, "%s: null GP description", who);
    GABO_REQUIRE(gp->manifold == GABO_SPHERE || gp->manifold == GABO_SPD, GABO_E_ARG, "%s: unknown manifold %d", who,
                 gp->manifold);
    GABO_REQUIRE(gp->n_train >= 1 && gp->n_train <= GABO_MAX_TRAIN, GABO_E_ARG, "%s: n_train=%d outside [1, %d]", who,
                 gp->n_train, GABO_MAX_TRAIN);
    GABO_REQUIRE(gp->x_train && gp->alpha && gp->minv, GABO_E_ARG, "%s: null GP array", who);
    GABO_REQUIRE(gp->compute == GABO_F32 || gp->compute == GABO_F64, GABO_E_ARG, "%s: bad compute dtype", who);
    if (gp->manifold == GABO_SPHERE) {
        GABO_REQUIRE(gp->dim >= 2 && gp->dim <= GABO_MAX_SPHERE_DIM, GABO_E_ARG, "%s: sphere dim %d outside [2, %d]",
                     who, gp->dim, GABO_MAX_SPHERE_DIM);
    } else {
        GABO_REQUIRE(gp->dim >= 1 && gp->dim <= GABO_MAX_SPD_DIM, GABO_E_ARG, "%s: SPD size %d outside [1, %d]", who,
                     gp->dim, GABO_MAX_SPD_DIM);
    }
    return GABO_OK;
}

int dispatch(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rcg_opts* opts, double* value, double* grad,
             int32_t* iters, int32_t* reason, cudaStream_t s) {
    if (gp->manifold == GABO_SPHERE) return launch_acq_sphere(gp, x, r, opts, value, grad, iters, reason, s);
    switch (gp->dim) {
#define GABO_CASE(DD) \
    case DD:          \
        return launch_acq_spd<DD>(gp, x, r, opts, value, grad, iters, reason, s);
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

int dispatch_ctr(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_ctr_opts* o, double* value, int32_t* iters,
                 int32_t* reason, cudaStream_t s) {
    const int d = gp->dim;
    CtrParams p{};
    p.tr.maxiter = o->tr.maxiter;
    p.tr.mininner = o->tr.mininner;
    p.tr.maxinner = o->tr.maxinner > 0 ? o->tr.maxinner : d * (d + 1) / 2;               // pymanopt PositiveDefinite.dim
    p.tr.mingradnorm = o->tr.mingradnorm;
    p.tr.kappa = o->tr.kappa;
    p.tr.theta = o->tr.theta;
    p.tr.rho_prime = o->tr.rho_prime;
    p.tr.rho_regularization = o->tr.rho_regularization;
    p.tr.delta_bar = o->tr.delta_bar > 0.0 ? o->tr.delta_bar : sqrt(0.5 * d * (d + 1));   // PositiveDefinite.typicaldist
    p.tr.delta0 = o->tr.delta0 > 0.0 ? o->tr.delta0 : p.tr.delta_bar / 8.0;
    p.tr.fd_eps = 1.0 / 16384.0;                                                         // approximate_hessian.py:43
    p.n_cons = o->n_constraints;
    p.strict = o->strict;
    for (int c = 0; c < 2; ++c) {
        p.kind[c] = o->kind[c];
        p.bound[c] = o->bound[c];
    }
    p.delta_cons = o->delta_cons;
    switch (d) {
#define GABO_CASE(DD) \
    case DD:          \
        return launch_rtr_spd<DD>(gp, x, r, p, value, iters, reason, s);
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

// Lexicographic (value desc, global index asc), NaN = -inf.  One CTA; n is the number of restarts (small).
__global__ void argmax_records_kernel(const double* __restrict__ values, const int64_t* __restrict__ gidx, int64_t n,
                                      int64_t* __restrict__ out_slot, double* __restrict__ out_value) {
    __shared__ double sv[32];
    __shared__ int64_t sg[32];
    __shared__ int64_t ss[32];
    const double ninf = __longlong_as_double(0xfff0000000000000LL);
    double bv = ninf;
    int64_t bg = INT64_MAX, bs = INT64_MAX;
    auto better = [](double v, int64_t g, int64_t s, double v2, int64_t g2, int64_t s2) {
        return v > v2 || (v == v2 && (g < g2 || (g == g2 && s < s2)));
    };
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double v = values[i];
        if (!(v == v)) v = ninf;
        const int64_t g = gidx ? gidx[i] : i;
        if (better(v, g, i, bv, bg, bs)) {
            bv = v;
            bg = g;
            bs = i;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
        const int64_t g2 = __shfl_xor_sync(0xffffffffu, bg, o);
        const int64_t s2 = __shfl_xor_sync(0xffffffffu, bs, o);
        if (better(v2, g2, s2, bv, bg, bs)) {
            bv = v2;
            bg = g2;
            bs = s2;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sv[warp] = bv;
        sg[warp] = bg;
        ss[warp] = bs;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        bv = lane < nw ? sv[lane] : ninf;
        bg = lane < nw ? sg[lane] : INT64_MAX;
        bs = lane < nw ? ss[lane] : INT64_MAX;
        for (int o = 16; o > 0; o >>= 1) {
            const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
            const int64_t g2 = __shfl_xor_sync(0xffffffffu, bg, o);
            const int64_t s2 = __shfl_xor_sync(0xffffffffu, bs, o);
            if (better(v2, g2, s2, bv, bg, bs)) {
                bv = v2;
                bg = g2;
                bs = s2;
            }
        }
        if (lane == 0) {
            out_slot[0] = bs;
            if (out_value) out_value[0] = bv;
        }
    }
}

}  // namespace
}  // namespace gabo

extern "C" int gabo_ei_eval(const gabo_gp_desc* gp, const double* x, int64_t r, double* ei, double* grad,
                            void* stream) {
    using namespace gabo;
    const int rc = validate_gp(gp, "gabo_ei_eval");
    if (rc != GABO_OK) return rc;
    GABO_REQUIRE(r >= 0, GABO_E_ARG, "gabo_ei_eval: negative size");
    if (r == 0) return GABO_OK;
    GABO_REQUIRE(x && ei, GABO_E_ARG, "gabo_ei_eval: null pointer");
    return dispatch(gp, const_cast<double*>(x), r, nullptr, ei, grad, nullptr, nullptr,
                    static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_acq_rcg(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rcg_opts* opts, double* value,
                            int32_t* iters, int32_t* reason, void* stream) {
    using namespace gabo;
    const int rc = validate_gp(gp, "gabo_acq_rcg");
    if (rc != GABO_OK) return rc;
    GABO_REQUIRE(r >= 0, GABO_E_ARG, "gabo_acq_rcg: negative size");
    if (r == 0) return GABO_OK;
    GABO_REQUIRE(x && value && opts, GABO_E_ARG, "gabo_acq_rcg: null pointer");
    GABO_REQUIRE(opts->maxiter >= 1 && opts->ls_maxiter >= 0, GABO_E_ARG, "gabo_acq_rcg: bad iteration limits");
    GABO_REQUIRE(opts->contraction > 0.0 && opts->contraction < 1.0, GABO_E_ARG, "gabo_acq_rcg: bad contraction factor");
    return dispatch(gp, x, r, opts, value, nullptr, iters, reason, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_acq_rtr(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rtr_opts* opts, double* value,
                            int32_t* iters, int32_t* reason, void* stream) {
    using namespace gabo;
    const int rc = validate_gp(gp, "gabo_acq_rtr");
    if (rc != GABO_OK) return rc;
    GABO_REQUIRE(r >= 0, GABO_E_ARG, "gabo_acq_rtr: negative size");
    if (r == 0) return GABO_OK;
    GABO_REQUIRE(x && value && opts, GABO_E_ARG, "gabo_acq_rtr: null pointer");
    GABO_REQUIRE(opts->maxiter >= 1 && opts->mininner >= 0, GABO_E_ARG, "gabo_acq_rtr: bad iteration limits");
    GABO_REQUIRE(opts->kappa > 0.0 && opts->rho_prime >= 0.0 && opts->rho_prime < 0.25, GABO_E_ARG,
                 "gabo_acq_rtr: need kappa > 0 and 0 <= rho_prime < 1/4");
    if (gp->manifold == GABO_SPD) {
        gabo_ctr_opts c{};
        c.tr = *opts;
        c.delta_cons = 1e-6;
        return dispatch_ctr(gp, x, r, &c, value, iters, reason, static_cast<cudaStream_t>(stream));
    }
    return launch_rtr_sphere(gp, x, r, opts, value, iters, reason, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_acq_ctr(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_ctr_opts* opts, double* value,
                            int32_t* iters, int32_t* reason, void* stream) {
    using namespace gabo;
    const int rc = validate_gp(gp, "gabo_acq_ctr");
    if (rc != GABO_OK) return rc;
    GABO_REQUIRE(r >= 0, GABO_E_ARG, "gabo_acq_ctr: negative size");
    if (r == 0) return GABO_OK;
    GABO_REQUIRE(x && value && opts, GABO_E_ARG, "gabo_acq_ctr: null pointer");
    GABO_REQUIRE(gp->manifold == GABO_SPD, GABO_E_UNSUPPORTED,
                 "gabo_acq_ctr: the constrained trust-region kernel is implemented for SPD(d)");
    GABO_REQUIRE(opts->tr.maxiter >= 1 && opts->tr.mininner >= 0, GABO_E_ARG, "gabo_acq_ctr: bad iteration limits");
    GABO_REQUIRE(opts->tr.kappa > 0.0 && opts->tr.rho_prime >= 0.0 && opts->tr.rho_prime < 0.25, GABO_E_ARG,
                 "gabo_acq_ctr: need kappa > 0 and 0 <= rho_prime < 1/4");
    GABO_REQUIRE(opts->n_constraints >= 0 && opts->n_constraints <= 2, GABO_E_ARG,
                 "gabo_acq_ctr: n_constraints=%d outside [0, 2]", opts->n_constraints);
    for (int c = 0; c < opts->n_constraints; ++c)
        GABO_REQUIRE(opts->kind[c] == GABO_CONS_MAX_EIG || opts->kind[c] == GABO_CONS_MIN_EIG, GABO_E_ARG,
                     "gabo_acq_ctr: unknown constraint kind %d", opts->kind[c]);
    GABO_REQUIRE(opts->delta_cons > 0.0, GABO_E_ARG, "gabo_acq_ctr: delta_cons must be positive");
    return dispatch_ctr(gp, x, r, opts, value, iters, reason, static_cast<cudaStream_t>(stream));
}

extern "C" int gabo_argmax_records(const double* values, const int64_t* gidx, int64_t n, int64_t* out_slot,
                                   double* out_value, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 1, GABO_E_ARG, "gabo_argmax_records: need at least one record");
    GABO_REQUIRE(values && out_slot, GABO_E_ARG, "gabo_argmax_records: null pointer");
    argmax_records_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(values, gidx, n, out_slot, out_value);
    return check_launch("argmax_records_kernel");
}
