/*
 * gabo_b200.h -- C ABI of libgabo_b200.so: the B200 (sm_100a) implementation of GaBOtorch's data-parallel hot path.
 *
 * Drop-in boundary.  The reference (NoemieJaquier/GaBOtorch @ 884f64a) is pure Python and has
 * no FFI; each entry point below replaces the arithmetic of the reference function cited next to it, and the Python
 * package `gabotorch_b200` binds them with ctypes behind the reference's own class / function names
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless stated otherwise;
 *   - row-major, contiguous, base pointers 16-byte aligned;
 *   - the caller owns every buffer (inputs, outputs, scratch); the library allocates nothing persistent;
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as void*), no device synchronisation;
 *   - return value 0 = ok, < 0 = error (GABO_E_*); gabo_last_error() gives a thread-local message;
 *   - points are fp64 (the reference's dtype).  O(N) per-point work runs in fp64; O(N^2) / O(R*T) per-pair and
 *     per-step work runs in the type selected by `compute` (GABO_F32 default, GABO_F64 reference-grade).
 */
#ifndef GABO_B200_H
#define GABO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GABO_VERSION 100 /* 0.1.0 */

/* error codes */
#define GABO_OK 0
#define GABO_E_ARG (-1)    /* null pointer, bad size, unsupported dimension */
#define GABO_E_ALIGN (-2)  /* base pointer not 16-byte aligned */
#define GABO_E_CUDA (-3)   /* CUDA runtime error at launch */
#define GABO_E_UNSUPPORTED (-4)

/* dtypes */
#define GABO_F32 0
#define GABO_F64 1

/* kernel kinds: what is written for each pair, given the geodesic distance d and the scalar `param` */
#define GABO_KIND_GAUSS 0    /* exp(-param * d^2)   kernels_sphere.py:93, kernels_spd.py:98   (param = beta)          */
#define GABO_KIND_LAPLACE 1  /* exp(-param * d)     kernels_sphere.py:132 (param = 1/l^2), kernels_spd.py:185 (beta) */
#define GABO_KIND_DIST 2     /* d itself            sphere_utils_torch.py:55, spd_utils_torch.py:120                 */

/* manifolds */
#define GABO_SPHERE 0
#define GABO_SPD 1

#define GABO_MAX_SPHERE_DIM 128 /* ambient dimension D of S^{D-1}                          */
#define GABO_MAX_SPD_DIM 8      /* matrix size d of SPD(d): Jacobi state lives in registers */
#define GABO_MAX_TRAIN 128      /* GP training points held in shared memory by the optimiser */

int gabo_version(void);
const char* gabo_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * G1 + G2: fused sphere Gram.  Replaces sphere_distance_torch (BoManifolds/Riemannian_utils/sphere_utils_torch.py:12-55)
 * followed by SphereGaussianKernel.forward / SphereLaplaceKernel.forward (kernel_utils/kernels_sphere.py:89-94, 129-134):
 *   out[i, j] = f(acos(clamp(<x1_i, x2_j>, -1+1e-15, 1-1e-15)))      x1: n1 x dim, x2: n2 x dim (fp64)
 * out: n1 x n2 with row stride ld_out (elements), dtype out_dtype.
 * ------------------------------------------------------------------------------------------------------------------ */
int gabo_sphere_gram(const double* x1, int64_t n1, const double* x2, int64_t n2, int dim, double param, int kind,
                     void* out, int out_dtype, int64_t ld_out, void* stream);

/* diag=True branch (sphere_utils_torch.py:45-49): out[i] = f(d(x1_i, x2_i)), n values. */
int gabo_sphere_gram_diag(const double* x1, const double* x2, int64_t n, int dim, double param, int kind, void* out,
                          int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * G3: Mandel notation.  Replaces vector_to_symmetric_matrix_mandel_torch / symmetric_matrix_to_vector_mandel_torch
 * (Riemannian_utils/spd_utils_torch.py:159-194 / :197-226).  vec: n x d(d+1)/2, mat: n x d x d (fp64).
 * ------------------------------------------------------------------------------------------------------------------ */
int gabo_mandel_unpack(const double* vec, int64_t n, int d, double* mat, void* stream);
int gabo_mandel_pack(const double* mat, int64_t n, int d, double* vec, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * G4 + G5: SPD affine-invariant Gram.  Replaces affine_invariant_distance_torch (spd_utils_torch.py:53-120) and
 * SpdAffineInvariantGaussianKernel.forward / ...LaplaceKernel.forward (kernel_utils/kernels_spd.py:90-100, 178-187).
 *
 * Step 1 (per point, fp64): Mandel unpack (or plain matrices), Cholesky X = L L^T, A = L^-1.  `fac` receives, per point,
 *   gabo_spd_factor_stride(d) doubles: [ L packed lower-triangular row-major | A packed the same way | padding ].
 *   `flags` (nullable, one int32): bit 0 is set when any point is not positive definite (torch.cholesky raises at
 *   spd_utils_torch.py:87; the library reports instead, the Python layer raises).
 * Step 2 (per pair): G = A_i L_j (fp64 FMAs), one-sided Jacobi on G in `compute` precision, eigenvalues
 *   lambda_k = |g_k|^2 rounded to fp32, d = sqrt(sum log(lambda)^2 + 1e-15) in fp32 exactly as spd_utils_torch.py:108-120.
 *   `symmetric` != 0 (only when fac1 == fac2): computes the upper triangle of tiles and mirrors it.
 * ------------------------------------------------------------------------------------------------------------------ */
int64_t gabo_spd_factor_stride(int d);
int gabo_spd_factor(const double* x, int64_t n, int d, int input_is_mandel, double* fac, int32_t* flags, void* stream);
/* Both operands of a Gram build in ONE launch (either set may be empty). */
int gabo_spd_factor2(const double* x1, int64_t n1, const double* x2, int64_t n2, int d, int input_is_mandel, double* fac1,
                     double* fac2, int32_t* flags, void* stream);
int gabo_spd_ai_gram(const double* fac1, int64_t n1, const double* fac2, int64_t n2, int d, double param, int kind,
                     int compute, int symmetric, void* out, int out_dtype, int64_t ld_out, void* stream);

/* Input gradient of the affine-invariant Gram (what the reference gets from torch.autograd through
 * affine_invariant_distance_torch, spd_utils_torch.py:53-120).  w: upstream weights dLoss/d(d_ij^2), n1 x n2 with row
 * stride ld_w (transpose_w != 0: w is stored n2 x n1 and read transposed).  out: n1 x d x d fp64,
 *   out_i = sum_j w_ij grad_{X1_i} d^2(X1_i, X2_j) = -2 A_i^T (sum_j w_ij logm(A_i X2_j A_i^T)) A_i;
 * its Mandel vector (gabo_mandel_pack) is the gradient with respect to the Mandel input.  The gradient with respect
 * to the second operand is the same call with the operands swapped and transpose_w = 1. */
int gabo_spd_ai_gram_backward(const double* fac1, int64_t n1, const double* fac2, int64_t n2, int d, const double* w,
                              int64_t ld_w, int transpose_w, int compute, double* out, void* stream);

/* Backward building blocks of the other kernels (every reference kernel is differentiable under torch.autograd with
 * respect to its inputs and its manifold-valued parameters: kernels_sphere.py:71-134, kernels_spd.py:190-313,
 * kernels_nested_spd.py:104-246).
 *   gabo_weighted_points_sum: out_i = sum_j W_ij b_j (out: rows x k, b: cols x k, k <= 128), the reduction a pairwise
 *     distance backward ends in.  g (and dist) are n1 x n2 with row stride ld; transpose != 0 reads them transposed
 *     (rows = n2: the gradient of the SECOND operand).  mode 0: W = g.  mode 1: W = -g / sin(dist), 0 where the
 *     reference's clamp of the inner product is active (d/dx_i acos(clamp(<x_i, y_j>)), sphere_utils_torch.py:49-55).
 *   gabo_spd_logm_backward: grad_in_n = adjoint of the Frechet derivative of logm at mat_n applied to grad_out_n
 *     (n x d x d each; what autograd gives through logm_torch, spd_utils_torch.py:13-30).
 *   gabo_nested_spd_project_backward: Y_n = W^T X_n W (nested_spd_utils.py:13-48):
 *     grad_x_n = W sym(G_n) W^T (n x D x D, nullable), grad_w = 2 sum_n sym(X_n) W sym(G_n) (D x d, nullable). */
int gabo_weighted_points_sum(const double* g, const double* dist, int64_t n1, int64_t n2, int64_t ld, int transpose,
                             int mode, const double* b, int k, double* out, void* stream);
int gabo_spd_logm_backward(const double* mat, const double* grad_out, int64_t n, int d, double* grad_in, void* stream);
int gabo_nested_spd_project_backward(const double* x, const double* w, const double* grad_y, int64_t n, int D, int d,
                                     double* grad_x, double* grad_w, void* stream);

/* Frobenius / log-Euclidean Gram (spd_utils_torch.py:124-156, kernels_spd.py:230-241, 283-313):
 *   out[i,j] = f(|| M1_i - M2_j + 1e-15 ||_F), m: n x d x d fp64 (apply gabo_spd_logm first for log-Euclidean);
 *   kind GAUSS uses exp(-param d^2) with param = 1/lengthscale^2. */
int gabo_frobenius_gram(const double* m1, int64_t n1, const double* m2, int64_t n2, int d, double param, int kind,
                        void* out, int out_dtype, int64_t ld_out, void* stream);
/* logm_torch (spd_utils_torch.py:13-30) batched: out_i = V diag(log lambda) V^T. */
int gabo_spd_logm(const double* mat, int64_t n, int d, double* out, void* stream);
/* Batched symmetric eigendecomposition for matrices beyond the register kernels (1 <= D <= 32), fp64: mats n x D x D
 * (symmetrised on load), evals n x D, evecs n x D x D with column k the eigenvector of evals[k] (nullable: values only),
 * flags (nullable) gets bit 0 when a matrix is not finite or Jacobi did not converge.  One launch replaces the per-matrix
 * torch.symeig calls of affine_invariant_distance_torch / logm_torch / sqrtm_torch (spd_utils_torch.py:108-112, :25-30,
 * :45-50) inside the reconstruction costs of nested_mappings/nested_spd_optimization.py:22-92. */
int gabo_sym_eig(const double* mats, int64_t n, int D, double* evals, double* evecs, int32_t* flags, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * M1: batched sphere manifold operations (pymanopt Sphere; reference formulas Riemannian_utils/sphere_utils.py:14-123).
 * All arrays n x dim fp64.  op codes below; `y` is the second point / vector where the op takes one.
 * ------------------------------------------------------------------------------------------------------------------ */
#define GABO_OP_PROJ 0      /* out = u - <x,u> x                       (proj / egrad2rgrad; a = x, b = u)      */
#define GABO_OP_RETR 1      /* out = (x+u)/|x+u|                       (a = x, b = u)                          */
#define GABO_OP_EXP 2       /* out = x cos|u| + u sin|u|/|u|           (a = x, b = u)  sphere_utils.py:33-36   */
#define GABO_OP_LOG 3       /* out = Log_x(y)                          (a = x, b = y)  sphere_utils.py:60-63   */
#define GABO_OP_TRANSP 4    /* out = proj(y, u)                        (a = y, b = u)  pymanopt Sphere.transp  */
#define GABO_OP_PTRANSP 5   /* true parallel transport x->y of u       (a = x, b = y, c = u) sphere_utils.py:93-123 / spd_utils.py:200-213 */
#define GABO_OP_EGRAD2RGRAD 6 /* spd: X sym(G) X                       (a = X, b = G)                          */
int gabo_sphere_op(int op, const double* a, const double* b, const double* c, int64_t n, int dim, double* out,
                   void* stream);
/* out[i] = acos(clip(<x_i,y_i>, -1, 1))  (pymanopt Sphere.dist; sphere_utils.py:68-90) */
int gabo_sphere_dist(const double* x, const double* y, int64_t n, int dim, double* out, void* stream);

/* Nested-sphere projection chain of HD-GaBO on spheres (nested_mappings/nested_spheres_utils.py:120-147 with
 * rotation_from_sphere_points_torch, sphere_utils_torch.py:58-93):  x: n x D unit vectors -> y: n x d_latent unit vectors.
 * axes: the unit axes of the levels D, D-1, ..., d_latent+1 concatenated (D + (D-1) + ... values); dists: one distance
 * to the axis per level (the kernel of kernels_nested_sphere.py fixes them to pi/2).  fp64, D <= 64. */
int gabo_nested_sphere_project(const double* x, int64_t n, int D, int d_latent, const double* axes, const double* dists,
                               double* y, void* stream);
/* Every level of the same chain (projection_from_sphere_to_subsphere returns the list, nested_spheres_utils.py:139-147):
 * levels = [n x (D-1) | n x (D-2) | ... | n x d_latent] blocks, level-major, each row-major. */
int gabo_nested_sphere_chain(const double* x, int64_t n, int D, int d_latent, const double* axes, const double* dists,
                             double* levels, void* stream);
/* projection_from_sphere_to_nested_sphere (nested_spheres_utils.py:13-67): x: n x dim on S^{dim-1} -> y: n x dim, the
 * closest points at geodesic distance `dist` from `axis` (still in the coordinates of S^{dim-1}). */
int gabo_nested_sphere_to_nested(const double* x, int64_t n, int dim, const double* axis, double dist, double* y,
                                 void* stream);
/* Inverse chain, projection_from_subsphere_to_sphere (nested_spheres_utils.py:149-213): y: n x d_latent ->
 * levels = [n x (d_latent+1) | ... | n x D]; axes / dists in the order of the projection (levels D, D-1, ...). */
int gabo_nested_sphere_reconstruct(const double* y, int64_t n, int d_latent, int D, const double* axes,
                                   const double* dists, double* levels, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * M2: batched SPD manifold operations under the affine-invariant metric (pymanopt PositiveDefinite; reference formulas
 * Riemannian_utils/spd_utils.py:104-213).  Matrices n x d x d fp64, d <= GABO_MAX_SPD_DIM.
 *   EXP/RETR: L expm(L^-1 U L^-T) L^T;  LOG: L logm(L^-1 Y L^-T) L^T;  EGRAD2RGRAD: X sym(G) X;
 *   PTRANSP: E U E^T, E = (Y X^-1)^(1/2);  TRANSP: identity (pymanopt 0.2.x), provided for completeness.
 * ------------------------------------------------------------------------------------------------------------------ */
int gabo_spd_op(int op, const double* a, const double* b, const double* c, int64_t n, int d, double* out,
                void* stream);
/* what: 0 = dist(X_i, Y_i) (b = Y);  1 = norm_X(U) = |L^-1 U L^-T|_F (b = U);  2 = inner_X(U, V) (b = U, c = V) */
int gabo_spd_scalar(int what, const double* x, const double* b, const double* c, int64_t n, int d, double* out,
                    void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * A4: acquisition value.  Replaces botorch ExpectedImprovement(model, best_f, maximize=False) over a SingleTaskGP with
 * ScaleKernel(SphereGaussianKernel | SpdAffineInvariantGaussianKernel) (call sites gabo_sphere.py:131-165,
 * gabo_spd.py:165-197).  The GP is described by precomputed device arrays:
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct gabo_gp_desc {
    int32_t manifold;       /* GABO_SPHERE | GABO_SPD                                                        */
    int32_t dim;            /* sphere: ambient D; spd: matrix size d                                          */
    int32_t n_train;        /* <= GABO_MAX_TRAIN                                                              */
    int32_t compute;        /* GABO_F32 | GABO_F64: arithmetic type of the per-restart work                   */
    const double* x_train;  /* sphere: n x D unit vectors; spd: n x gabo_spd_factor_stride(d) (gabo_spd_factor) */
    const double* alpha;    /* n:     (s K + noise I)^-1 (y - mean)                                           */
    const double* minv;     /* n x n: (s K + noise I)^-1                                                      */
    double mean;            /* constant mean                                                                  */
    double outputscale;     /* s                                                                              */
    double beta;            /* kernel parameter                                                               */
    double best_f;          /* incumbent (minimisation)                                                       */
    double kxx;             /* base-kernel k(x,x) (1 up to the reference's 1e-15 guards)                       */
} gabo_gp_desc;

/* EI and (optionally) its Riemannian gradient at r points.  x: sphere r x D, spd r x d x d (fp64).
 * ei: r (fp64).  grad (nullable): same shape as x.  Points that are not SPD get ei = NaN. */
int gabo_ei_eval(const gabo_gp_desc* gp, const double* x, int64_t r, double* ei, double* grad, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * A1: batched multi-start Riemannian conjugate gradient on cost = -EI.  Replaces the serial restart loop of
 * gen_candidates_manifold (manifold_optimization/manifold_optimize.py:207-221) with pymanopt ConjugateGradient
 * (Hestenes-Stiefel) + LineSearchAdaptive per restart.  One warp per restart, every step inside one launch.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct gabo_rcg_opts {
    int32_t maxiter;          /* pymanopt Solver maxiter (default 1000)            */
    int32_t ls_maxiter;       /* LineSearchAdaptive maxiter (10)                   */
    double mingradnorm;       /* 1e-6                                              */
    double minstepsize;       /* 1e-10                                             */
    double contraction;       /* 0.5                                               */
    double suff_decr;         /* 0.5                                               */
    double initial_stepsize;  /* 1.0                                               */
} gabo_rcg_opts;

/* x (in/out): r starting points -> r candidates.  value: r, EI at the candidate (fp64).  iters / reason: r (int32,
 * nullable); reason 1 = maxiter, 2 = gradnorm, 3 = stepsize. */
int gabo_acq_rcg(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rcg_opts* opts, double* value,
                 int32_t* iters, int32_t* reason, void* stream);

/* Trust-region variant: the reference's own TrustRegions solver
 * (manifold_optimization/robust_trust_regions.py:116-352 with _truncated_conjugate_gradient :410-520) over the
 * finite-difference Hessian of manifold_optimization/approximate_hessian.py:11-62, one warp per restart: spheres of
 * ambient dimension <= 8 (or <= 16 with n_train <= 64) and SPD(d), d <= 8 (fp64; see gabo_acq_ctr).
 * reason 1 = maxiter, 2 = gradnorm. */
typedef struct gabo_rtr_opts {
    int32_t maxiter;            /* pymanopt Solver maxiter (1000)                                   */
    int32_t mininner;           /* TrustRegions.solve mininner (1)                                  */
    int32_t maxinner;           /* <= 0: manifold.dim                                               */
    int32_t reserved;
    double mingradnorm;         /* 1e-6                                                             */
    double kappa;               /* 0.1   (robust_trust_regions.py:95-96)                            */
    double theta;               /* 1.0                                                              */
    double rho_prime;           /* 0.1                                                              */
    double rho_regularization;  /* 1e3                                                              */
    double delta_bar;           /* <= 0: manifold.typicaldist (pi on the sphere)                    */
    double delta0;              /* <= 0: delta_bar / 8                                              */
} gabo_rtr_opts;
int gabo_acq_rtr(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_rtr_opts* opts, double* value,
                 int32_t* iters, int32_t* reason, void* stream);

/* Constrained variant on SPD(d) (what examples/bo_spd/benchmark_examples/gabo_spd.py:183,200-203 runs): the reference's
 * ConstrainedTrustRegions (manifold_optimization/constrained_trust_regions.py:120-439, constrained tCG :441-735) and
 * StrictConstrainedTrustRegions (:737-1415; infeasible proposals rejected, :932-952, :1036) with up to two eigenvalue
 * inequality constraints (Riemannian_utils/spd_constraints_utils_torch.py:17-50):
 *   GABO_CONS_MAX_EIG: bound - lambda_max(X) >= 0,   GABO_CONS_MIN_EIG: lambda_min(X) - bound >= 0.
 * One warp per restart, the whole solve inside one launch, fp64 (gp->compute is ignored: the stopping rule of the
 * solver lies below the fp32 noise floor of the SPD gradient).  n_constraints = 0 is plain TrustRegions on SPD(d), which
 * is also what gabo_acq_rtr runs for gp->manifold == GABO_SPD.  x: r x d x d (in/out).  reason 1 = maxiter,
 * 2 = gradnorm, -1 = start not positive definite (value NaN). */
#define GABO_CONS_MAX_EIG 0
#define GABO_CONS_MIN_EIG 1
typedef struct gabo_ctr_opts {
    gabo_rtr_opts tr;          /* delta_bar <= 0: sqrt(d(d+1)/2) (pymanopt PositiveDefinite.typicaldist); maxinner <= 0: d(d+1)/2 */
    int32_t n_constraints;     /* 0, 1 or 2                                                       */
    int32_t strict;            /* != 0: StrictConstrainedTrustRegions                             */
    int32_t kind[2];           /* GABO_CONS_*                                                     */
    double bound[2];
    double delta_cons;         /* Delta_cons of ConstrainedTrustRegions.solve (1e-6)              */
} gabo_ctr_opts;
int gabo_acq_ctr(const gabo_gp_desc* gp, double* x, int64_t r, const gabo_ctr_opts* opts, double* value,
                 int32_t* iters, int32_t* reason, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * A3: candidate selection.  Replaces botorch get_best_candidates (argmax of batch values, manifold_optimize.py:118-120).
 * Lexicographic (value descending, global index ascending), NaN = -inf: identical on 1 or N GPUs.
 * values: n fp64; gidx: n int64 global restart indices (nullable = 0..n-1); out_slot: one int64 = winning position in
 * the arrays; out_value: one fp64.
 * ------------------------------------------------------------------------------------------------------------------ */
int gabo_argmax_records(const double* values, const int64_t* gidx, int64_t n, int64_t* out_slot, double* out_value,
                        void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * P1: nested SPD projection Y = W^T X W (nested_mappings/nested_spd_utils.py:13-48) in Mandel coordinates:
 *   y_mandel[n x dvl] = x_mandel[n x dvh] * P^T,  dvh = D(D+1)/2, dvl = d(d+1)/2, P built from W (D x d, fp64).
 * fp32 I/O, 3xTF32 tensor-core contraction with fp32 accumulation (error ~1e-6 relative).
 * gabo_nested_projection_matrix writes P, split into tf32 hi/lo parts and arranged per MMA lane, into `p_pack`
 * (gabo_nested_projection_pack_size(D, d) floats, 16-byte aligned, caller-owned); gabo_nested_spd_project consumes it.
 * ------------------------------------------------------------------------------------------------------------------ */
/* Reference-precision (fp64) form for the nested kernels (kernel_utils/kernels_nested_spd.py:122-127):
 * y_mandel[n x d(d+1)/2] = Mandel(W^T X W), x_mandel: n x D(D+1)/2, w: D x d row-major, all fp64. */
int gabo_nested_spd_project_f64(const double* x_mandel, int64_t n, int D, int d, const double* w, double* y_mandel,
                                void* stream);
int64_t gabo_nested_projection_pack_size(int D, int d);
int gabo_nested_projection_matrix(const double* w, int D, int d, float* p_pack, void* stream);
int gabo_nested_spd_project(const float* x_mandel, int64_t n, int D, int d, const float* p_pack, float* y_mandel,
                            void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Approximate inverse of P1, projection_from_nested_spd_to_spd (nested_mappings/nested_spd_utils.py:51-118):
 *   X = R [Y B; B^T C] R^T,  R = [W V],  B = Y^(1/2) K C^(1/2);   fp64, d <= GABO_MAX_SPD_DIM, d < D <= 32.
 * gabo_spd_sqrtm is sqrtm_torch (Riemannian_utils/spd_utils_torch.py:33-50) for a batch of d x d matrices (NaN output
 * for a matrix that is not positive definite).  gabo_nested_spd_reconstruct_setup folds the point-independent factors
 * (W: D x d, V: D x (D-d), C: (D-d)^2, K: d x (D-d), row-major) into `pack`
 * (gabo_nested_spd_reconstruct_pack_size(D, d) doubles, caller-owned); *flag (device int, zeroed by the caller) is set
 * when C is not positive definite.  gabo_nested_spd_reconstruct maps y (n x d x d) with y_sqrt = sqrtm(y) to
 * x (n x D x D).
 * ------------------------------------------------------------------------------------------------------------------ */
int gabo_spd_sqrtm(const double* mat, int64_t n, int d, double* out, void* stream);
int64_t gabo_nested_spd_reconstruct_pack_size(int D, int d);
int gabo_nested_spd_reconstruct_setup(const double* w, const double* v, const double* c, const double* k, int D, int d,
                                      double* pack, int* flag, void* stream);
int gabo_nested_spd_reconstruct(const double* y, const double* y_sqrt, int64_t n, int D, int d, const double* pack,
                                double* x, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * GP hyper-parameter fit: the objective botorch.fit_gpytorch_model evaluates at every L-BFGS step
 * (examples/bo_sphere/benchmark_examples/gabo_sphere.py:131-165: SingleTaskGP = constant mean +
 * ScaleKernel(geodesic kernel) + GaussianLikelihood under gpytorch.mlls.ExactMarginalLogLikelihood), for `batch`
 * hyper-parameter sets at once, one CTA each:
 *   K = s exp(-beta dmat) + noise I,  ll = log N(y | m 1, K),  theta[b] = {beta, s, noise, m}.
 * dmat: n x n (squared geodesic distances for the Gaussian kernels, distances for the Laplace kernels), y: n.
 * out_ll: batch (NOT divided by n, priors not included); out_grad: batch x 4 = d ll / d{beta, s, noise, m} (nullable);
 * out_alpha: batch x n = K^-1 (y - m) and out_kinv: batch x n x n = K^-1 (nullable; what gabo_gp_desc consumes);
 * flags: batch ints, 1 where K is not positive definite (outputs NaN).  fp64, n <= 128.
 * ------------------------------------------------------------------------------------------------------------------ */
int gabo_gp_mll(const double* dmat, int64_t n, const double* y, const double* theta, int64_t batch, double* out_ll,
                double* out_grad, double* out_alpha, double* out_kinv, int* flags, void* stream);
/* The whole fit in ONE launch (what botorch.fit_gpytorch_model does with scipy's L-BFGS-B, one objective evaluation =
 * one Gram rebuild + Cholesky + autograd per call, gabo_sphere.py:162): `batch` independent starts, one CTA each, run BFGS
 * (dense 4 x 4 inverse Hessian = full-memory L-BFGS, Armijo backtracking) on
 *   f(raw) = -(ll(theta) + log-priors(theta)) / n,
 *   theta = (beta_min + softplus(raw[0]), softplus(raw[1]), noise_min + softplus(raw[2]), raw[3])
 * i.e. gpytorch's constraints and botorch's objective.  raw0 / out_raw: batch x 4; priors: 6 doubles = Gamma
 * (concentration, rate) for beta, outputscale, noise (concentration <= 0: no prior); fixed: 4 ints, != 0 holds that raw
 * parameter at its start value; stop: max|g| <= pgtol, relative decrease <= ftol, or maxiter.
 * out_f: batch (inf when the start is not positive definite); out_info: batch x 3 ints = {status (0 pgtol, 1 ftol,
 * 2 maxiter, 3 line search failed, 4 not PD), iterations, objective evaluations}. */
int gabo_gp_fit(const double* dmat, int64_t n, const double* y, const double* raw0, int64_t batch, double beta_min,
                double noise_min, const double* priors, const int* fixed, int maxiter, double pgtol, double ftol,
                double* out_raw, double* out_f, int* out_info, void* stream);
/* Factorisation only, from a base-kernel matrix the Gram kernels already produced (what gabo_gp_desc needs once the
 * hyper-parameters are known): alpha = (s kmat + noise I)^-1 (y - m), kinv = (s kmat + noise I)^-1; *flag = 1 when the
 * matrix is not positive definite.  Same kernel, same limits. */
int gabo_gp_factor(const double* kmat, int64_t n, const double* y, double outputscale, double noise, double mean,
                   double* out_alpha, double* out_kinv, int* flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GABO_B200_H */
