"""Multi-start Riemannian acquisition optimisation with the reference's function names and signatures
(``BoManifolds/manifold_optimization/manifold_optimize.py``), batched on the B200.

Reference flow (manifold_optimize.py:36-321): raw manifold samples -> acquisition values -> botorch's
``initialize_q_batch[_nonneg]`` picks ``num_restarts`` starts -> ONE pymanopt solve PER restart in a Python loop
(:207-221), each cost / gradient call going through torch autograd -> acquisition re-evaluated at the candidates (:227)
-> ``get_best_candidates`` = argmax (:118-120).

Here the restarts are solved together by ``gabo_acq_rcg`` (one warp per restart, every CG step inside one launch, the
EI gradient in closed form), the raw samples are scored by ``gabo_ei_eval`` and the winner is picked by
``gabo_argmax_records``.  With ``torch.distributed`` initialised and ``options={'distributed': True}`` the restarts are
sharded across ranks and ONE all-gather of (value, global index, candidate) records ends the solve.

The fast path needs a recognised triple (no CPU fallback):
  manifold    : ``Sphere`` / ``PositiveDefinite`` (ours, or pymanopt's by class name),
  solver      : conjugate gradient (``ConjugateGradient`` below, or pymanopt's by class name),
  acquisition : ``ExpectedImprovement`` over a GP whose kernel is ``ScaleKernel(SphereGaussianKernel |
                SpdAffineInvariantGaussianKernel)`` (our ``ManifoldGP`` or a botorch ``SingleTaskGP`` duck-typed).
Anything else raises ``NotImplementedError``.
"""
import math
import warnings

import torch

from . import _lib, ops
from .kernel_utils import (SphereGaussianKernel, SpdAffineInvariantGaussianKernel, SpdLogEuclideanGaussianKernel,
                           SpdFrobeniusGaussianKernel)


class BadInitialCandidatesWarning(RuntimeWarning):
    pass


# ----------------------------------------------------------------------------------------------------------------
# solver description
# ----------------------------------------------------------------------------------------------------------------

class ConjugateGradient:
    """Options of the batched Riemannian CG (pymanopt ``ConjugateGradient`` constructor surface: Hestenes-Stiefel,
    adaptive backtracking line search; defaults are pymanopt's ``Solver`` / ``LineSearchAdaptive`` defaults)."""

    def __init__(self, beta_type='HestenesStiefel', orth_value=float('inf'), linesearch=None, maxtime=1000,
                 maxiter=1000, mingradnorm=1e-6, minstepsize=1e-10, maxcostevals=5000, logverbosity=0,
                 contraction_factor=0.5, suff_decr=0.5, ls_maxiter=10, initial_stepsize=1.0):
        if beta_type not in ('HestenesStiefel', 2):
            raise NotImplementedError('the batched solver implements the Hestenes-Stiefel rule (pymanopt default)')
        self._maxtime = maxtime
        self._maxiter = maxiter
        self._mingradnorm = mingradnorm
        self._minstepsize = minstepsize
        self._maxcostevals = maxcostevals
        self._logverbosity = logverbosity
        self.contraction_factor = contraction_factor
        self.suff_decr = suff_decr
        self.ls_maxiter = ls_maxiter
        self.initial_stepsize = initial_stepsize


class TrustRegions:
    """Options of the batched Riemannian trust-region solver: constructor surface of the reference's own
    ``TrustRegions`` (manifold_optimization/robust_trust_regions.py:91-108; pymanopt ``Solver`` keywords behind it).
    Hessian-vector products are the finite differences of ``approximate_hessian.py`` -- what the reference uses with
    ``approx_hessian=True``; with ``approx_hessian=False`` the reference differentiates the gradient with autograd
    instead, which this path does not do (the two differ by O(2^-14) relative)."""

    def __init__(self, miniter=3, kappa=0.1, theta=1.0, rho_prime=0.1, use_rand=False, rho_regularization=1e3,
                 maxtime=1000, maxiter=1000, mingradnorm=1e-6, minstepsize=1e-10, maxcostevals=5000, logverbosity=0,
                 mininner=1, maxinner=None, Delta_bar=None, Delta0=None):
        if use_rand:
            raise NotImplementedError('use_rand=True (randomised tCG start + Cauchy point) is not supported')
        self.miniter, self.kappa, self.theta, self.rho_prime = miniter, kappa, theta, rho_prime
        self.use_rand, self.rho_regularization = use_rand, rho_regularization
        self._maxtime, self._maxiter, self._mingradnorm = maxtime, maxiter, mingradnorm
        self._minstepsize, self._maxcostevals, self._logverbosity = minstepsize, maxcostevals, logverbosity
        # arguments of TrustRegions.solve in the reference (:110-111); the batched entry takes them here
        self.mininner, self.maxinner, self.Delta_bar, self.Delta0 = mininner, maxinner, Delta_bar, Delta0


def _trust_region_options(solver):
    if getattr(solver, 'use_rand', False):
        raise NotImplementedError('use_rand=True (randomised tCG start + Cauchy point) is not supported')
    return dict(
        maxiter=int(getattr(solver, '_maxiter', 1000)),
        mingradnorm=float(getattr(solver, '_mingradnorm', 1e-6)),
        kappa=float(getattr(solver, 'kappa', 0.1)),
        theta=float(getattr(solver, 'theta', 1.0)),
        rho_prime=float(getattr(solver, 'rho_prime', 0.1)),
        rho_regularization=float(getattr(solver, 'rho_regularization', 1e3)),
        mininner=int(getattr(solver, 'mininner', 1)),
        maxinner=getattr(solver, 'maxinner', None),
        delta_bar=getattr(solver, 'Delta_bar', None),
        delta0=getattr(solver, 'Delta0', None),
    )


class ConstrainedTrustRegions(TrustRegions):
    """Constructor surface of the reference's ``ConstrainedTrustRegions`` (constrained_trust_regions.py:102-118;
    ``Delta_cons`` is an argument of its ``solve``, :120-121)."""

    def __init__(self, *args, Delta_cons=1e-6, **kwargs):
        super().__init__(*args, **kwargs)
        self.Delta_cons = Delta_cons


class StrictConstrainedTrustRegions(ConstrainedTrustRegions):
    """The reference's ``StrictConstrainedTrustRegions`` (constrained_trust_regions.py:737-1415): proposals that violate
    a constraint are rejected outright (``hd_gabo_spd.py``'s acquisition solver)."""


class AugmentedLagrangeMethod:
    """Constructor surface of the reference's ``AugmentedLagrangeMethod`` (augmented_Lagrange_method.py:30-62; pymanopt
    ``Solver`` keywords behind it).  The batched acquisition path (``gen_candidates_manifold``) drives ``TrustRegions`` as
    inner solver; the host path of the reconstruction fit (``manifold_gp_fit.riemannian_alm``) also takes
    ``ConjugateGradient`` (the inner solver of hd_gabo_spd.py:192)."""

    def __init__(self, inner_solver, bound=20, rho_init=1, thetarho=0.3, tau=0.8, starting_tolgradnorm=1e-3,
                 ending_tolgradnorm=1e-6, lambdas_fact=1., gammas_fact=1., maxtime=1000, maxiter=1000, mingradnorm=1e-6,
                 minstepsize=1e-10, maxcostevals=5000, logverbosity=0):
        if type(inner_solver).__name__ not in ('TrustRegions', 'ConjugateGradient'):
            raise NotImplementedError('AugmentedLagrangeMethod: inner_solver must be TrustRegions or ConjugateGradient')
        self.inner_solver = inner_solver
        self._bound, self._rho_init, self._thetarho, self._tau = bound, rho_init, thetarho, tau
        self._starting_tolgradnorm, self._ending_tolgradnorm = starting_tolgradnorm, ending_tolgradnorm
        self._lambdas_fact, self._gammas_fact = lambdas_fact, gammas_fact
        self._maxtime, self._maxiter, self._mingradnorm = maxtime, maxiter, mingradnorm
        self._minstepsize, self._maxcostevals, self._logverbosity = minstepsize, maxcostevals, logverbosity


def batched_constraints(constraints, manifold_kind):
    """Turn the reference's list of inequality constraints (callables of ONE point returning a zero-dim tensor,
    positive when satisfied; examples/bo_spd/benchmark_examples/gabo_spd.py:136-138) into one callable of the whole
    batch: ``X (R, ...) -> (values (R, C), [C Riemannian gradients shaped like X])``.

    ``functools.partial(max_eigenvalue_constraint_torch | min_eigenvalue_constraint_torch, ...)`` from
    ``gabotorch_b200.riemannian_utils`` is evaluated in closed form for the batch (extreme eigenpair: value
    ``+-(bound - lambda)``, Euclidean gradient ``-+ v v^T``; the eigenpairs come from one ``gabo_sym_eig`` launch).  A
    callable that declares ``supports_batch = True`` (the nested-SPD eigenvalue constraints of ``nested_mappings``) is
    differentiated with torch.autograd over the whole batch at once; any other callable one restart at a time, which is
    what the reference's ``Problem`` does (pymanopt_addons/problem.py:118-137)."""
    from . import riemannian_utils as ru
    constraints = list(constraints)

    def rgrad(X, E):
        if manifold_kind == _lib.SPD:
            return ops.spd_op(_lib.OP_EGRAD2RGRAD, X, E)                   # X sym(E) X
        return E - (X * E).sum(-1, keepdim=True) * X                      # projection onto the tangent space

    def closed_form(c):
        f = getattr(c, 'func', None)
        kw = getattr(c, 'keywords', None) or {}
        args = getattr(c, 'args', ())
        if f is ru.max_eigenvalue_constraint_torch:
            return 'max', float(kw.get('maximum_eigenvalue', args[0] if args else float('nan')))
        if f is ru.min_eigenvalue_constraint_torch:
            return 'min', float(kw.get('minimum_eigenvalue', args[0] if args else float('nan')))
        return None

    kinds = [closed_form(c) for c in constraints]

    def evaluate(X):
        vals, grads = [], []
        eig = None
        for c, k in zip(constraints, kinds):
            if k is not None:
                if eig is None:
                    eig = ops.sym_eig(X)[:2]                               # unsorted eigenpairs, one launch
                lam, vec = eig
                i = lam.argmax(-1) if k[0] == 'max' else lam.argmin(-1)
                ext = lam.gather(-1, i.unsqueeze(-1)).squeeze(-1)
                v = vec.gather(-1, i[..., None, None].expand(vec.shape[:-1] + (1,))).squeeze(-1)
                outer = v.unsqueeze(-1) * v.unsqueeze(-2)
                vals.append(k[1] - ext if k[0] == 'max' else ext - k[1])
                grads.append(rgrad(X, -outer if k[0] == 'max' else outer))
            elif getattr(getattr(c, 'func', c), 'supports_batch', False):
                xs = X.detach().clone().requires_grad_(True)
                with torch.enable_grad():
                    f = c(xs)
                    e, = torch.autograd.grad(f.sum(), xs)
                vals.append(f.detach())
                grads.append(rgrad(X, e))
            else:
                xs = X.detach().clone().requires_grad_(True)
                with torch.enable_grad():
                    f = torch.stack([c(xs[i]) for i in range(xs.shape[0])])
                    e, = torch.autograd.grad(f.sum(), xs)
                vals.append(f.detach())
                grads.append(rgrad(X, e))
        return torch.stack(vals, dim=-1), grads
    return evaluate


_CONSTRAINED_SOLVERS = ('ConstrainedTrustRegions', 'StrictConstrainedTrustRegions')


def _rtr_kernel_covers(gp):
    """The one-launch trust-region kernels (gabo_acq_rtr): spheres of ambient dimension <= 8 (or <= 16 with at most 64
    training points, register-resident iterate) and SPD(d) (one warp per restart, whitened coordinates)."""
    if getattr(gp, 'is_tensor_gp', False):
        return False
    if gp.manifold == _lib.SPD:
        return True
    return gp.dim <= 8 or (gp.dim <= 16 and gp.n_train <= 64)


def eigenvalue_constraint_specs(constraints):
    """``[('max' | 'min', bound), ...]`` when EVERY constraint is a ``functools.partial`` of
    ``max_eigenvalue_constraint_torch`` / ``min_eigenvalue_constraint_torch`` (the constraints of gabo_spd.py:136-138) and
    there are at most two of them -- the set ``gabo_acq_ctr`` evaluates inside the solver kernel; None otherwise."""
    from . import riemannian_utils as ru
    specs = []
    for c in constraints:
        f = getattr(c, 'func', None)
        kw = getattr(c, 'keywords', None) or {}
        args = getattr(c, 'args', ())
        if f is ru.max_eigenvalue_constraint_torch and ('maximum_eigenvalue' in kw or args):
            specs.append(('max', float(kw.get('maximum_eigenvalue', args[0] if args else 0.0))))
        elif f is ru.min_eigenvalue_constraint_torch and ('minimum_eigenvalue' in kw or args):
            specs.append(('min', float(kw.get('minimum_eigenvalue', args[0] if args else 0.0))))
        else:
            return None
    return specs if 0 < len(specs) <= 2 else None


def batched_trust_regions(gp, x0, maxiter=1000, mingradnorm=1e-6, kappa=0.1, theta=1.0, rho_prime=0.1,
                          rho_regularization=1e3, mininner=1, maxinner=None, delta_bar=None, delta0=None,
                          ineq_constraints=None, delta_cons=1e-6, strict=False, eq_constraints=None, penalty=None):
    """The reference's ``TrustRegions.solve`` (robust_trust_regions.py:116-352, tCG :410-520, finite-difference Hessian
    approximate_hessian.py:11-62) for ALL restarts in lock-step, for the cases the single-launch kernel does not cover:
    SPD(d) and spheres of large ambient dimension.  Every cost / gradient evaluation is one launch of ``gabo_ei_eval``
    over the whole batch, every retraction / inner product one launch of ``gabo_spd_op`` / ``gabo_spd_scalar``
    (pymanopt ``PositiveDefinite``: exp-map retraction, identity transport, affine-invariant inner product); the
    per-restart scalars of the recurrences live in (R,) device tensors and finished restarts are masked out, so each
    restart follows exactly the serial algorithm.  Returns (candidates, values, iters, reasons) like ``ops.acq_rtr``.
    Launch-bound by construction (about 8 launches per inner iteration); fusing it into the CTA-per-restart SPD kernel
    is the next step.

    ``ineq_constraints`` (a callable ``X -> (values (R, C), [C Riemannian gradients shaped like X])``, see
    ``batched_constraints``) switches to the reference's ``ConstrainedTrustRegions``
    (constrained_trust_regions.py:120-439, constrained tCG :441-735): the linearised constraints
    ``c(x) + <grad c, eta>`` are kept within ``delta_cons`` of feasibility along the tCG path (only negative inequality
    terms count), the step is cut by the root of the corresponding quadratic, and the radius also grows after a step
    that stopped on the constraints.  ``strict=True`` is ``StrictConstrainedTrustRegions`` (:737-1415): a proposal that
    violates a constraint gets an infinite cost, is rejected and shrinks the radius (:936-952, :972).
    ``eq_constraints`` (same callable form, satisfied when zero) are held in the same way with every term counting
    (``gabo_sphere_equality_constraints.py``).  Equality and inequality constraints together are refused: the reference
    builds its active set for that case with ``np.where(term < 0)[0] + n_eq`` over the whole vector (:573-577), which
    indexes past the inequality block.
    ``penalty`` (``X -> (values (R,), Riemannian gradients)``) is added to the cost and to its gradient but NOT to the
    gradient the finite-difference Hessian differentiates: the augmented Lagrangian subproblem of ``batched_alm``
    (augmented_Lagrange_method.py:322 binds the Hessian to the original problem)."""
    X = ops.to_dev64(x0).clone()
    R = X.shape[0]
    dev = X.device
    lead = (R,) + (1,) * (X.dim() - 1)

    def bc(v):
        return v.reshape(lead)

    if gp.manifold == _lib.SPD:
        d = X.shape[-1]
        dim, typical = d * (d + 1) // 2, math.sqrt(d * (d + 1) // 2)      # pymanopt PositiveDefinite.dim / typicaldist

        def inner(P, U, V):
            return ops.spd_scalar(2, P, U, V)

        def norm(P, U):
            return ops.spd_scalar(1, P, U)

        def retr(P, U):
            return ops.spd_op(_lib.OP_RETR, P, U)

        def transp(P1, P, G):                                             # pymanopt 0.2.x: identity
            return G
    else:
        dim, typical = X.shape[-1] - 1, math.pi                           # pymanopt Sphere.dim / typicaldist

        def inner(P, U, V):
            return (U * V).sum(-1)

        def norm(P, U):
            return (U * U).sum(-1).sqrt()

        def retr(P, U):
            Y = P + U
            return Y / Y.norm(dim=-1, keepdim=True)

        def transp(P1, P, G):                                             # projection onto the tangent space at P
            return G - (P * G).sum(-1, keepdim=True) * P

    inf = torch.full((R,), float('inf'), dtype=torch.float64, device=dev)

    def plain_cost_grad(P):
        ei, g = ops.ei_eval(gp, P, want_grad=True)
        c = -ei
        return torch.where(torch.isfinite(c), c, inf), -g

    def cost(P):
        c = -ops.ei_eval(gp, P)
        if penalty is not None:
            c = c + penalty(P)[0]
        return torch.where(torch.isfinite(c), c, inf)

    def cost_grad(P):
        c, g = plain_cost_grad(P)
        if penalty is not None:
            pc, pg = penalty(P)
            c, g = c + pc, g + pg
        return torch.where(torch.isfinite(c), c, inf), g

    maxinner = int(dim if maxinner is None else maxinner)
    delta_bar = float(typical if delta_bar is None else delta_bar)
    delta0 = float(delta_bar / 8 if delta0 is None else delta0)
    fd_eps = 2.0 ** -14                                                   # approximate_hessian.py:43
    eps = 2.220446049250313e-16                                           # np.spacing(1)
    NEG, EXC, LIN, SUP, MAXI, INC, CONS = range(7)
    dc2 = float(delta_cons) ** 2

    if eq_constraints is not None and ineq_constraints is not None:
        raise NotImplementedError('equality and inequality constraints together are not supported (see the docstring)')
    constraints = eq_constraints if eq_constraints is not None else ineq_constraints
    is_eq = eq_constraints is not None

    def step_to_constraints(fc, pe, pd, step):
        """(violated (R,), tau (R,)): does the linearised constraint term leave the delta_cons ball at ``step``, and the
        step that brings it back onto it (constrained_trust_regions.py:565-592 / :622-650; inequality constraints only,
        where the reference's index set is exactly the negative terms)."""
        term = fc + pe + step.unsqueeze(-1) * pd
        if not is_eq:
            term = torch.clamp(term, max=0.0)
        bad = (term * term).sum(-1) > dc2
        m = torch.ones_like(term) if is_eq else (term < 0).to(fc.dtype)
        qa = (m * pd * pd).sum(-1)
        qb = 2.0 * ((m * fc * pd).sum(-1) + (m * pe * pd).sum(-1))
        qc = (m * fc * fc).sum(-1) + 2.0 * (m * fc * pe).sum(-1) + (m * pe * pe).sum(-1) - dc2
        disc = qb * qb - 4.0 * qa * qc
        tau = torch.where(disc >= 0, (-qb + disc.clamp(min=0).sqrt()) / (2.0 * qa), torch.zeros_like(disc))
        return bad, tau

    def hess(P, G, A):
        na = norm(P, A)
        small = na < 1e-15
        c = bc(fd_eps / torch.where(small, torch.ones_like(na), na))
        P1 = retr(P, c * A)
        _, G1 = plain_cost_grad(P1)
        H = transp(P1, P, G1) / c - G / c
        return torch.where(bc(small), torch.zeros_like(H), H)

    fx, G = cost_grad(X)
    Gh = G if penalty is None else plain_cost_grad(X)[1]      # what the finite-difference Hessian differentiates
    ng = norm(X, G)
    radius = torch.full((R,), delta0, dtype=torch.float64, device=dev)
    k = torch.zeros(R, dtype=torch.int32, device=dev)
    reason = torch.zeros(R, dtype=torch.int32, device=dev)
    active = torch.ones(R, dtype=torch.bool, device=dev)
    zero = torch.zeros(R, dtype=torch.float64, device=dev)
    while bool(active.any()):
        # ---- truncated CG (use_rand=False, identity preconditioner) ----
        eta, heta, r = torch.zeros_like(X), torch.zeros_like(X), G.clone()
        e_pe, e_pd, model_value = zero.clone(), zero.clone(), zero.clone()
        r_r = inner(X, r, r)
        norm_r0 = r_r.sqrt()
        z_r, d_pd, delta = r_r.clone(), r_r.clone(), -r
        stop = torch.full((R,), MAXI, dtype=torch.int32, device=dev)
        live = active.clone()
        r2 = radius * radius
        pw = norm_r0 ** theta
        if constraints is not None:
            fc, gcs = constraints(X)                            # (R, C) values, C Riemannian gradients
            pe = torch.zeros_like(fc)                           # <grad c, eta>, eta = 0
        for j in range(maxinner):
            if not bool(live.any()):
                break
            hdelta = hess(X, Gh, delta)
            d_hd = inner(X, delta, hdelta)
            nz = d_hd != 0
            alpha = torch.where(nz, z_r / torch.where(nz, d_hd, torch.ones_like(d_hd)), zero)
            e_pe_new = torch.where(nz, e_pe + 2 * alpha * e_pd + alpha * alpha * d_pd, e_pe)
            out1 = live & ((d_hd <= 0) | (e_pe_new >= r2))
            tau = (-e_pd + (e_pd * e_pd + d_pd * (r2 - e_pe)).sqrt()) / d_pd
            code1 = torch.where(d_hd <= 0, NEG, EXC).to(torch.int32)
            if constraints is not None:
                pd = torch.stack([inner(X, gc, delta) for gc in gcs], dim=-1)
                tau = torch.where(torch.isnan(tau), torch.zeros_like(tau), tau)          # :562-563
                bad, tau_c = step_to_constraints(fc, pe, pd, tau)
                tau = torch.where(bad, tau_c, tau)
                code1 = torch.where((d_hd > 0) & bad, torch.full_like(code1, CONS), code1)
            eta = torch.where(bc(out1), eta + bc(tau) * delta, eta)
            heta = torch.where(bc(out1), heta + bc(tau) * hdelta, heta)
            stop = torch.where(out1, code1, stop)
            live = live & ~out1
            if constraints is not None:                         # the full CG step would violate the constraints
                bad, tau_c = step_to_constraints(fc, pe, pd, alpha)
                outc = live & bad
                eta = torch.where(bc(outc), eta + bc(tau_c) * delta, eta)
                heta = torch.where(bc(outc), heta + bc(tau_c) * hdelta, heta)
                stop = torch.where(outc, torch.full_like(stop, CONS), stop)
                live = live & ~outc
            new_eta = eta + bc(alpha) * delta
            new_heta = heta + bc(alpha) * hdelta
            new_mv = inner(X, new_eta, G) + 0.5 * inner(X, new_eta, new_heta)
            out2 = live & (new_mv >= model_value)
            stop = torch.where(out2, torch.full_like(stop, INC), stop)
            live = live & ~out2
            eta = torch.where(bc(live), new_eta, eta)
            heta = torch.where(bc(live), new_heta, heta)
            model_value = torch.where(live, new_mv, model_value)
            e_pe = torch.where(live, e_pe_new, e_pe)
            r = torch.where(bc(live), r + bc(alpha) * hdelta, r)
            r_r = inner(X, r, r)
            out3 = live & (r_r.sqrt() <= norm_r0 * torch.minimum(pw, torch.full_like(pw, kappa)))
            if j < mininner:
                out3 = out3 & False
            stop = torch.where(out3, torch.where(kappa < pw, LIN, SUP).to(torch.int32), stop)
            live = live & ~out3
            beta = r_r / z_r
            delta = torch.where(bc(live), -r + bc(beta) * delta, delta)
            e_pd = torch.where(live, beta * (e_pd + alpha * d_pd), e_pd)
            d_pd = torch.where(live, r_r + beta * beta * d_pd, d_pd)
            z_r = torch.where(live, r_r, z_r)
            if constraints is not None:
                pe = torch.where(live.unsqueeze(-1), pe + alpha.unsqueeze(-1) * pd, pe)
        # ---- proposal, rho, radius update, acceptance (robust_trust_regions.py:225-311) ----
        x_prop = retr(X, eta)
        fx_prop = cost(x_prop)
        invalid = torch.zeros_like(active)
        if strict and constraints is not None:
            fcp = constraints(x_prop)[0]
            invalid = (fcp if is_eq else torch.clamp(fcp, max=0.0)).abs().sum(-1) != 0
            fx_prop = torch.where(invalid, inf, fx_prop)
        rho_reg = torch.clamp(fx.abs(), min=1.0) * eps * rho_regularization
        rhonum = fx - fx_prop + rho_reg
        rhoden = -inner(X, G, eta) - 0.5 * inner(X, eta, heta) + rho_reg
        model_decreased = rhoden >= 0
        rho = rhonum / rhoden
        shrink = (rho < 0.25) | ~model_decreased | torch.isnan(rho) | invalid
        grow = ~shrink & (rho > 0.75) & ((stop == NEG) | (stop == EXC) | (stop == CONS))
        radius = torch.where(active & shrink, radius / 4,
                             torch.where(active & grow, torch.clamp(2 * radius, max=delta_bar), radius))
        accept = active & model_decreased & (rho > rho_prime)
        X = torch.where(bc(accept), x_prop, X)
        fx = torch.where(accept, fx_prop, fx)
        _, g_new = cost_grad(X)
        G = torch.where(bc(accept), g_new, G)
        if penalty is not None:
            Gh = torch.where(bc(accept), plain_cost_grad(X)[1], Gh)
        else:
            Gh = G
        ng = torch.where(accept, norm(X, G), ng)
        k = k + active.to(torch.int32)
        hit_iter = active & (k >= maxiter)
        hit_grad = active & ~hit_iter & (ng < mingradnorm)
        reason = torch.where(hit_iter, torch.ones_like(reason), torch.where(hit_grad, 2 * torch.ones_like(reason), reason))
        active = active & ~(hit_iter | hit_grad)
    return X, ops.ei_eval(gp, X), k, reason


def batched_alm(gp, x0, inner_opts, ineq_constraints=None, eq_constraints=None, maxiter=1000, minstepsize=1e-10, bound=20.0,
                rho_init=1.0, thetarho=0.3, tau=0.8, starting_tolgradnorm=1e-3, ending_tolgradnorm=1e-6,
                lambdas_fact=1.0, gammas_fact=1.0):
    """The reference's ``AugmentedLagrangeMethod`` (augmented_Lagrange_method.py:66-203, subproblem :205-328) around the
    lock-step trust-region driver, all restarts together: multipliers, penalty parameters and the stopping test live in
    per-restart device tensors; the inner tolerance follows the reference's schedule.  ``ineq_constraints`` /
    ``eq_constraints`` are ``batched_constraints`` callables.  Returns (candidates, values, outer iterations, reasons)."""
    X = ops.to_dev64(x0).clone()
    R, dev = X.shape[0], X.device
    lead = (R,) + (1,) * (X.dim() - 1)

    def both(P):
        fi, gi = ineq_constraints(P) if ineq_constraints is not None else (P.new_zeros(R, 0), [])
        fe, ge = eq_constraints(P) if eq_constraints is not None else (P.new_zeros(R, 0), [])
        return fi, gi, fe, ge
    fi0, _, fe0, _ = both(X)
    lambdas = torch.full_like(fi0, lambdas_fact)
    gammas = torch.full_like(fe0, gammas_fact)
    rho = torch.full((R,), float(rho_init), dtype=torch.float64, device=dev)
    oldacc = torch.full((R,), float('inf'), dtype=torch.float64, device=dev)
    tol = float(starting_tolgradnorm)
    theta_tol = (ending_tolgradnorm / starting_tolgradnorm) ** (1.0 / maxiter)
    active = torch.ones(R, dtype=torch.bool, device=dev)
    iters = torch.zeros(R, dtype=torch.int32, device=dev)
    reason = torch.zeros(R, dtype=torch.int32, device=dev)
    if gp.manifold == _lib.SPD:
        def dist(A, B):
            return ops.spd_scalar(0, A, B)
    else:
        def dist(A, B):
            return torch.acos(torch.clamp((A * B).sum(-1), -1.0, 1.0))       # pymanopt Sphere.dist
    k = 0
    while bool(active.any()):
        lam, gam, r = lambdas.clone(), gammas.clone(), rho.clone()

        def penalty(P):
            fi, gi, fe, ge = both(P)
            slack = lam / r.unsqueeze(-1) - fi
            on = slack > 0
            val = (r.unsqueeze(-1) / 2.0 * torch.clamp(slack, min=0.0) ** 2).sum(-1) \
                + (r.unsqueeze(-1) / 2.0 * (gam / r.unsqueeze(-1) + fe) ** 2).sum(-1)
            grad = torch.zeros_like(P)
            for c, g_ in enumerate(gi):
                coef = torch.where(on[:, c], fi[:, c] * r - lam[:, c], torch.zeros_like(r))
                grad = grad + coef.reshape(lead) * g_
            for c, g_ in enumerate(ge):
                grad = grad + (fe[:, c] * r + gam[:, c]).reshape(lead) * g_
            return val, grad
        xbest = batched_trust_regions(gp, X, **dict(inner_opts, mingradnorm=tol), penalty=penalty)[0]
        xbest = torch.where(active.reshape(lead), xbest, X)                 # finished restarts keep their result
        fi, _, fe, _ = both(xbest)
        newacc = torch.zeros(R, dtype=torch.float64, device=dev)
        if fi.shape[1]:
            newacc = torch.maximum(newacc, torch.maximum(-lambdas / rho.unsqueeze(-1), fi).abs().amax(-1))
            new_l = torch.clamp(lambdas + rho.unsqueeze(-1) * fi, min=0.0, max=bound)
            lambdas = torch.where(active.unsqueeze(-1), new_l, lambdas)
        if fe.shape[1]:
            newacc = torch.maximum(newacc, fe.abs().amax(-1))
            new_g = torch.clamp(gammas + rho.unsqueeze(-1) * fe, min=-bound, max=bound)
            gammas = torch.where(active.unsqueeze(-1), new_g, gammas)
        grow = active & ((newacc > tau * oldacc) if k > 0 else torch.ones_like(active))
        rho = torch.where(grow, rho / thetarho, rho)
        oldacc = torch.where(active, newacc, oldacc)
        tol = max(ending_tolgradnorm, tol * theta_tol)
        k += 1
        iters = iters + active.to(torch.int32)
        step = dist(xbest, X)
        hit_iter = active & torch.full_like(active, k >= maxiter)
        hit_step = active & ~hit_iter & (step < minstepsize)
        hit_tol = active & ~hit_iter & ~hit_step & torch.full_like(active, tol <= ending_tolgradnorm)
        reason = torch.where(hit_iter, torch.ones_like(reason),
                             torch.where(hit_step, 3 * torch.ones_like(reason),
                                         torch.where(hit_tol, 2 * torch.ones_like(reason), reason)))
        X = xbest
        active = active & ~(hit_iter | hit_step | hit_tol)
    return X, ops.ei_eval(gp, X), iters, reason


def _solver_options(solver):
    name = type(solver).__name__
    if name != 'ConjugateGradient':
        raise NotImplementedError(
            'solver %s: the B200 path batches ConjugateGradient, TrustRegions and [Strict]ConstrainedTrustRegions (the '
            'ALM solver is SURVEY 8f "next"); there is no CPU fallback' % name)
    ls = getattr(solver, '_linesearch', None) or getattr(solver, 'linesearch', None)
    return dict(
        maxiter=int(getattr(solver, '_maxiter', 1000)),
        mingradnorm=float(getattr(solver, '_mingradnorm', 1e-6)),
        minstepsize=float(getattr(solver, '_minstepsize', 1e-10)),
        contraction=float(getattr(solver, 'contraction_factor', getattr(ls, 'contraction_factor', 0.5))),
        suff_decr=float(getattr(solver, 'suff_decr', getattr(ls, 'suff_decr', 0.5))),
        ls_maxiter=int(getattr(solver, 'ls_maxiter', getattr(ls, 'maxiter', 10))),
        initial_stepsize=float(getattr(solver, 'initial_stepsize', getattr(ls, 'initial_stepsize', 1.0))),
    )


def _manifold_kind(manifold):
    name = type(manifold).__name__
    if name == 'Sphere':
        return _lib.SPHERE
    if name in ('PositiveDefinite', 'SymmetricPositiveDefinite'):
        return _lib.SPD
    raise NotImplementedError('manifold %s is not supported by the B200 acquisition optimiser' % name)


# ----------------------------------------------------------------------------------------------------------------
# GP + acquisition
# ----------------------------------------------------------------------------------------------------------------

class ManifoldGP:
    """Exact GP with constant mean over a geodesic kernel: the slice of botorch ``SingleTaskGP`` the acquisition needs
    (reference call sites gabo_sphere.py:131-165).  Hyper-parameters are given, or fitted in place by ``gp_fit.fit_gpytorch_model``."""

    def __init__(self, train_x, train_y, covar_module, noise=1e-2, mean=None, noise_prior=None, noise_min=1e-8):
        self.train_inputs = (torch.as_tensor(train_x, dtype=torch.float64),)
        self.train_targets = torch.as_tensor(train_y, dtype=torch.float64).reshape(-1)
        self.covar_module = covar_module
        self.noise = float(noise)
        self.mean = float(self.train_targets.mean()) if mean is None else float(mean)
        # used by gp_fit.fit_gpytorch_model only: (concentration, rate) of the Gamma prior on the noise and the lower
        # bound of its constraint (GaussianLikelihood(noise_prior=..., noise_constraint=GreaterThan(1e-8)))
        if noise_prior is not None and not isinstance(noise_prior, tuple):
            noise_prior = (float(noise_prior.concentration), float(noise_prior.rate))
        self.noise_prior = noise_prior
        self.noise_min = float(noise_min)
        self.likelihood = None


def _unwrap_model(model):
    """(x_train, y, base_kernel, outputscale, noise, mean) from ManifoldGP or a botorch / gpytorch exact GP."""
    x = model.train_inputs[0]
    y = model.train_targets
    cov = model.covar_module
    base = getattr(cov, 'base_kernel', None)
    if base is None:
        base, scale = cov, 1.0
    else:
        scale = float(cov.outputscale.detach())
    if isinstance(model, ManifoldGP):
        noise, mean = model.noise, model.mean
    else:  # gpytorch ExactGP duck-typing
        noise = float(model.likelihood.noise.reshape(-1)[0])
        mean = float(model.mean_module.constant.reshape(-1)[0])
    return x.reshape(-1, x.shape[-1]), y.reshape(-1), base, scale, noise, mean


class ExpectedImprovement:
    """Analytic EI, botorch semantics (``ExpectedImprovement(model, best_f, maximize=False)``): evaluated on the
    device by ``gabo_ei_eval``.  ``__call__`` takes ``b x 1 x dvec`` (or ``b x dvec``) points in the GP's input
    representation (unit vectors; Mandel vectors for SPD) and returns ``b`` values."""

    nonnegative = True

    def __init__(self, model, best_f, maximize=True, compute='f32'):
        # botorch's default is maximize=True; every reference call site passes maximize=False explicitly
        # (gabo_sphere.py:165, gabo_spd.py:197, hd_gabo_spd.py:259)
        self.model = model
        self.best_f = float(best_f)
        self.maximize = bool(maximize)
        self.compute = compute
        self._gp = None

    def device_gp(self):
        if self._gp is None:
            self._gp = build_device_gp(self.model, self.best_f, self.maximize, self.compute)
        return self._gp

    def __call__(self, X):
        X = torch.as_tensor(X)
        gp = self.device_gp()
        pts = X.reshape(-1, X.shape[-1])
        if gp.manifold == _lib.SPD:
            pts = ops.mandel_unpack(pts)
        out = ops.ei_eval(gp, pts)
        out = out.reshape(X.shape[:-2]) if X.dim() >= 3 else out
        return out if X.is_cuda else out.to(X.device)


def build_device_gp(model, best_f, maximize=False, compute='f32'):
    """Precompute alpha = (sK + noise I)^-1 (y - m) and (sK + noise I)^-1 on the device (n <= 128, fp64)."""
    x, y, base, scale, noise, mean = _unwrap_model(model)
    comp = _lib.GABO_F64 if compute == 'f64' else _lib.GABO_F32
    beta = float(base.beta.detach()) if hasattr(base, 'raw_beta') else float('nan')
    x = ops.to_dev64(x)
    y = ops.to_dev64(y)
    sign = 1.0
    if maximize:
        # EI for maximisation of f == EI for minimisation of -f: flip targets, mean and incumbent
        sign = -1.0
    if isinstance(base, SphereGaussianKernel):
        manifold, dim = _lib.SPHERE, x.shape[-1]
        k = ops.sphere_gram(x, x, beta, _lib.KIND_GAUSS)
        x_dev = x
        kxx = math.exp(-beta * math.acos(1.0 - 1e-15) ** 2)
    elif isinstance(base, SpdAffineInvariantGaussianKernel):
        manifold, dim = _lib.SPD, ops.mandel_dim(x.shape[-1])
        x_dev = ops.spd_factor(x, dim, True)
        k = ops.spd_ai_gram_from_factors(x_dev, x_dev, dim, beta, _lib.KIND_GAUSS, compute=_lib.GABO_F64,
                                         symmetric=True)
        kxx = 1.0  # diagonal_distance=True returns zero distances (spd_utils_torch.py:72-75)
    elif isinstance(base, (SpdLogEuclideanGaussianKernel, SpdFrobeniusGaussianKernel)):
        # kernels_spd.py:190-313 (the latent model of hd_gabo_spd.py): EI as differentiable device tensor code
        dim = ops.mandel_dim(x.shape[-1])
        use_log = isinstance(base, SpdLogEuclideanGaussianKernel)
        mats = ops.mandel_unpack(x)
        s_train = ops.spd_logm(mats) if use_log else mats
        ls = float(base.lengthscale.detach().reshape(-1)[0])
        inv = 1.0 / (ls * ls)
        k = ops.frobenius_gram(s_train, s_train, inv, _lib.KIND_GAUSS)
        alpha, minv = ops.gp_factor(k, sign * y, scale, noise, sign * mean)
        return ops.TensorGP(dim, s_train, alpha, minv, sign * mean, scale, inv, sign * best_f, use_log)
    else:
        raise NotImplementedError('acquisition kernels support SphereGaussianKernel, SpdAffineInvariantGaussianKernel '
                                  'and (through the lock-step solvers) SpdLogEuclidean / SpdFrobeniusGaussianKernel, got %s'
                                  % type(base).__name__)
    # Cholesky factorisation, alpha and the inverse in one launch of the GP kernel (no cuSOLVER on the path)
    alpha, minv = ops.gp_factor(k, sign * y, scale, noise, sign * mean)
    return ops.DeviceGP(manifold, dim, x_dev, alpha, minv, sign * mean, scale, beta, sign * best_f, kxx, comp)


# ----------------------------------------------------------------------------------------------------------------
# initial conditions (botorch initialize_q_batch_nonneg semantics)
# ----------------------------------------------------------------------------------------------------------------

def initialize_q_batch_nonneg(X, Y, n, eta=1.0, alpha=1e-4, generator=None):
    """botorch.optim.initializers.initialize_q_batch_nonneg: keep the best point, sample the others with weights
    exp(eta (Y / max - 1)) among the points with Y >= alpha max."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError('n cannot be larger than the number of provided samples')
    if n == n_samples:
        return X
    max_val, max_idx = torch.max(Y, dim=0)
    if bool(max_val <= 0):
        warnings.warn('All acquisition values for raw sampled points are nonpositive, so initial conditions are '
                      'being selected randomly.', BadInitialCandidatesWarning)
        return X[torch.randperm(n_samples, device=X.device, generator=generator)][:n]
    alpha_pos = Y >= alpha * max_val
    while int(alpha_pos.sum()) < n:
        alpha = 0.1 * alpha
        alpha_pos = Y >= alpha * max_val
    idcs_all = torch.arange(len(Y), device=Y.device)[alpha_pos]
    weights = torch.exp(eta * (Y[alpha_pos] / max_val - 1))
    idcs = idcs_all[torch.multinomial(weights, n, generator=generator)]
    if not bool((idcs == max_idx).any()):
        idcs[-1] = max_idx
    return X[idcs]


def initialize_q_batch(X, Y, n, eta=1.0, generator=None):
    """botorch.optim.initializers.initialize_q_batch (the branch manifold_optimize.py:277-281 takes for acquisition
    functions that are not known to be non-negative): keep the best point, sample the others with weights
    exp(eta Z), Z the standardised acquisition values (eta halved until the weights are finite)."""
    n_samples = X.shape[0]
    if n > n_samples:
        raise RuntimeError('n (%d) cannot be larger than the number of provided samples (%d)' % (n, n_samples))
    if n == n_samples:
        return X
    Ystd = Y.std()
    if bool(Ystd == 0):
        warnings.warn('All acquisition values for raw samples points are the same. Choosing initial conditions at '
                      'random.', BadInitialCandidatesWarning)
        return X[torch.randperm(n_samples, device=X.device, generator=generator)][:n]
    max_val, max_idx = torch.max(Y, dim=0)
    etaZ = eta * (Y - Y.mean()) / Ystd
    weights = torch.exp(etaZ)
    while bool(torch.isinf(weights).any()):
        etaZ = etaZ * 0.5
        weights = torch.exp(etaZ)
    idcs = torch.multinomial(weights, n, generator=generator)
    if not bool((idcs == max_idx).any()):
        idcs[-1] = max_idx
    return X[idcs]


def is_nonnegative(acq_function):
    """botorch.acquisition.utils.is_nonnegative: True for the acquisition classes known to be non-negative (of the ones
    this package evaluates: ExpectedImprovement); anything else is treated as signed."""
    return bool(getattr(acq_function, 'nonnegative', False)) or type(acq_function).__name__ in (
        'ExpectedImprovement', 'NoisyExpectedImprovement', 'ProbabilityOfImprovement', 'qExpectedImprovement',
        'qNoisyExpectedImprovement', 'qProbabilityOfImprovement')


def _rand_points(manifold, n, generator):
    if 'rand' in getattr(manifold, '__dict__', {}):
        # ``rand`` re-bound on the manifold OBJECT, the reference's way of installing a problem-specific sampler
        # (gabo_spd.py:102, hd_gabo_spd.py:236-239): honour it, one point per call
        import numpy as np
        return ops.to_dev64(np.stack([np.asarray(manifold.rand()) for _ in range(n)]))
    if hasattr(manifold, 'rand_batch'):
        return manifold.rand_batch(n, generator=generator)
    import numpy as np  # foreign (pymanopt) manifold object: its own sampler, one point per call as in the reference
    return ops.to_dev64(np.stack([manifold.rand() for _ in range(n)]))


def sharded_acq_values(acq_function, X, group=None):
    """Raw-sample screening sharded over the ranks: rank g evaluates the acquisition on its contiguous
    block of the samples, ONE all-gather of the values (a few KB) gives every rank the whole vector.  ``X`` must be
    identical on every rank (same seed); shards are padded to equal length for the collective."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = X.shape[0]
    lo, hi = shard_range(n, rank, world)
    width = -(-n // world) + 1                                   # >= the longest shard
    with torch.no_grad():
        vals = ops.to_dev64(acq_function(X[lo:hi])).reshape(-1) if hi > lo else X.new_zeros(0, dtype=torch.float64)
    buf = torch.zeros(width, dtype=torch.float64, device=vals.device)
    buf[:hi - lo] = vals
    out = torch.empty(world, width, dtype=torch.float64, device=vals.device)
    dist.all_gather_into_tensor(out, buf, group=group) if buf.is_cuda else \
        dist.all_gather(list(out.unbind(0)), buf, group=group)
    parts = []
    for g_ in range(world):
        a_, b_ = shard_range(n, g_, world)
        parts.append(out[g_, :b_ - a_])
    return torch.cat(parts)


def gen_batch_initial_conditions_manifold(acq_function, manifold, bounds, q, num_restarts, raw_samples,
                                          sample_type=torch.float64, options=None, post_processing_manifold=None):
    """``num_restarts x q x dvec`` starting points (manifold_optimize.py:232-321).  Only q = 1 (as the reference)."""
    options = options or {}
    if q is None:
        q = 1
    if q != 1:
        raise NotImplementedError('q != 1 is not handled (neither by the reference, manifold_optimize.py:206)')
    seed = options.get('seed')
    distributed = bool(options.get('distributed', False))
    if distributed and seed is None:
        # the ranks must draw the same raw samples: rank 0 picks the seed
        import torch.distributed as dist
        box = [int(torch.seed() % (2 ** 31)) if dist.get_rank() == 0 else 0]
        dist.broadcast_object_list(box, src=0)
        seed = box[0]
    gen = None
    if seed is not None:
        gen = torch.Generator(device=ops.device())
        gen.manual_seed(int(seed))
    batch_limit = options.get('batch_limit')
    init_kwargs = {}
    if 'eta' in options:
        init_kwargs['eta'] = options.get('eta')
    if options.get('nonnegative') or is_nonnegative(acq_function):       # manifold_optimize.py:277-281
        init_func = initialize_q_batch_nonneg
        if 'alpha' in options:
            init_kwargs['alpha'] = options.get('alpha')
    else:
        init_func = initialize_q_batch

    def acq_chunked(X):                                                   # manifold_optimize.py:297-307
        step = X.shape[0] if batch_limit is None else max(1, int(batch_limit))
        with torch.no_grad():
            parts = [ops.to_dev64(acq_function(X[lo:lo + step])).reshape(-1) for lo in range(0, X.shape[0], step)]
        return parts[0] if len(parts) == 1 else torch.cat(parts)
    factor, max_factor = 1, 5
    batch_initial_conditions = None
    while factor < max_factor:
        with warnings.catch_warnings(record=True) as ws:
            warnings.simplefilter('always')
            pts = _rand_points(manifold, raw_samples * factor * q, gen)          # (n, ...) on the device
            X_rnd = pts[:, None].to(sample_type)
            if post_processing_manifold is not None:
                X_rnd = post_processing_manifold(X_rnd)
            if distributed:
                Y_rnd = sharded_acq_values(acq_chunked, X_rnd)
            else:
                Y_rnd = acq_chunked(X_rnd)
            batch_initial_conditions = init_func(X_rnd, Y_rnd, num_restarts, generator=gen, **init_kwargs)
            if not any(issubclass(w.category, BadInitialCandidatesWarning) for w in ws):
                return batch_initial_conditions
            factor += 1
    warnings.warn('Unable to find non-zero acquisition function values - initial conditions are being selected '
                  'randomly.', BadInitialCandidatesWarning)
    return batch_initial_conditions


# ----------------------------------------------------------------------------------------------------------------
# candidates
# ----------------------------------------------------------------------------------------------------------------

def gen_candidates_manifold(initial_conditions, acquisition_function, manifold, solver, pre_processing_manifold=None,
                            post_processing_manifold=None, lower_bounds=None, upper_bounds=None,
                            inequality_constraints=None, equality_constraints=None, approx_hessian=False,
                            solver_init_conds=False, options=None, return_info=False):
    """All restarts solved together (manifold_optimize.py:124-228).  Returns ``(candidates, acquisition values)``
    with the shapes of the reference: ``R x 1 x dvec`` and ``R``.  Bounds are accepted and ignored, as in the
    reference (:131-132 are never used by its body).

    Recognised solvers: ``ConjugateGradient`` (one launch, ``gabo_acq_rcg``), ``TrustRegions`` (one launch on spheres
    the register kernel covers, ``gabo_acq_rtr``; lock-step over the batched kernels otherwise) and
    ``ConstrainedTrustRegions`` / ``StrictConstrainedTrustRegions`` with inequality OR equality constraints and
    ``AugmentedLagrangeMethod(inner_solver=TrustRegions(...))`` (lock-step).
    Hessian-vector products are always the finite differences of approximate_hessian.py (the reference's
    ``approx_hessian=True``); with ``approx_hessian=False`` the reference differentiates the gradient with autograd
    instead.  Anything else raises ``NotImplementedError``: there is no CPU fallback."""
    if type(solver).__name__ == 'AugmentedLagrangeMethod':
        return _gen_candidates_alm(initial_conditions, acquisition_function, manifold, solver, pre_processing_manifold,
                                   post_processing_manifold, inequality_constraints, equality_constraints, return_info)
    if equality_constraints is not None and inequality_constraints is not None:
        raise NotImplementedError('equality and inequality constraints together are not supported')
    if equality_constraints is not None:
        inequality_constraints, eq_mode = equality_constraints, True
    else:
        eq_mode = False
    if inequality_constraints is not None and type(solver).__name__ not in _CONSTRAINED_SOLVERS:
        raise NotImplementedError('constraints need [Strict]ConstrainedTrustRegions (the ALM solver is '
                                  'SURVEY 8f "next"); there is no CPU fallback')
    if solver_init_conds:
        raise NotImplementedError('solver-side initialisation (population methods) is not supported')
    kind = _manifold_kind(manifold)
    trust_region = type(solver).__name__ in ('TrustRegions',) + _CONSTRAINED_SOLVERS
    constrained = type(solver).__name__ in _CONSTRAINED_SOLVERS and inequality_constraints is not None
    sopts = _trust_region_options(solver) if trust_region else _solver_options(solver)
    if not isinstance(acquisition_function, ExpectedImprovement):
        raise NotImplementedError('the B200 optimiser evaluates ExpectedImprovement in closed form; got %s'
                                  % type(acquisition_function).__name__)
    gp = acquisition_function.device_gp()
    if gp.manifold != kind:
        raise ValueError('the manifold and the GP kernel live on different manifolds')
    x0 = torch.as_tensor(initial_conditions).detach()
    like = x0
    if pre_processing_manifold is not None:
        x0 = pre_processing_manifold(x0)
    x0 = ops.to_dev64(x0)
    if x0.dim() < 3 or x0.shape[1] != 1:
        raise NotImplementedError('initial_conditions must be R x 1 x ... (q = 1, manifold_optimize.py:206)')
    pts = x0[:, 0]
    tensor_gp = getattr(gp, 'is_tensor_gp', False)
    if not trust_region:
        if tensor_gp:
            raise NotImplementedError('log-Euclidean / Frobenius kernels: the acquisition is optimised by the trust-region '
                                      'solvers (TrustRegions, [Strict]ConstrainedTrustRegions), as in hd_gabo_spd.py')
        solve = ops.acq_rcg
    elif constrained:
        cons = inequality_constraints if isinstance(inequality_constraints, (list, tuple)) else [inequality_constraints]
        strict = type(solver).__name__ == 'StrictConstrainedTrustRegions'
        delta_cons = float(getattr(solver, 'Delta_cons', 1e-6))
        specs = None if (eq_mode or kind != _lib.SPD or tensor_gp) else eigenvalue_constraint_specs(cons)
        if specs is not None:
            # gabo_spd.py's configuration: eigenvalue constraints on SPD(d) -> the whole constrained solve in one launch
            solve = ops.acq_ctr
            sopts = dict(sopts, constraints=specs, strict=strict, delta_cons=delta_cons)
        else:
            # any other constraint callable (differentiated with torch.autograd per restart, as the reference's Problem
            # does) or equality constraints: lock-step driver with the constrained tCG, fp64 evaluator
            solve = batched_trust_regions
            gp = gp.with_compute(_lib.GABO_F64)
            sopts = dict(sopts, **{'eq_constraints' if eq_mode else 'ineq_constraints': batched_constraints(cons, kind)},
                         delta_cons=delta_cons, strict=strict)
    elif _rtr_kernel_covers(gp):
        solve = ops.acq_rtr                      # one launch, one warp per restart (SPD: fp64 whatever gp.compute says)
    else:
        # spheres beyond the register kernel: lock-step over the batched kernels, evaluated in fp64
        solve = batched_trust_regions
        gp = gp.with_compute(_lib.GABO_F64)
    cand, val, iters, reason = solve(gp, pts, **sopts)
    candidates = cand[:, None]
    if post_processing_manifold is not None:
        candidates = post_processing_manifold(candidates)
    if not like.is_cuda:
        candidates, val = candidates.to(like.device), val.to(like.device)
    if return_info:
        return candidates, val, dict(iters=iters, reason=reason)
    return candidates, val


def _gen_candidates_alm(initial_conditions, acquisition_function, manifold, solver, pre, post, ineq, eq, return_info):
    """``gen_candidates_manifold`` for ``AugmentedLagrangeMethod`` (equality and inequality constraints, also together)."""
    if not isinstance(acquisition_function, ExpectedImprovement):
        raise NotImplementedError('the B200 optimiser evaluates ExpectedImprovement in closed form; got %s'
                                  % type(acquisition_function).__name__)
    kind = _manifold_kind(manifold)
    gp = acquisition_function.device_gp()
    if gp.manifold != kind:
        raise ValueError('the manifold and the GP kernel live on different manifolds')
    gp = gp.with_compute(_lib.GABO_F64)
    x0 = torch.as_tensor(initial_conditions).detach()
    like = x0
    if pre is not None:
        x0 = pre(x0)
    x0 = ops.to_dev64(x0)
    if x0.dim() < 3 or x0.shape[1] != 1:
        raise NotImplementedError('initial_conditions must be R x 1 x ... (q = 1, manifold_optimize.py:206)')

    def as_batched(c):
        if c is None:
            return None
        return batched_constraints(c if isinstance(c, (list, tuple)) else [c], kind)
    if type(solver.inner_solver).__name__ != 'TrustRegions':
        raise NotImplementedError('AugmentedLagrangeMethod: the batched path drives TrustRegions as inner solver')
    inner = _trust_region_options(solver.inner_solver)
    cand, val, iters, reason = batched_alm(
        gp, x0[:, 0], inner, ineq_constraints=as_batched(ineq), eq_constraints=as_batched(eq),
        maxiter=int(solver._maxiter), minstepsize=float(solver._minstepsize), bound=float(solver._bound),
        rho_init=float(solver._rho_init), thetarho=float(solver._thetarho), tau=float(solver._tau),
        starting_tolgradnorm=float(solver._starting_tolgradnorm), ending_tolgradnorm=float(solver._ending_tolgradnorm),
        lambdas_fact=float(solver._lambdas_fact), gammas_fact=float(solver._gammas_fact))
    candidates = cand[:, None]
    if post is not None:
        candidates = post(candidates)
    if not like.is_cuda:
        candidates, val = candidates.to(like.device), val.to(like.device)
    if return_info:
        return candidates, val, dict(iters=iters, reason=reason)
    return candidates, val


def get_best_candidates(batch_candidates, batch_values):
    """botorch.gen.get_best_candidates: the candidate with the highest value (first index on ties, NaN loses)."""
    slot, _ = ops.argmax_records(batch_values)
    return batch_candidates[int(slot.item())]


def shard_range(num, rank, world):
    """Contiguous block partition of ``num`` restarts: rank g owns [g*num/G, (g+1)*num/G)."""
    return (rank * num) // world, ((rank + 1) * num) // world


def allgather_records(value, gidx, candidate, group=None):
    """ONE all-gather of fixed-size fp64 records ``[value, global index, candidate...]``.  Device-agnostic
    plumbing (NCCL on GPUs, gloo in the CPU tests).  Returns ``(values (G,), gidx (G,), candidates (G, ...))``."""
    import torch.distributed as dist
    flat = candidate.reshape(-1).to(torch.float64)
    rec = torch.cat([value.reshape(1).to(torch.float64), gidx.reshape(1).to(torch.float64), flat])
    world = dist.get_world_size(group)
    out = torch.empty(world, rec.numel(), dtype=torch.float64, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group) if rec.is_cuda else \
        dist.all_gather(list(out.unbind(0)), rec, group=group)
    return out[:, 0], out[:, 1].to(torch.int64), out[:, 2:].reshape((world,) + tuple(candidate.shape))


def joint_optimize_manifold(acq_function, manifold, solver, q, num_restarts, raw_samples, bounds=None,
                            sample_type=torch.float64, options=None, inequality_constraints=None,
                            equality_constraints=None, pre_processing_manifold=None, post_processing_manifold=None,
                            approx_hessian=False, solver_init_conds=False):
    """``q x dvec`` best candidate of a multi-start optimisation (manifold_optimize.py:36-120)."""
    options = options or {}
    distributed = bool(options.get('distributed', False))
    ics = gen_batch_initial_conditions_manifold(acq_function=acq_function, manifold=manifold, bounds=bounds, q=None,
                                                num_restarts=num_restarts, raw_samples=raw_samples,
                                                sample_type=sample_type, options=options,
                                                post_processing_manifold=post_processing_manifold)
    lo, hi = 0, num_restarts
    if distributed:
        import torch.distributed as dist
        # every rank must hold the same starts: take rank 0's
        dist.broadcast(ics, src=0)
        lo, hi = shard_range(num_restarts, dist.get_rank(), dist.get_world_size())
    sub = {k: v for k, v in options.items() if k not in ('batch_limit', 'nonnegative', 'distributed', 'seed')}
    # manifold_optimize.py:95-116: restarts are handed to gen_candidates_manifold in chunks of `batch_limit` (default: all
    # of them -- one launch); the chunks are independent, so the result does not depend on the chunking
    batch_limit = max(1, int(options.get('batch_limit', num_restarts) or num_restarts))
    cl, vl = [], []
    for start in range(lo, hi, batch_limit):
        c_, v_ = gen_candidates_manifold(initial_conditions=ics[start:min(start + batch_limit, hi)],
                                         acquisition_function=acq_function, manifold=manifold, solver=solver,
                                         pre_processing_manifold=pre_processing_manifold,
                                         post_processing_manifold=post_processing_manifold,
                                         lower_bounds=None if bounds is None else bounds[0],
                                         upper_bounds=None if bounds is None else bounds[1], options=sub,
                                         inequality_constraints=inequality_constraints,
                                         equality_constraints=equality_constraints, approx_hessian=approx_hessian,
                                         solver_init_conds=solver_init_conds)
        cl.append(c_)
        vl.append(v_)
    cands, vals = (cl[0], vl[0]) if len(cl) == 1 else (torch.cat(cl), torch.cat(vl))
    if not distributed:
        return get_best_candidates(cands, vals)
    gidx = torch.arange(lo, hi, device=vals.device)
    slot, best = ops.argmax_records(vals, gidx)
    s = int(slot.item())
    v, g, c = allgather_records(best.reshape(()), gidx[s], ops.to_dev64(cands[s]))
    win, _ = ops.argmax_records(v, g)
    return c[int(win.item())]
