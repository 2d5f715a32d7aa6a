"""GP fit with manifold-valued kernel parameters: ``fit_gpytorch_manifold``.

Mirrors ``BoManifolds/manifold_optimization/manifold_gp_fit.py:54-222`` of the reference: the negative marginal
log-likelihood is minimised over the PRODUCT of the parameter manifolds -- Euclidean for the raw noise, the constant
mean, the raw outputscale and the raw beta / lengthscale; ``Grassmann(D, d)`` for the projection matrix of the nested
SPD kernels and one ``Sphere(k)`` per axis of the nested-sphere kernel (the ``<name>_manifold`` attributes the reference
looks up with ``attrgetter``, :158-166) -- with a pymanopt solver (default ``ConjugateGradient(maxiter=500)``) started
from the best of ``nb_init_candidates`` random parameter sets, three quarters of which keep the current values of the
first four (Euclidean) parameters (:186-196).

What runs where.  The parameter manifolds are tiny (a 20 x 5 matrix, a few unit vectors, four scalars): their
retractions / projections are host numpy, exactly like pymanopt's.  Every objective evaluation is device work through the
package's kernels: projection -> factorisation -> geodesic distance matrix (``kernel_utils``), then ONE launch of
``gabo_gp_mll`` (Cholesky, solves, log-determinant, closed-form gradient of the four Euclidean parameters, ``alpha`` and
``K^-1``); the gradient of the manifold-valued parameters goes back through the distance matrix with the backward kernels
(``gabo_spd_ai_gram_backward``, ``gabo_nested_spd_project_backward``, ``gabo_weighted_points_sum``).  The reference gets
both from torch.autograd through its per-pair Python loops.

pymanopt (``Product``, ``Grassmann``, ``Sphere``, ``Euclidean``, ``ConjugateGradient``, ``LineSearchAdaptive``) is
third-party and absent from the image: PARITY UNPINNED for the trajectory (restated from pymanopt 0.2.x); the objective and
its gradients are checked against finite differences and the oracle.
"""
import math
import time

import numpy as np
import torch

from . import _lib, ops
from . import kernel_utils as ku
from .gp_fit import MarginalLogLikelihood, NOISE_MIN, _prior_of, _model_parts
from .manifold_optimization import ConjugateGradient, ManifoldGP


# ----------------------------------------------------------------------------------------------------------------
# host-side parameter manifolds (pymanopt 0.2.x formulas; points are small numpy arrays)
# ----------------------------------------------------------------------------------------------------------------

class EuclideanParam:
    """pymanopt ``Euclidean(n)``."""

    def __init__(self, *shape):
        self._shape = tuple(shape)

    @property
    def typicaldist(self):
        return math.sqrt(float(np.prod(self._shape)))

    @property
    def dim(self):
        return int(np.prod(self._shape))

    def rand(self):
        return np.random.randn(*self._shape)

    def inner(self, x, u, v):
        return float(np.tensordot(u, v, axes=u.ndim))

    def norm(self, x, u):
        return float(np.linalg.norm(u))

    def proj(self, x, u):
        return u

    egrad2rgrad = proj

    def retr(self, x, u):
        return x + u

    def transp(self, x, y, u):
        return u

    def dist(self, x, y):
        return float(np.linalg.norm(x - y))


class SphereParam:
    """pymanopt ``Sphere(n)`` (vectors): projection retraction, projection transport."""

    def __init__(self, n):
        self._n = int(n)

    typicaldist = math.pi

    @property
    def dim(self):
        return self._n - 1

    def rand(self):
        v = np.random.randn(self._n)
        return v / np.linalg.norm(v)

    def inner(self, x, u, v):
        return float(np.dot(u.ravel(), v.ravel()))

    def norm(self, x, u):
        return float(np.linalg.norm(u))

    def proj(self, x, u):
        return u - np.dot(x.ravel(), u.ravel()) * x

    egrad2rgrad = proj

    def retr(self, x, u):
        y = x + u
        return y / np.linalg.norm(y)

    def transp(self, x, y, u):
        return self.proj(y, u)

    def dist(self, x, y):
        return float(np.arccos(np.clip(np.dot(x.ravel(), y.ravel()), -1.0, 1.0)))


class GrassmannParam:
    """pymanopt ``Grassmann(n, p)`` (one subspace): orthonormal n x p representatives, horizontal projection
    U - X X^T U, polar retraction (SVD of X + U), projection transport."""

    def __init__(self, n, p):
        self._n, self._p = int(n), int(p)

    @property
    def typicaldist(self):
        return math.sqrt(self._p)

    @property
    def dim(self):
        return self._p * (self._n - self._p)

    def rand(self):
        q, _ = np.linalg.qr(np.random.randn(self._n, self._p))
        return q

    def inner(self, x, u, v):
        return float(np.tensordot(u, v, axes=2))

    def norm(self, x, u):
        return float(np.linalg.norm(u))

    def proj(self, x, u):
        return u - x @ (x.T @ u)

    egrad2rgrad = proj

    def retr(self, x, u):
        a, _, bt = np.linalg.svd(x + u, full_matrices=False)
        return a @ bt

    def transp(self, x, y, u):
        return self.proj(y, u)

    def dist(self, x, y):
        s = np.linalg.svd(x.T @ y, compute_uv=False)
        return float(np.linalg.norm(np.arccos(np.clip(s, None, 1.0))))


def _sym(a):
    return 0.5 * (a + a.T)


class SpdParam:
    """pymanopt 0.2.x ``PositiveDefinite(n)`` (k = 1): affine-invariant metric tr(X^-1 U X^-1 V), Riemannian gradient
    X sym(G) X, exponential map as the retraction, identity transport, random points U diag(1 + rand) U^T."""

    def __init__(self, n):
        self._n = int(n)

    @property
    def typicaldist(self):
        return math.sqrt(self._n * (self._n + 1) / 2.0)

    @property
    def dim(self):
        return self._n * (self._n + 1) // 2

    def rand(self):
        d = 1.0 + np.random.rand(self._n)
        u, _ = np.linalg.qr(np.random.randn(self._n, self._n))
        return (u * d) @ u.T

    def inner(self, x, u, v):
        return float(np.tensordot(np.linalg.solve(x, u), np.linalg.solve(x, v).T, axes=2))

    def norm(self, x, u):
        c = np.linalg.cholesky(x)
        w = np.linalg.solve(c, np.linalg.solve(c, u).T)       # L^-1 U L^-T
        return float(np.linalg.norm(w))

    def proj(self, x, u):
        return _sym(u)

    def egrad2rgrad(self, x, u):
        return x @ _sym(u) @ x

    def retr(self, x, u):
        c = np.linalg.cholesky(x)
        w = _sym(np.linalg.solve(c, np.linalg.solve(c, _sym(u)).T))
        lam, q = np.linalg.eigh(w)
        cq = c @ q
        return _sym((cq * np.exp(lam)) @ cq.T)

    def transp(self, x, y, u):
        return u

    def dist(self, x, y):
        c = np.linalg.cholesky(x)
        w = _sym(np.linalg.solve(c, np.linalg.solve(c, y).T))
        lam = np.linalg.eigvalsh(w)
        return float(np.linalg.norm(np.log(lam)))


class ProductParam:
    """pymanopt ``Product``: points and tangent vectors are lists, one entry per factor."""

    def __init__(self, manifolds):
        self.manifolds = list(manifolds)

    @property
    def typicaldist(self):
        return math.sqrt(sum(m.typicaldist ** 2 for m in self.manifolds))

    @property
    def dim(self):
        return sum(m.dim for m in self.manifolds)

    def rand(self):
        return [m.rand() for m in self.manifolds]

    def inner(self, x, u, v):
        return float(sum(m.inner(a, b, c) for m, a, b, c in zip(self.manifolds, x, u, v)))

    def norm(self, x, u):
        return math.sqrt(max(self.inner(x, u, u), 0.0))

    def proj(self, x, u):
        return [m.proj(a, b) for m, a, b in zip(self.manifolds, x, u)]

    def egrad2rgrad(self, x, u):
        return [m.egrad2rgrad(a, b) for m, a, b in zip(self.manifolds, x, u)]

    def retr(self, x, u):
        return [m.retr(a, b) for m, a, b in zip(self.manifolds, x, u)]

    def transp(self, x, y, u):
        return [m.transp(a, b, c) for m, a, b, c in zip(self.manifolds, x, y, u)]

    def dist(self, x, y):
        return math.sqrt(sum(m.dist(a, b) ** 2 for m, a, b in zip(self.manifolds, x, y)))


def _lin(a, x, b=None, y=None):
    """a x (+ b y) on lists of arrays."""
    if y is None:
        return [a * xi for xi in x]
    return [a * xi + b * yi for xi, yi in zip(x, y)]


def host_manifold(obj):
    """Host parameter manifold for what a kernel stores as ``<parameter>_manifold`` (the package's stubs, or a pymanopt
    ``Grassmann`` / ``Sphere`` / ``Euclidean`` object, recognised by class name and its ``_n`` / ``_p`` attributes)."""
    name = type(obj).__name__
    if isinstance(obj, (EuclideanParam, SphereParam, GrassmannParam, SpdParam, ProductParam)):
        return obj
    if 'PositiveDefinite' in name:
        return SpdParam(obj._n)
    if 'Grassmann' in name:
        return GrassmannParam(obj._n, obj._p)
    if 'Sphere' in name:
        return SphereParam(getattr(obj, '_n', None) or obj._shape[0])
    if 'Euclidean' in name:
        return EuclideanParam(*obj._shape)
    raise NotImplementedError('parameter manifold %s is not supported (Grassmann, Sphere, PositiveDefinite, Euclidean)' % name)


# ----------------------------------------------------------------------------------------------------------------
# pymanopt ConjugateGradient + LineSearchAdaptive on a (product) manifold, host loop over device evaluations
# ----------------------------------------------------------------------------------------------------------------

def riemannian_cg(manifold, cost, cost_grad, x0, solver=None, callback=None):
    """pymanopt 0.2.x ``ConjugateGradient.solve`` (Hestenes-Stiefel, ``orth_value = inf``, ``LineSearchAdaptive``) for a
    problem given by ``cost(x) -> float`` and ``cost_grad(x) -> (float, euclidean gradient)``.
    Returns (x, log) with log = {'iterations', 'stop', 'cost', 'gradnorm', 'costevals'}."""
    s = solver or ConjugateGradient()
    maxiter, mingradnorm = int(s._maxiter), float(s._mingradnorm)
    minstepsize, maxcostevals, maxtime = float(s._minstepsize), int(s._maxcostevals), float(s._maxtime)
    contraction, suff_decr = float(s.contraction_factor), float(s.suff_decr)
    ls_maxiter, init_step = int(s.ls_maxiter), float(s.initial_stepsize)
    t0 = time.time()
    x = x0
    f, eg = cost_grad(x)
    costevals = 1
    grad = manifold.egrad2rgrad(x, eg)
    gradnorm = manifold.norm(x, grad)
    gPg = manifold.inner(x, grad, grad)
    desc = _lin(-1.0, grad)
    it, stepsize, oldalpha, stop = 0, float('nan'), None, ''
    while True:
        if callback is not None:
            callback(it, f, gradnorm)
        # Solver._check_stopping_criterion(time0, iter = it + 1, gradnorm, stepsize, costevals)
        if time.time() - t0 >= maxtime:
            stop = 'maxtime'
        elif it + 1 >= maxiter:
            stop = 'maxiter'
        elif gradnorm < mingradnorm:
            stop = 'mingradnorm'
        elif stepsize < minstepsize:
            stop = 'minstepsize'
        elif costevals >= maxcostevals:
            stop = 'maxcostevals'
        if stop:
            break
        df0 = manifold.inner(x, grad, desc)
        if df0 >= 0:                                   # not a descent direction: restart from steepest descent
            desc = _lin(-1.0, grad)
            df0 = -gPg
        # LineSearchAdaptive.search
        norm_d = manifold.norm(x, desc)
        alpha = oldalpha if oldalpha is not None else init_step / norm_d
        newx = manifold.retr(x, _lin(alpha, desc))
        newf = cost(newx)
        evals = 1
        while newf > f + suff_decr * alpha * df0 and evals <= ls_maxiter:
            alpha *= contraction
            newx = manifold.retr(x, _lin(alpha, desc))
            newf = cost(newx)
            evals += 1
        if newf > f:
            alpha, newx = 0.0, x
        costevals += evals
        stepsize = alpha * norm_d
        oldalpha = alpha if evals == 2 else 2.0 * alpha
        newf, neweg = cost_grad(newx)
        costevals += 1
        newgrad = manifold.egrad2rgrad(newx, neweg)
        newgradnorm = manifold.norm(newx, newgrad)
        newgPg = manifold.inner(newx, newgrad, newgrad)
        oldgrad = manifold.transp(x, newx, grad)
        desc = manifold.transp(x, newx, desc)
        diff = _lin(1.0, newgrad, -1.0, oldgrad)
        ip_diff = manifold.inner(newx, newgrad, diff)
        den = manifold.inner(newx, diff, desc)
        try:
            beta = max(0.0, ip_diff / den)             # Hestenes-Stiefel; NaN -> 0 as Python's max(0, nan)
        except ZeroDivisionError:
            beta = 1.0
        desc = _lin(-1.0, newgrad, beta, desc)
        x, f, grad, gradnorm, gPg = newx, newf, newgrad, newgradnorm, newgPg
        it += 1
    return x, {'iterations': it, 'stop': stop, 'cost': f, 'gradnorm': gradnorm, 'costevals': costevals,
               'time': time.time() - t0}


def riemannian_trust_regions(manifold, cost, cost_grad, x0, solver=None):
    """pymanopt 0.2.x ``TrustRegions.solve`` (Steihaug-Toint truncated CG, no preconditioner, ``use_rand=False``) on a
    (product) manifold with numpy points, the Hessian-vector product being the finite difference of the gradient that the
    reference installs everywhere (``get_hessianfd``, approximate_hessian.py:11-62: step 2^-14 / |a|, transport back).
    ``solver``: a ``manifold_optimization.TrustRegions`` options holder.  Returns (x, log)."""
    from .manifold_optimization import TrustRegions
    s = solver or TrustRegions()
    maxiter, mingradnorm, maxtime = int(s._maxiter), float(s._mingradnorm), float(s._maxtime)
    kappa, theta, rho_prime, rho_reg_fact = float(s.kappa), float(s.theta), float(s.rho_prime), float(s.rho_regularization)
    mininner = int(getattr(s, 'mininner', 1))
    maxinner = int(manifold.dim if getattr(s, 'maxinner', None) is None else s.maxinner)
    delta_bar = float(manifold.typicaldist if getattr(s, 'Delta_bar', None) is None else s.Delta_bar)
    delta = float(delta_bar / 8 if getattr(s, 'Delta0', None) is None else s.Delta0)
    eps = float(np.spacing(1))
    t0 = time.time()

    def grad_at(x):
        f, eg = cost_grad(x)
        return f, manifold.egrad2rgrad(x, eg)

    def hess(x, gx, a):
        na = manifold.norm(x, a)
        if na < 1e-15:
            return _lin(0.0, a)
        c = 2.0 ** -14 / na
        x1 = manifold.retr(x, _lin(c, a))
        g1 = grad_at(x1)[1]
        return _lin(1.0 / c, manifold.transp(x1, x, g1), -1.0 / c, gx)

    x = x0
    fx, gx = grad_at(x)
    ng = manifold.norm(x, gx)
    k, stop = 0, ''
    while True:
        if time.time() - t0 >= maxtime:
            stop = 'maxtime'
        elif k >= maxiter:
            stop = 'maxiter'
        elif ng < mingradnorm:
            stop = 'mingradnorm'
        if stop:
            break
        # ---- truncated CG from eta = 0 ----
        eta, heta = _lin(0.0, gx), _lin(0.0, gx)
        r = gx
        e_pe, r_r = 0.0, manifold.inner(x, gx, gx)
        norm_r0 = math.sqrt(r_r)
        z_r, d_pd, e_pd = r_r, r_r, 0.0
        dlt = _lin(-1.0, r)
        model_value, inner_stop = 0.0, 'maxinner'
        for j in range(maxinner):
            hd = hess(x, gx, dlt)
            d_hd = manifold.inner(x, dlt, hd)
            alpha = z_r / d_hd if d_hd != 0 else float('inf')
            e_pe_new = e_pe + 2.0 * alpha * e_pd + alpha * alpha * d_pd
            if d_hd <= 0 or e_pe_new >= delta * delta:
                tau = (-e_pd + math.sqrt(e_pd * e_pd + d_pd * (delta * delta - e_pe))) / d_pd
                eta = _lin(1.0, eta, tau, dlt)
                heta = _lin(1.0, heta, tau, hd)
                inner_stop = 'negative curvature' if d_hd <= 0 else 'exceeded'
                break
            e_pe = e_pe_new
            new_eta = _lin(1.0, eta, alpha, dlt)
            new_heta = _lin(1.0, heta, alpha, hd)
            new_model = manifold.inner(x, new_eta, gx) + 0.5 * manifold.inner(x, new_eta, new_heta)
            if new_model >= model_value:
                inner_stop = 'model increased'
                break
            eta, heta, model_value = new_eta, new_heta, new_model
            r = _lin(1.0, r, alpha, hd)
            r_r = manifold.inner(x, r, r)
            norm_r = math.sqrt(r_r)
            if j + 1 >= mininner and norm_r <= norm_r0 * min(norm_r0 ** theta, kappa):
                inner_stop = 'target'
                break
            zold = z_r
            z_r = r_r
            beta = z_r / zold
            dlt = manifold.proj(x, _lin(-1.0, r, beta, dlt))
            e_pd = beta * (e_pd + alpha * d_pd)
            d_pd = z_r + beta * beta * d_pd
        # ---- accept / reject, radius update ----
        x_prop = manifold.retr(x, eta)
        f_prop = cost(x_prop)
        rhonum = fx - f_prop
        rhoden = -manifold.inner(x, gx, eta) - 0.5 * manifold.inner(x, heta, eta)
        reg = max(1.0, abs(fx)) * eps * rho_reg_fact
        rhonum, rhoden = rhonum + reg, rhoden + reg
        model_decreased = rhoden >= 0
        rho = rhonum / rhoden if rhoden != 0 else float('nan')
        if rho < 0.25 or not model_decreased or math.isnan(rho):
            delta /= 4.0
        elif rho > 0.75 and inner_stop in ('negative curvature', 'exceeded'):
            delta = min(2.0 * delta, delta_bar)
        if model_decreased and rho > rho_prime:
            x, fx = x_prop, f_prop
            fx, gx = grad_at(x)
            ng = manifold.norm(x, gx)
        k += 1
    return x, {'iterations': k, 'stop': stop, 'cost': fx, 'gradnorm': ng, 'time': time.time() - t0}


def solve_on_manifold(manifold, cost, cost_grad, x0, solver):
    """Dispatch on the solver options object (``ConjugateGradient`` or ``TrustRegions``), as ``solver.solve(problem, x=x0)``."""
    name = type(solver).__name__
    if name == 'ConjugateGradient':
        return riemannian_cg(manifold, cost, cost_grad, x0, solver)
    if name == 'TrustRegions':
        return riemannian_trust_regions(manifold, cost, cost_grad, x0, solver)
    raise NotImplementedError('solver %s is not supported here (ConjugateGradient, TrustRegions)' % name)


def riemannian_alm(manifold, cost, cost_grad, x0, solver, eq_constraints=None, ineq_constraints=None, lambdas=None,
                   gammas=None, rho=None):
    """The reference's ``AugmentedLagrangeMethod.solve`` (augmented_Lagrange_method.py:66-229, sub-problem :231-328; Liu &
    Boumal 2019) for ONE start on a (product) manifold with numpy points: the host loop of the reconstruction fit.

    ``cost(x) -> float`` and ``cost_grad(x) -> (float, Euclidean gradient)`` as for ``solve_on_manifold``; every
    constraint is a callable ``x -> (value, Euclidean gradient)`` (equalities hold at 0, inequalities at >= 0).  ``solver``
    is a ``manifold_optimization.AugmentedLagrangeMethod`` options holder whose inner solver (``TrustRegions`` or
    ``ConjugateGradient``) is run by ``solve_on_manifold`` with the tolerance schedule of the reference
    (``starting_tolgradnorm`` shrinking geometrically to ``ending_tolgradnorm`` over ``maxiter`` outer iterations).
    Returns (x, log)."""
    import copy
    eqs = [] if eq_constraints is None else (list(eq_constraints) if isinstance(eq_constraints, (list, tuple))
                                             else [eq_constraints])
    ineqs = [] if ineq_constraints is None else (list(ineq_constraints) if isinstance(ineq_constraints, (list, tuple))
                                                 else [ineq_constraints])
    bound, thetarho, tau = float(solver._bound), float(solver._thetarho), float(solver._tau)
    start_tol, end_tol = float(solver._starting_tolgradnorm), float(solver._ending_tolgradnorm)
    maxiter, minstepsize, maxtime = max(1, int(solver._maxiter)), float(solver._minstepsize), float(solver._maxtime)
    lambdas = (float(solver._lambdas_fact) * np.ones(len(ineqs))) if lambdas is None else np.array(lambdas, dtype=float)
    gammas = (float(solver._gammas_fact) * np.ones(len(eqs))) if gammas is None else np.array(gammas, dtype=float)
    rho = float(solver._rho_init if rho is None else rho)
    inner = copy.copy(solver.inner_solver)
    tolgradnorm = start_tol
    theta_tol = (end_tol / start_tol) ** (1.0 / maxiter)
    oldacc = float('inf')
    t0 = time.time()
    x = x0
    k, stop, inner_logs = 0, '', []
    while True:
        lam_k, gam_k, rho_k = lambdas.copy(), gammas.copy(), rho

        def sub_cost(p):
            f = cost(p)
            for c, l in zip(ineqs, lam_k):
                f += 0.5 * rho_k * max(0.0, l / rho_k - c(p)[0]) ** 2
            for c, g in zip(eqs, gam_k):
                f += 0.5 * rho_k * (g / rho_k + c(p)[0]) ** 2
            return f

        def sub_cost_grad(p):
            f, grad = cost_grad(p)
            grad = [np.array(g, dtype=float) for g in grad]
            for c, l in zip(ineqs, lam_k):
                v, gc = c(p)
                if l / rho_k - v > 0:
                    f += 0.5 * rho_k * (l / rho_k - v) ** 2
                    grad = [g + (v * rho_k - l) * np.asarray(h) for g, h in zip(grad, gc)]
            for c, g_mult in zip(eqs, gam_k):
                v, gc = c(p)
                f += 0.5 * rho_k * (g_mult / rho_k + v) ** 2
                grad = [g + (v * rho_k + g_mult) * np.asarray(h) for g, h in zip(grad, gc)]
            return f, grad

        inner._mingradnorm = tolgradnorm
        x_new, ilog = solve_on_manifold(manifold, sub_cost, sub_cost_grad, x, inner)
        inner_logs.append((ilog['iterations'], ilog['stop']))
        newacc = 0.0
        for i, c in enumerate(ineqs):
            v = c(x_new)[0]
            newacc = max(newacc, abs(max(-lambdas[i] / rho, v)))
            lambdas[i] = min(bound, max(lambdas[i] + rho * v, 0.0))
        for i, c in enumerate(eqs):
            v = c(x_new)[0]
            newacc = max(newacc, abs(v))
            gammas[i] = min(bound, max(-bound, gammas[i] + rho * v))
        if k == 0 or newacc > tau * oldacc:
            rho = rho / thetarho
        oldacc = newacc
        tolgradnorm = max(end_tol, tolgradnorm * theta_tol)
        k += 1
        stepsize = manifold.dist(x_new, x)
        x = x_new
        if time.time() - t0 >= maxtime:
            stop = 'maxtime'
        elif k >= maxiter:
            stop = 'maxiter'
        elif stepsize < minstepsize:
            stop = 'minstepsize'
        if tolgradnorm <= end_tol:
            stop = 'mingradnorm'
        if stop:
            break
    return x, {'iterations': k, 'stop': stop, 'cost': cost(x), 'violation': oldacc, 'rho': rho,
               'lambdas': lambdas, 'gammas': gammas, 'inner': inner_logs, 'time': time.time() - t0}


# ----------------------------------------------------------------------------------------------------------------
# the objective
# ----------------------------------------------------------------------------------------------------------------

def _manifold_parameters(base):
    """[(name, parameter, host manifold)] for the parameters of the base kernel that carry a ``<name>_manifold``
    attribute, in registration order (manifold_gp_fit.py:158-166 walks ``named_parameters`` the same way)."""
    out = []
    for name, p in base.named_parameters():
        if not p.requires_grad:
            continue
        man = getattr(base, name + '_manifold', None)
        if man is not None:
            out.append((name, p, host_manifold(man)))
    return out


def _distance_matrix(base, x):
    """Dm with base kernel = exp(-theta Dm), through the differentiable operator chain of ``kernel_utils`` (under
    ``torch.no_grad()`` the same chain simply runs forward, so objective-only and gradient evaluations see bit-identical
    distances -- the fp32 tail of the affine-invariant distance differs by one rounding between the mirrored and the
    general tile path of the Gram kernel, which would otherwise show up in the line search)."""
    comp = _lib.GABO_F64
    if isinstance(base, ku.NestedSpdAffineInvariantGaussianKernel):
        y = base._project(x, autograd=True)
        return ku._SpdAiDistance2.apply(y, y, comp)
    if isinstance(base, ku.NestedSpdLogEuclideanGaussianKernel):
        a = ku._SpdLogm.apply(ku._MandelUnpack.apply(base._project(x, autograd=True)))
        return ku._FrobeniusDistance2.apply(a, a)
    if isinstance(base, ku.NestedSphereGaussianKernel):
        q = ku._nested_sphere_project_autograd(x, base.axes, base.distances_to_axis)
        d = ku._SphereDistance.apply(q, q)
        return d * d
    from .gp_fit import kernel_distance_matrix
    return kernel_distance_matrix(base, x)[0]


def _kernel_parameter(base):
    """(lower bound, prior, getter, setter) of the scalar the base kernel multiplies its distance matrix with:
    ``beta >= beta_min`` for the beta kernels, ``1 / lengthscale^2`` for the lengthscale kernels (fitted through
    ``raw_lengthscale`` there)."""
    if hasattr(base, 'raw_beta'):
        return 'beta'
    return 'lengthscale'


class ManifoldObjective:
    """-(log-likelihood + log-priors) / n of a ``ManifoldGP`` as a function of the parameter LIST
    [raw_noise, mean, raw_outputscale, raw_beta | raw_lengthscale, *manifold-valued kernel parameters]
    (the order of ``mll.named_parameters()`` in the reference's model) together with its Euclidean gradient."""

    def __init__(self, model):
        if not isinstance(model, ManifoldGP):
            raise NotImplementedError('fit_gpytorch_manifold expects a gabotorch_b200.ManifoldGP (or an mll holding one)')
        self.model = model
        self.cov, self.base, self.scaled = _model_parts(model)
        self.kind = _kernel_parameter(self.base)
        x = model.train_inputs[0]
        self.x = ops.to_dev64(x.reshape(-1, x.shape[-1]))
        self.mparams = _manifold_parameters(self.base)
        if self.kind == 'beta':
            lower, prior = float(self.base.beta_min), _prior_of(self.base, 'beta_prior')
        else:
            lower, prior = 0.0, _prior_of(self.base, 'lengthscale_prior')
        self.inner = MarginalLogLikelihood(
            torch.zeros(1, 1, dtype=torch.float64), model.train_targets, lower, getattr(model, 'noise_min', NOISE_MIN),
            outputscale_prior=_prior_of(self.cov, 'outputscale_prior') if self.scaled else None,
            noise_prior=getattr(model, 'noise_prior', None), beta_prior=prior)
        self.n = self.inner.n
        self.evaluations = 0
        self.manifold = ProductParam([EuclideanParam(1)] * 4 + [m for _, _, m in self.mparams])

    # -- parameter list <-> model ---------------------------------------------------------------------------
    def current(self):
        """The model's parameters as the list the solver works on."""
        if self.kind == 'beta':
            kp = float(self.base.beta.detach().reshape(-1)[0])
        else:
            ls = float(self.base.lengthscale.detach().reshape(-1)[0])
            kp = 1.0 / (ls * ls)
        s = float(self.cov.outputscale.detach()) if self.scaled else 1.0
        noise = max(self.model.noise, self.inner.noise_min * (1 + 1e-6) + 1e-300)
        raw = self.inner.inverse_transform((self._kp_to_internal(kp), s, noise, self.model.mean))
        # order: raw_noise, mean, raw_outputscale, raw kernel parameter
        out = [np.array([raw[2]]), np.array([raw[3]]), np.array([raw[1]]), np.array([raw[0]])]
        for _, p, _ in self.mparams:
            a = p.detach().cpu().double().numpy()
            out.append(a.copy() if (a.ndim > 1 and a.shape[0] > 1) else a.reshape(-1).copy())
        return out

    def _kp_to_internal(self, kp):
        # the inner objective's first slot is "lower + softplus(raw)"; for lengthscale kernels it holds the lengthscale
        return kp if self.kind == 'beta' else 1.0 / math.sqrt(kp)

    def _raw4(self, x):
        return np.array([float(x[3][0]), float(x[2][0]), float(x[0][0]), float(x[1][0])])   # inner order: kp, s, noise, mean

    def _theta(self, raw4):
        kp, s, noise, mean = self.inner.transform(raw4)
        if self.kind != 'beta':
            kp = 1.0 / (kp * kp)                              # lengthscale -> 1 / lengthscale^2
        return kp, s, noise, mean

    def _set_manifold_params(self, x):
        with torch.no_grad():
            for (name, p, _), val in zip(self.mparams, x[4:]):
                p.copy_(torch.as_tensor(np.asarray(val), dtype=p.dtype).reshape(p.shape))

    def apply(self, x):
        """Write the parameter list into the model (what ``set_params_with_list_of_array`` does in the reference)."""
        self._set_manifold_params(x)
        raw4 = self._raw4(x)
        kp, s, noise, mean = self.inner.transform(raw4)
        if self.kind == 'beta':
            self.base.beta = kp
        else:
            self.base.lengthscale = kp
        if self.scaled:
            self.cov.outputscale = s
        self.model.noise, self.model.mean = noise, mean

    # -- evaluations ----------------------------------------------------------------------------------------
    def _prior(self, raw4):
        kp, s, noise, mean = self.inner.transform(raw4)
        return self.inner._prior_terms((kp, s, noise, mean))

    def cost_device(self, x):
        """Objective as a 0-d DEVICE tensor without synchronising (candidate screening reads all values back at once)."""
        self._set_manifold_params(x)
        raw4 = self._raw4(x)
        with torch.no_grad():
            dm = _distance_matrix(self.base, self.x)
        theta = torch.tensor([self._theta(raw4)], dtype=torch.float64)
        ll, _, _, _, flags = ops.gp_mll(dm, self.inner.y, theta, want_grad=False)
        self.evaluations += 1
        lp = self._prior(raw4)[0]
        val = -(ll[0] + lp) / self.n
        return torch.where((flags[0] != 0) | ~torch.isfinite(val), torch.full_like(val, float('inf')), val)

    def cost(self, x):
        return float(self.cost_device(x))

    def cost_grad(self, x):
        """(objective, Euclidean gradient list): one forward through the kernels, one ``gabo_gp_mll`` launch, one
        backward through the distance matrix."""
        self._set_manifold_params(x)
        raw4 = self._raw4(x)
        kp, s, noise, mean = self._theta(raw4)
        with torch.enable_grad():
            dm = _distance_matrix(self.base, self.x)
        theta = torch.tensor([[kp, s, noise, mean]], dtype=torch.float64)
        ll, g4, alpha, kinv, flags = ops.gp_mll(dm.detach(), self.inner.y, theta, want_grad=True, want_factors=True)
        self.evaluations += 1
        head = torch.cat([ll, g4.reshape(-1), flags.double()]).cpu().numpy()
        zero = [np.zeros(1)] * 4 + [np.zeros_like(np.asarray(v, dtype=np.float64)) for v in x[4:]]
        if head[5] != 0 or not np.isfinite(head[0]):
            return 1e10, zero
        t_in = self.inner.transform(raw4)
        lp, dlp = self.inner._prior_terms(t_in)
        g = head[1:5].copy()                                   # d ll / d (kp, s, noise, mean)
        if self.kind != 'beta':
            g[0] = g[0] * (-2.0 / t_in[0] ** 3)                # d (1 / l^2) / d l
        from .gp_fit import _sigmoid
        for i in range(3):
            g[i] = (g[i] + dlp[i]) * _sigmoid(raw4[i])
        g = -g / self.n
        grads = [np.array([g[2]]), np.array([g[3]]), np.array([g[1] if self.scaled else 0.0]), np.array([g[0]])]
        if self.mparams and dm.requires_grad:
            # d ll / d Dm_ij = 1/2 (alpha alpha^T - Ktilde^-1)_ij * (-kp s exp(-kp Dm_ij))
            a = alpha[0]
            gd = 0.5 * (torch.outer(a, a) - kinv[0]) * (-kp * s) * torch.exp(-kp * dm.detach())
            for _, p, _ in self.mparams:
                p.grad = None
            dm.backward(-gd / self.n)
            for (name, p, _), val in zip(self.mparams, x[4:]):
                gp = p.grad.detach().cpu().double().numpy() if p.grad is not None else np.zeros(tuple(p.shape))
                grads.append(gp.reshape(np.asarray(val).shape))
                p.grad = None
        return -(head[0] + lp) / self.n, grads


# ----------------------------------------------------------------------------------------------------------------
# the entry point
# ----------------------------------------------------------------------------------------------------------------

def fit_gpytorch_manifold(mll, bounds=None, solver=None, nb_init_candidates=200, last_x_as_candidate_prob=0.9,
                          options=None, track_iterations=True, approx_mll=False, **kwargs):
    """Drop-in for the reference's ``fit_gpytorch_manifold`` (manifold_gp_fit.py:54-222) on a ``ManifoldGP`` (or an
    object with ``.model``): fits noise, mean, outputscale, beta / lengthscale AND the manifold-valued kernel parameters
    in place.  Returns ``(mll, info_dict)`` with ``fopt``, ``wall_time``, ``opt_log`` (and ``iterations`` when
    ``track_iterations``)."""
    if bounds is not None:
        import warnings
        warnings.warn('Bounds handling not supported yet in fit_gpytorch_manifold')      # as in the reference (:108-110)
    if approx_mll:
        raise NotImplementedError('approx_mll=True (gpytorch stochastic log-determinant) is not provided: the exact '
                                  'marginal likelihood is one kernel launch')
    solver = solver or ConjugateGradient(maxiter=500)
    if type(solver).__name__ != 'ConjugateGradient':
        raise NotImplementedError('fit_gpytorch_manifold drives ConjugateGradient (the reference default), got %s'
                                  % type(solver).__name__)
    model = getattr(mll, 'model', mll)
    t1 = time.time()
    obj = ManifoldObjective(model)
    x0 = obj.current()
    man = obj.manifold
    nb = int(nb_init_candidates)
    # initial candidates (:186-196): x0 with probability last_x_as_candidate_prob, random draws otherwise; the first
    # three quarters keep the current Euclidean hyper-parameters x0[0:4]
    if np.random.rand() < last_x_as_candidate_prob:
        cands = [x0] + [man.rand() for _ in range(nb - 1)]
    else:
        cands = [man.rand() for _ in range(nb)]
    for i in range(int(3 * nb / 4)):
        cands[i][0:4] = [v.copy() for v in x0[0:4]]
    vals = torch.stack([obj.cost_device(c) for c in cands]).cpu().numpy()   # ONE read-back for all candidates
    x_init = cands[int(np.argmin(vals))]
    iterations = []
    cb = (lambda it, f, gn: iterations.append((it, float(f), time.time() - t1))) if track_iterations else None
    opt_x, log = riemannian_cg(man, obj.cost, obj.cost_grad, x_init, solver, callback=cb)
    fopt = obj.cost(opt_x)
    obj.apply(opt_x)
    model.fit_result = {'objective': fopt, 'iterations': log['iterations'], 'evaluations': obj.evaluations,
                        'stop': log['stop'], 'gradnorm': log['gradnorm']}
    info = {'fopt': fopt, 'wall_time': time.time() - t1, 'opt_log': log}
    if track_iterations:
        info['iterations'] = iterations
    return mll, info
