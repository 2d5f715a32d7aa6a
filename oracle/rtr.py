"""Oracle (CPU, fp64) for the trust-region acquisition solver.  Test infrastructure only.

Restates the reference's OWN solver, ``BoManifolds/manifold_optimization/robust_trust_regions.py`` (``TrustRegions.solve``
``:116-352``: radius update ``:271-300``, acceptance ``:303-311``; ``_truncated_conjugate_gradient`` ``:410-520`` with
the ``d_Hd != 0`` guard ``:461-465``), with the finite-difference Hessian of
``manifold_optimization/approximate_hessian.py:11-62`` (``epsilon = 2**-14``, retraction + projection transport) that
``gen_candidates_manifold`` installs for ``approx_hessian=True`` (``manifold_optimize.py:199-200``).

PINNED on the reference's code: ``tests/golden/make_golden.py`` runs the reference's ``TrustRegions`` class itself (the
third-party base class ``pymanopt.solvers.solver.Solver`` -- stopping rules only -- stubbed from its published
definition, see ``reference_loader.load_trust_regions``) on the oracle's EI problem and stores the trajectories
(``rtr_*`` arrays); ``tests/test_oracle_golden.py`` compares this restatement against them.
"""
from dataclasses import dataclass

import numpy as np

from . import gp as _gp
from . import sphere as _sph
from . import spd as _spd

NEGATIVE_CURVATURE, EXCEEDED_TR, REACHED_TARGET_LINEAR, REACHED_TARGET_SUPERLINEAR, MAX_INNER_ITER, MODEL_INCREASED = \
    range(6)


@dataclass
class TROptions:
    maxiter: int = 1000            # pymanopt Solver defaults
    mingradnorm: float = 1e-6
    kappa: float = 0.1             # robust_trust_regions.py:95-96
    theta: float = 1.0
    rho_prime: float = 0.1
    rho_regularization: float = 1e3
    mininner: int = 1              # solve(): mininner=1, maxinner=manifold.dim
    maxinner: int = None
    delta_bar: float = None        # manifold.typicaldist
    delta0: float = None           # delta_bar / 8
    fd_epsilon: float = 2.0 ** -14  # approximate_hessian.py:43


class _Man:
    def __init__(self, name, x0):
        m = _sph if name == 'sphere' else _spd
        self.inner, self.norm, self.retr, self.transp = m.inner, m.norm, m.retr, m.transp
        self.zerovec = lambda x: np.zeros(np.shape(x))
        self.dist = m.dist
        if name == 'sphere':
            self.dim = x0.shape[-1] - 1            # pymanopt Sphere.dim
            self.typicaldist = np.pi               # pymanopt Sphere.typicaldist
        else:
            d = x0.shape[-1]
            self.dim = d * (d + 1) // 2            # pymanopt PositiveDefinite.dim
            self.typicaldist = np.sqrt(self.dim)   # pymanopt PositiveDefinite.typicaldist


def ei_problem(gp):
    """cost = -EI and its Riemannian gradient (manifold_optimize.py:178-186 with the closed-form gradient)."""
    def cost(x):
        return -_gp.ei_and_grad(gp, x, want_grad=False)[0]

    def grad(x):
        return -_gp.ei_and_grad(gp, x, want_grad=True)[1]
    return cost, grad


def hessian_fd(man, grad, x, a, epsilon=2.0 ** -14):
    """approximate_hessian.py:11-62."""
    norm_a = man.norm(x, a)
    g = grad(x)
    if norm_a < 1e-15:
        return np.zeros(g.shape)
    c = epsilon / norm_a
    x1 = man.retr(x, c * a)
    g1 = man.transp(x1, x, grad(x1))
    return g1 / c - g / c


def truncated_cg(man, hess, x, fgradx, delta_radius, theta, kappa, mininner, maxinner):
    """robust_trust_regions.py:410-520 (use_rand=False, identity preconditioner)."""
    inner = man.inner
    eta = np.zeros_like(fgradx)
    heta = np.zeros_like(fgradx)
    r = fgradx
    e_pe = 0.0
    r_r = inner(x, r, r)
    norm_r0 = np.sqrt(r_r)
    z = r
    z_r = inner(x, z, r)
    d_pd = z_r
    delta = -z
    e_pd = 0.0
    model_value = 0.0
    stop = MAX_INNER_ITER
    j = 0
    for j in range(int(maxinner)):
        hdelta = hess(x, delta)
        d_hd = inner(x, delta, hdelta)
        if d_hd != 0:
            alpha = z_r / d_hd
            e_pe_new = e_pe + 2 * alpha * e_pd + alpha ** 2 * d_pd
        else:
            e_pe_new = e_pe
        if d_hd <= 0 or e_pe_new >= delta_radius ** 2:
            tau = (-e_pd + np.sqrt(e_pd * e_pd + d_pd * (delta_radius ** 2 - e_pe))) / d_pd
            eta = eta + tau * delta
            heta = heta + tau * hdelta
            stop = NEGATIVE_CURVATURE if d_hd <= 0 else EXCEEDED_TR
            break
        e_pe = e_pe_new
        new_eta = eta + alpha * delta
        new_heta = heta + alpha * hdelta
        new_model_value = inner(x, new_eta, fgradx) + 0.5 * inner(x, new_eta, new_heta)
        if new_model_value >= model_value:
            stop = MODEL_INCREASED
            break
        eta, heta, model_value = new_eta, new_heta, new_model_value
        r = r + alpha * hdelta
        r_r = inner(x, r, r)
        norm_r = np.sqrt(r_r)
        if j >= mininner and norm_r <= norm_r0 * min(norm_r0 ** theta, kappa):
            stop = REACHED_TARGET_LINEAR if kappa < norm_r0 ** theta else REACHED_TARGET_SUPERLINEAR
            break
        z = r
        zold_rold = z_r
        z_r = inner(x, z, r)
        beta = z_r / zold_rold
        delta = -z + beta * delta
        e_pd = beta * (e_pd + alpha * d_pd)
        d_pd = z_r + beta * beta * d_pd
    return eta, heta, j, stop


def solve_tr(gp, x0, opts=None, trace=None, cost=None, grad=None, hess_grad=None):
    """One ``TrustRegions.solve`` on cost = -EI with the finite-difference Hessian.  Returns (x, cost, iters).
    ``cost`` / ``grad`` replace the EI problem (the augmented Lagrangian of oracle/alm.py); ``hess_grad`` is the gradient
    the finite-difference Hessian differentiates (the ALM subproblem binds ``get_hessianfd`` to the ORIGINAL problem,
    augmented_Lagrange_method.py:322)."""
    opts = opts or TROptions()
    x = np.array(x0, dtype=np.float64)
    man = _Man(gp.manifold, x)
    ei_cost, ei_grad = ei_problem(gp)
    cost = cost or ei_cost
    grad = grad or ei_grad
    hess_grad = hess_grad or grad
    maxinner = man.dim if opts.maxinner is None else opts.maxinner
    delta_bar = man.typicaldist if opts.delta_bar is None else opts.delta_bar
    delta0 = delta_bar / 8 if opts.delta0 is None else opts.delta0

    def hess(p, a):
        return hessian_fd(man, hess_grad, p, a, opts.fd_epsilon)

    k = 0
    fx = cost(x)
    fgradx = grad(x)
    norm_grad = man.norm(x, fgradx)
    radius = delta0
    while True:
        if trace is not None:
            trace.append((k, x.copy(), fx, norm_grad, radius))
        eta, heta, _, stop_inner = truncated_cg(man, hess, x, fgradx, radius, opts.theta, opts.kappa, opts.mininner,
                                                maxinner)
        x_prop = man.retr(x, eta)
        fx_prop = cost(x_prop)
        rhonum = fx - fx_prop
        rhoden = -man.inner(x, fgradx, eta) - 0.5 * man.inner(x, eta, heta)
        rho_reg = max(1, abs(fx)) * np.spacing(1) * opts.rho_regularization
        rhonum = rhonum + rho_reg
        rhoden = rhoden + rho_reg
        model_decreased = rhoden >= 0
        rho = rhonum / rhoden if rhoden != 0 else np.nan
        if rho < 0.25 or not model_decreased or np.isnan(rho):
            radius = radius / 4
        elif rho > 0.75 and stop_inner in (NEGATIVE_CURVATURE, EXCEEDED_TR):
            radius = min(2 * radius, delta_bar)
        if model_decreased and rho > opts.rho_prime:
            x = x_prop
            fx = fx_prop
            fgradx = grad(x)
            norm_grad = man.norm(x, fgradx)
        k += 1
        # Solver._check_stopping_criterion(time0, gradnorm=norm_grad, iter=k)
        if k >= opts.maxiter or norm_grad < opts.mingradnorm:
            break
    return x, fx, k


def gen_candidates(gp, x0s, opts=None):
    """manifold_optimize.py:207-227 with the trust-region solver: candidates and their acquisition values."""
    xs, vals, its = [], [], []
    for x0 in x0s:
        x, c, k = solve_tr(gp, x0, opts)
        xs.append(x)
        vals.append(-c)
        its.append(k)
    return np.array(xs), np.array(vals), np.array(its)
