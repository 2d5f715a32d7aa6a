"""Oracle (CPU, fp64) for the augmented Lagrangian acquisition solver.  Test infrastructure only.

Restates the reference's OWN ``AugmentedLagrangeMethod`` (``BoManifolds/manifold_optimization/augmented_Lagrange_method.py``:
``solve`` ``:66-203`` -- multiplier updates ``:160-169``, penalty update ``:171-172``, tolerance schedule ``:141-142,
:175``; ``subproblem_alm`` ``:205-328`` -- augmented cost / gradient, and the finite-difference Hessian bound to the
ORIGINAL problem ``:322``, i.e. it differentiates the gradient of the plain cost) around the reference's own
``TrustRegions`` inner solver (oracle/rtr.py) -- the configuration of
``examples/bo_sphere/constrained_benchmark_examples/gabo_sphere_inequality_constraints.py:251-256``.

PINNED on the reference's code: ``tests/golden/make_golden.py`` runs the reference's class itself (``alm_*`` arrays).
"""
import numpy as np

from . import rtr as _rtr


def solve_alm(gp, x0, eq_constraints=(), ineq_constraints=(), maxiter=1000, minstepsize=1e-10, inner_opts=None,
              bound=20.0, rho_init=1.0, thetarho=0.3, tau=0.8, starting_tolgradnorm=1e-3, ending_tolgradnorm=1e-6,
              lambdas_fact=1.0, gammas_fact=1.0):
    """Constraints are (value, Riemannian gradient) pairs (oracle/ctr.py).  Returns (x, outer iterations)."""
    man = _rtr._Man(gp.manifold, np.asarray(x0))
    ei_cost, ei_grad = _rtr.ei_problem(gp)
    x = np.array(x0, dtype=np.float64)
    xbest_prev = x
    lambdas = lambdas_fact * np.ones(len(ineq_constraints))
    gammas = gammas_fact * np.ones(len(eq_constraints))
    rho = rho_init
    oldacc = np.inf
    tolgradnorm = starting_tolgradnorm
    theta_tol = (ending_tolgradnorm / starting_tolgradnorm) ** (1.0 / maxiter)
    k = 0
    while True:
        lam, gam, r = lambdas.copy(), gammas.copy(), rho

        def cost(p):
            c = ei_cost(p)
            for i, con in enumerate(ineq_constraints):
                c += r / 2.0 * max(0.0, lam[i] / r - con[0](p)) ** 2
            for i, con in enumerate(eq_constraints):
                c += r / 2.0 * (gam[i] / r + con[0](p)) ** 2
            return c

        def grad(p):
            g = ei_grad(p).copy()
            for i, con in enumerate(ineq_constraints):
                v = con[0](p)
                if lam[i] / r - v > 0:
                    g += (v * r - lam[i]) * con[1](p)
            for i, con in enumerate(eq_constraints):
                g += (con[0](p) * r + gam[i]) * con[1](p)
            return g
        opts = _rtr.TROptions(**{**(inner_opts or {}), 'mingradnorm': tolgradnorm})
        xbest, _, _ = _rtr.solve_tr(gp, x, opts, cost=cost, grad=grad, hess_grad=ei_grad)
        newacc = 0.0
        for i, con in enumerate(ineq_constraints):
            v = con[0](xbest)
            newacc = max(newacc, abs(max(-lambdas[i] / rho, v)))
            lambdas[i] = min(bound, max(lambdas[i] + rho * v, 0.0))
        for i, con in enumerate(eq_constraints):
            v = con[0](xbest)
            newacc = max(newacc, abs(v))
            gammas[i] = min(bound, max(-bound, gammas[i] + rho * v))
        if k == 0 or newacc > tau * oldacc:
            rho = rho / thetarho
        oldacc = newacc
        tolgradnorm = max(ending_tolgradnorm, tolgradnorm * theta_tol)
        k += 1
        # Solver._check_stopping_criterion(time0, iter=k, stepsize=man.dist(xbest, xbest_prev))
        if k >= maxiter or man.dist(xbest, xbest_prev) < minstepsize or tolgradnorm <= ending_tolgradnorm:
            break
        x = xbest
        xbest_prev = xbest
    return xbest, k
