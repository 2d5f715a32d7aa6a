"""Oracle (CPU, fp64) for the multi-start Riemannian acquisition optimiser.  Test infrastructure only.

Restates ``BoManifolds/manifold_optimization/manifold_optimize.py`` of the reference:
``gen_candidates_manifold`` (``:124-228``: one ``solver.solve`` per restart, serial loop ``:207-221``, cost
``-acq`` ``:178-186``, acquisition re-evaluated at the candidates ``:227``) and ``joint_optimize_manifold``
(``:36-120``: argmax over restarts through botorch ``get_best_candidates``).

The solver itself is pymanopt 0.2.x ``ConjugateGradient`` (Hestenes-Stiefel, ``orth_value=inf``) with
``LineSearchAdaptive`` and ``Solver._check_stopping_criterion`` -- third-party, absent from
``/root/reference``: PARITY UNPINNED, restated from the published algorithm.
"""
from dataclasses import dataclass

import numpy as np

from . import gp as _gp
from . import sphere as _sph
from . import spd as _spd


@dataclass
class CGOptions:
    maxiter: int = 1000
    mingradnorm: float = 1e-6
    minstepsize: float = 1e-10
    maxcostevals: int = 5000
    # LineSearchAdaptive
    contraction_factor: float = 0.5
    suff_decr: float = 0.5
    ls_maxiter: int = 10
    initial_stepsize: float = 1.0


class _Man:
    def __init__(self, name):
        m = _sph if name == 'sphere' else _spd
        self.inner, self.norm, self.retr, self.transp = m.inner, m.norm, m.retr, m.transp


def solve_cg(gp, x0, opts=None, trace=None):
    """One pymanopt ``ConjugateGradient.solve`` on cost(x) = -EI(x).  Returns (x, cost, iters, stop_reason)."""
    opts = opts or CGOptions()
    man = _Man(gp.manifold)

    def objective(x):
        # A trial point that is numerically not SPD makes the reference raise inside torch.cholesky
        # (spd_utils_torch.py:87).  The GPU path rejects such a trial instead (cost = +inf, so the line search
        # backtracks); the oracle states the same rule.
        try:
            v = -_gp.ei_and_grad(gp, x, want_grad=False)[0]
        except np.linalg.LinAlgError:
            return np.inf
        return v if np.isfinite(v) else np.inf

    def cost_grad(x):
        ei, g = _gp.ei_and_grad(gp, x, want_grad=True)
        return -ei, -g

    x = np.array(x0, dtype=np.float64)
    it = 0
    stepsize = np.nan
    oldalpha = None
    cost, grad = cost_grad(x)
    gradnorm = float(man.norm(x, grad))
    gradPgrad = float(man.inner(x, grad, grad))
    desc = -grad
    reason = 0
    while True:
        if trace is not None:
            trace.append((it, x.copy(), cost, gradnorm))
        # Solver._check_stopping_criterion(iter=iter+1, gradnorm, stepsize)
        if it + 1 >= opts.maxiter:
            reason = 1
            break
        if gradnorm < opts.mingradnorm:
            reason = 2
            break
        if stepsize < opts.minstepsize:
            reason = 3
            break
        df0 = float(man.inner(x, grad, desc))
        if df0 >= 0:
            desc = -grad
            df0 = -gradPgrad
        # LineSearchAdaptive.search
        norm_d = float(man.norm(x, desc))
        alpha = oldalpha if oldalpha is not None else opts.initial_stepsize / norm_d
        newx = man.retr(x, alpha * desc)
        newf = objective(newx)
        evals = 1
        while newf > cost + opts.suff_decr * alpha * df0 and evals <= opts.ls_maxiter:
            alpha *= opts.contraction_factor
            newx = man.retr(x, alpha * desc)
            newf = objective(newx)
            evals += 1
        if newf > cost:
            alpha = 0.0
            newx = x
        stepsize = alpha * norm_d
        oldalpha = alpha if evals == 2 else 2.0 * alpha
        # new cost-related quantities
        newcost, newgrad = cost_grad(newx)
        newgradnorm = float(man.norm(newx, newgrad))
        newgPg = float(man.inner(newx, newgrad, newgrad))
        oldgrad = man.transp(x, newx, grad)
        desc = man.transp(x, newx, desc)
        diff = newgrad - oldgrad
        ip_diff = float(man.inner(newx, newgrad, diff))
        den = float(man.inner(newx, diff, desc))
        if den == 0.0 and gp.manifold == 'sphere':
            beta = 1.0   # Sphere.inner returns a Python float -> ZeroDivisionError branch of pymanopt
        else:
            with np.errstate(divide='ignore', invalid='ignore'):
                q = np.float64(ip_diff) / np.float64(den)
            beta = float(q) if q > 0 else 0.0   # max(0, q); NaN -> 0 as Python's max(0, nan)
        desc = -newgrad + beta * desc
        x, cost, grad, gradnorm, gradPgrad = newx, newcost, newgrad, newgradnorm, newgPg
        it += 1
    return x, cost, it, reason


def gen_candidates(gp, x0s, opts=None):
    """manifold_optimize.py:207-227: serial restarts, then acquisition at the candidates."""
    xs, vals, its = [], [], []
    for x0 in x0s:
        x, cost, it, _ = solve_cg(gp, x0, opts)
        xs.append(x)
        vals.append(_gp.ei_and_grad(gp, x, want_grad=False)[0])
        its.append(it)
    return np.array(xs), np.array(vals), np.array(its)


def best_candidate(values):
    """botorch ``get_best_candidates``: argmax of the batch values; first index on ties, NaN loses."""
    v = np.where(np.isnan(values), -np.inf, values)
    return int(np.argmax(v))


def lexi_argmax_records(values, indices):
    """Multi-GPU selection rule: highest value, then lowest global index; NaN = -inf."""
    v = np.where(np.isnan(values), -np.inf, np.asarray(values, dtype=np.float64))
    best = 0
    for i in range(1, len(v)):
        if v[i] > v[best] or (v[i] == v[best] and indices[i] < indices[best]):
            best = i
    return best
