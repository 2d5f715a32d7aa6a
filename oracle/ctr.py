"""Oracle (CPU, fp64) for the constrained trust-region acquisition solver.  Test infrastructure only.

Restates the reference's OWN ``ConstrainedTrustRegions`` and ``StrictConstrainedTrustRegions``
(``BoManifolds/manifold_optimization/constrained_trust_regions.py``:
``solve`` ``:120-439`` -- the radius update additionally grows on ``REACHED_CONSTRAINTS`` ``:321-323``;
``_constrained_truncated_conjugate_gradient`` ``:441-735`` -- the linearised constraints
``c(x) + <grad c, eta>`` are kept within ``Delta_cons`` (1e-6) of feasibility along the tCG path, inequality terms
counting only when negative) with the finite-difference Hessian of ``approximate_hessian.py`` -- the configuration
``examples/bo_spd/benchmark_examples/gabo_spd.py:98-203`` runs: ``ConstrainedTrustRegions(mingradnorm=1e-4,
maxiter=100)``, ``approx_hessian=True``, one inequality constraint ``max_eigenvalue_constraint_torch``
(``Riemannian_utils/spd_constraints_utils_torch.py:17-32``).

PINNED on the reference's code: ``tests/golden/make_golden.py`` runs the reference's class itself, with the
reference's own ``pymanopt_addons.problem.Problem`` + PyTorch autodiff backend for the constraint (value and gradient
through ``torch.symeig``), on the oracle's EI problem (``ctr_*`` arrays); ``tests/test_oracle_golden.py`` compares.
The product side is ``gabotorch_b200.manifold_optimization.batched_trust_regions(..., ineq_constraints=...)``.
"""
import numpy as np

from . import rtr as _rtr

(NEGATIVE_CURVATURE, EXCEEDED_TR, REACHED_TARGET_LINEAR, REACHED_TARGET_SUPERLINEAR, MAX_INNER_ITER, MODEL_INCREASED,
 REACHED_CONSTRAINTS) = range(7)


def _sym(a):
    return 0.5 * (a + np.swapaxes(a, -1, -2))


def max_eigenvalue_constraint(maximum_eigenvalue):
    """(value, Riemannian gradient) of ``maximum_eigenvalue - lambda_max(X)`` on SPD(d) under the affine-invariant
    metric: Euclidean gradient ``-v v^T`` (what autograd returns through symeig), ``egrad2rgrad = X sym(G) X``."""
    def value(x):
        return maximum_eigenvalue - np.linalg.eigvalsh(x)[-1]

    def grad(x):
        _, vec = np.linalg.eigh(x)
        v = vec[:, -1]
        return x @ _sym(-np.outer(v, v)) @ x
    return value, grad


def min_eigenvalue_constraint(minimum_eigenvalue):
    """``lambda_min(X) - minimum_eigenvalue`` (spd_constraints_utils_torch.py:35-50)."""
    def value(x):
        return np.linalg.eigvalsh(x)[0] - minimum_eigenvalue

    def grad(x):
        _, vec = np.linalg.eigh(x)
        v = vec[:, 0]
        return x @ _sym(np.outer(v, v)) @ x
    return value, grad


def sphere_domain_constraint(centre, angle):
    """``angle - acos(clamp(<x, centre>))`` on the sphere
    (examples/bo_sphere/constrained_benchmark_examples/gabo_sphere_inequality_constraints.py:113-120); Riemannian
    gradient = projection of the Euclidean one, ``centre / sqrt(1 - <x, centre>^2)`` (zero where the clamp is active)."""
    centre = np.asarray(centre, dtype=np.float64)

    def value(x):
        return angle - np.arccos(np.clip(x @ centre, -1.0, 1.0))

    def grad(x):
        u = x @ centre
        e = centre / np.sqrt(1.0 - u * u) if -1.0 < u < 1.0 else np.zeros_like(centre)
        return e - (x @ e) * x
    return value, grad


def _tau_constraints(fc, pe, pd, idx, delta_cons):
    """Largest step along delta keeping |fc + pe + tau pd|_idx <= delta_cons (the quadratic of :578-592, :636-650)."""
    qa = np.inner(pd[idx], pd[idx])
    qb = 2.0 * (np.inner(fc[idx], pd[idx]) + np.inner(pe[idx], pd[idx]))
    qc = np.inner(fc[idx], fc[idx]) + 2.0 * np.inner(fc[idx], pe[idx]) + np.inner(pe[idx], pe[idx]) - delta_cons ** 2
    disc = qb * qb - 4.0 * qa * qc
    with np.errstate(divide='ignore', invalid='ignore'):
        return (-qb + np.sqrt(disc)) / (2.0 * qa) if disc >= 0.0 else 0.0


def constrained_truncated_cg(man, hess, x, fgradx, radius, theta, kappa, mininner, maxinner, f_eq, g_eq, f_ineq, g_ineq,
                             delta_cons):
    """constrained_trust_regions.py:441-735 (use_rand=False, identity preconditioner)."""
    inner = man.inner
    eta = np.zeros_like(fgradx)
    heta = np.zeros_like(fgradx)
    r = fgradx
    e_pe = 0.0
    r_r = inner(x, r, r)
    norm_r0 = np.sqrt(r_r)
    z_r = r_r
    d_pd = z_r
    delta = -r
    e_pd = 0.0
    model_value = 0.0
    neq, nineq = len(f_eq), len(f_ineq)
    fc = np.array(list(f_eq) + list(f_ineq), dtype=np.float64)
    grads = list(g_eq) + list(g_ineq)
    pe = np.array([inner(x, gc, eta) for gc in grads], dtype=np.float64)
    stop = MAX_INNER_ITER
    alpha = None
    j = 0

    def violated(step):
        term = fc + pe + step * pd
        term[neq:] = np.minimum(0.0, term[neq:])
        # as in the reference (:573-577): np.where over the WHOLE vector, shifted by the number of equality constraints
        # (identical to indexing the inequality block when there are no equality constraints)
        idx = np.hstack((np.arange(neq, dtype=int), np.where(term < 0)[0] + neq)) if nineq > 0 \
            else np.arange(neq, dtype=int)
        return np.inner(term, term) > delta_cons ** 2, idx

    for j in range(int(maxinner)):
        hdelta = hess(x, delta)
        d_hd = inner(x, delta, hdelta)
        if d_hd != 0:
            alpha = z_r / d_hd
            e_pe_new = e_pe + 2 * alpha * e_pd + alpha ** 2 * d_pd
        else:
            e_pe_new = e_pe
        pd = np.array([inner(x, gc, delta) for gc in grads], dtype=np.float64)
        if d_hd <= 0 or e_pe_new >= radius ** 2:
            with np.errstate(divide='ignore', invalid='ignore'):
                tau_tr = (-e_pd + np.sqrt(e_pd * e_pd + d_pd * (radius ** 2 - e_pe))) / d_pd
            if np.isnan(tau_tr):
                tau_tr = 0.0
            bad, idx = violated(tau_tr)
            tau = _tau_constraints(fc, pe, pd, idx, delta_cons) if bad else tau_tr
            eta = eta + tau * delta
            heta = heta + tau * hdelta
            stop = NEGATIVE_CURVATURE if d_hd <= 0 else (REACHED_CONSTRAINTS if bad else EXCEEDED_TR)
            break
        bad, idx = violated(alpha)
        if bad:
            tau = _tau_constraints(fc, pe, pd, idx, delta_cons)
            eta = eta + tau * delta
            heta = heta + tau * hdelta
            stop = REACHED_CONSTRAINTS
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
        zold_rold = z_r
        z_r = r_r
        beta = z_r / zold_rold
        delta = -r + beta * delta
        e_pd = beta * (e_pd + alpha * d_pd)
        d_pd = z_r + beta * beta * d_pd
        pe = pe + alpha * pd
    return eta, heta, j, stop


def solve_ctr(gp, x0, eq_constraints=(), ineq_constraints=(), opts=None, delta_cons=1e-6, trace=None, strict=False):
    """One ``ConstrainedTrustRegions.solve`` on cost = -EI.  Constraints are (value, Riemannian gradient) pairs.
    ``strict=True`` is ``StrictConstrainedTrustRegions`` (:737-1415): the only difference is that a proposal violating
    a constraint gets an infinite cost and shrinks the radius (:936-952, :972).  Returns (x, cost, iters)."""
    opts = opts or _rtr.TROptions()
    x = np.array(x0, dtype=np.float64)
    man = _rtr._Man(gp.manifold, x)
    cost, grad = _rtr.ei_problem(gp)
    maxinner = man.dim if opts.maxinner is None else opts.maxinner
    delta_bar = man.typicaldist if opts.delta_bar is None else opts.delta_bar
    delta0 = delta_bar / 8 if opts.delta0 is None else opts.delta0

    def hess(p, a):
        return _rtr.hessian_fd(man, grad, p, a, opts.fd_epsilon)

    k = 0
    fx = cost(x)
    fgradx = grad(x)
    norm_grad = man.norm(x, fgradx)
    radius = delta0
    while True:
        if trace is not None:
            trace.append((k, x.copy(), fx, norm_grad, radius))
        f_eq = [c[0](x) for c in eq_constraints]
        g_eq = [c[1](x) for c in eq_constraints]
        f_in = [c[0](x) for c in ineq_constraints]
        g_in = [c[1](x) for c in ineq_constraints]
        eta, heta, _, stop_inner = constrained_truncated_cg(man, hess, x, fgradx, radius, opts.theta, opts.kappa,
                                                            opts.mininner, maxinner, f_eq, g_eq, f_in, g_in, delta_cons)
        x_prop = man.retr(x, eta)
        invalid_prop = False
        if strict:
            fcp = [c[0](x_prop) for c in eq_constraints] + [min(c[0](x_prop), 0.0) for c in ineq_constraints]
            invalid_prop = float(np.sum(np.abs(np.array(fcp, dtype=np.float64)))) != 0.0
        fx_prop = np.inf if invalid_prop else cost(x_prop)
        rho_reg = max(1, abs(fx)) * np.spacing(1) * opts.rho_regularization
        rhonum = fx - fx_prop + rho_reg
        rhoden = -man.inner(x, fgradx, eta) - 0.5 * man.inner(x, eta, heta) + rho_reg
        model_decreased = rhoden >= 0
        with np.errstate(divide='ignore', invalid='ignore'):
            rho = np.float64(rhonum) / np.float64(rhoden)
        if rho < 0.25 or not model_decreased or np.isnan(rho) or invalid_prop:
            radius = radius / 4
        elif rho > 0.75 and stop_inner in (NEGATIVE_CURVATURE, EXCEEDED_TR, REACHED_CONSTRAINTS):
            radius = min(2 * radius, delta_bar)
        if model_decreased and rho > opts.rho_prime:
            x = x_prop
            fx = fx_prop
            fgradx = grad(x)
            norm_grad = man.norm(x, fgradx)
        k += 1
        if k >= opts.maxiter or norm_grad < opts.mingradnorm:
            break
    return x, fx, k
