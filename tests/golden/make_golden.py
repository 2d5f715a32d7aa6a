"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN CODE.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

It imports the importable slices of the reference (``oracle/reference_loader.py``: torch/numpy-only modules,
with a ``torch.symeig`` -> ``torch.linalg.eigh`` shim because torch 2.11 removed ``symeig``) and records
inputs + outputs of

* ``sphere_distance_torch`` (sphere_utils_torch.py:12-55) and ``exp(-beta d^2)`` as in kernels_sphere.py:89-94
* ``vector_to_symmetric_matrix_mandel_torch`` / ``symmetric_matrix_to_vector_mandel_torch`` (spd_utils_torch.py:159-226)
* ``affine_invariant_distance_torch`` (spd_utils_torch.py:53-120) and ``exp(-beta d^2)`` as in kernels_spd.py:96-98
* ``frobenius_distance_torch`` (spd_utils_torch.py:124-156), ``logm_torch`` (:13-30)
* ``projection_from_spd_to_nested_spd`` (nested_spd_utils.py:13-48)
* ``projection_from_sphere_to_subsphere`` (nested_spheres_utils.py:120-147), ``projection_from_sphere_to_nested_sphere``
  (:13-67), ``projection_from_subsphere_to_sphere`` (:149-213)
* ``sqrtm_torch`` (spd_utils_torch.py:33-50), ``projection_from_nested_spd_to_spd`` (nested_spd_utils.py:51-118)
* the numpy manifold formulas ``sphere_utils.py:14-123`` and ``spd_utils.py:104-213`` (exp/log/dist/transport)

The kernel classes themselves (``kernels_sphere.py`` / ``kernels_spd.py``) import gpytorch, which is not
installed, so ``K`` is formed here by the one line the class adds on top of the distance
(``torch.exp(-distance2.mul(beta.double()))``) with ``beta = beta_min + softplus(0) = beta_min + ln 2``.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_loader  # noqa: E402

SEED = 1234
LN2 = math.log(2.0)


def sphere_points(rng, n, dim):
    x = rng.standard_normal((n, dim))
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def spd_points(rng, n, d, max_cond=100.0):
    """Law of the reference's spd_sample (spd_utils.py:298-305) with the cond<=100 filter of
    examples/kernels/spd/spd_gaussian_kernel_parameters.py:91-96."""
    out = []
    while len(out) < n:
        lam = 0.001 + (5.0 - 0.001) * rng.random(d)
        q, _ = np.linalg.qr(rng.standard_normal((d, d)))
        if lam.max() / lam.min() > max_cond:
            continue
        m = (q * lam) @ q.T
        out.append(0.5 * (m + m.T))
    return np.array(out)


def main():
    ref = reference_loader.load()
    st = ref.sphere_utils_torch
    pt = ref.spd_utils_torch
    rng = np.random.default_rng(SEED)
    out = {}

    # ---- sphere -------------------------------------------------------------------------------
    for name, n, dim, beta_min in (('s2_n256', 256, 3, 6.5), ('s5_n1024', 1024, 6, 1.0), ('s8_n96', 96, 9, 0.6)):
        x = sphere_points(rng, n, dim)
        xt = torch.from_numpy(x)
        d = st.sphere_distance_torch(xt, xt)
        beta = beta_min + LN2
        k = torch.exp(-torch.mul(d, d).mul(torch.tensor(beta, dtype=torch.float64)))
        stride = max(1, n // 128)
        out[name + '_x'] = x
        out[name + '_beta'] = np.float64(beta)
        out[name + '_stride'] = np.int64(stride)
        out[name + '_d'] = d.numpy()[::stride, ::stride].copy()
        out[name + '_k'] = k.numpy()[::stride, ::stride].copy()
        out[name + '_ksum'] = np.float64(k.sum().item())
        out[name + '_ddiag'] = st.sphere_distance_torch(xt, xt, diag=True).numpy()
    # rectangular + near-duplicate / antipodal edge cases
    a = sphere_points(rng, 37, 4)
    b = sphere_points(rng, 53, 4)
    b[0] = a[0]
    b[1] = -a[1]
    tiny = 1e-4 * rng.standard_normal(4)
    b[2] = (a[2] + tiny) / np.linalg.norm(a[2] + tiny)
    out['s3_rect_a'] = a
    out['s3_rect_b'] = b
    out['s3_rect_d'] = st.sphere_distance_torch(torch.from_numpy(a), torch.from_numpy(b)).numpy()

    # ---- Mandel + SPD -------------------------------------------------------------------------
    for name, n, d, beta_min in (('spd3_n128', 128, 3, 0.5), ('spd8_n64', 64, 8, 0.22), ('spd2_n40', 40, 2, 0.6),
                                 ('spd5_n48', 48, 5, 0.25)):
        m = spd_points(rng, n, d)
        mt = torch.from_numpy(m)
        v = pt.symmetric_matrix_to_vector_mandel_torch(mt)
        back = pt.vector_to_symmetric_matrix_mandel_torch(v)
        dist = pt.affine_invariant_distance_torch(back, back)
        beta = beta_min + LN2
        k = torch.exp(-torch.mul(dist, dist).mul(torch.tensor(beta, dtype=torch.float64)))
        out[name + '_mat'] = m
        out[name + '_vec'] = v.numpy()
        out[name + '_unpacked'] = back.numpy()
        out[name + '_beta'] = np.float64(beta)
        out[name + '_d'] = dist.numpy()
        out[name + '_k'] = k.numpy()
        out[name + '_frob'] = pt.frobenius_distance_torch(back, back).numpy()
        out[name + '_logm'] = torch.stack([pt.logm_torch(back[i]) for i in range(n)]).numpy()
    # rectangular SPD(3)
    a = spd_points(rng, 19, 3)
    b = spd_points(rng, 27, 3)
    out['spd3_rect_a'] = a
    out['spd3_rect_b'] = b
    out['spd3_rect_d'] = pt.affine_invariant_distance_torch(torch.from_numpy(a), torch.from_numpy(b)).numpy()

    # ---- nested projection --------------------------------------------------------------------
    for name, n, D, d in (('proj_20_5', 64, 20, 5), ('proj_5_2', 16, 5, 2)):
        x = spd_points(rng, n, D, max_cond=1e9)
        w, _ = np.linalg.qr(rng.standard_normal((D, d)))
        y = ref.nested_spd_utils.projection_from_spd_to_nested_spd(torch.from_numpy(x), torch.from_numpy(w))
        out[name + '_x'] = x
        out[name + '_w'] = w
        out[name + '_y'] = y.numpy()
        out[name + '_xvec'] = pt.symmetric_matrix_to_vector_mandel_torch(torch.from_numpy(x)).numpy()
        out[name + '_yvec'] = pt.symmetric_matrix_to_vector_mandel_torch(y).numpy()

    # ---- numpy manifold formulas (sphere_utils.py / spd_utils.py) -------------------------------
    su, sp = ref.sphere_utils, ref.spd_utils
    n = 24
    xs = sphere_points(rng, n, 6)
    ys = sphere_points(rng, n, 6)
    logs = np.array([su.logmap(ys[i], xs[i])[:, 0] for i in range(n)])
    exps = np.array([su.expmap(0.7 * logs[i], xs[i])[:, 0] for i in range(n)])
    vs = rng.standard_normal((n, 6))
    vs = vs - np.sum(vs * xs, axis=-1, keepdims=True) * xs
    pts = np.array([su.parallel_transport_operator(xs[i], ys[i]) @ vs[i] for i in range(n)])
    out['man_sphere_x'], out['man_sphere_y'], out['man_sphere_v'] = xs, ys, vs
    out['man_sphere_log'], out['man_sphere_exp07'], out['man_sphere_pt'] = logs, exps, pts
    out['man_sphere_dist'] = np.array([su.sphere_distance(xs[i], ys[i]) for i in range(n)]).reshape(n)

    for d in (3, 8):
        xs = spd_points(rng, n, d)
        ys = spd_points(rng, n, d)
        logs = np.array([np.real(sp.logmap(ys[i], xs[i])) for i in range(n)])
        exps = np.array([np.real(sp.expmap(0.5 * logs[i], xs[i])) for i in range(n)])
        dist = np.array([np.real(sp.affine_invariant_distance(xs[i], ys[i])) for i in range(n)])
        us = rng.standard_normal((n, d, d))
        us = 0.5 * (us + np.swapaxes(us, -1, -2))
        pts = []
        for i in range(n):
            e = np.real(sp.parallel_transport_operator(xs[i], ys[i]))
            pts.append(e @ us[i] @ e.T)
        tag = 'man_spd%d_' % d
        out[tag + 'x'], out[tag + 'y'], out[tag + 'u'] = xs, ys, us
        out[tag + 'log'], out[tag + 'exp05'], out[tag + 'dist'], out[tag + 'pt'] = logs, exps, dist, np.array(pts)

    # ---- nested spheres (appended last so that the arrays above keep their values) ----------------------------
    nsu = ref.nested_spheres_utils
    for name, n, D, dl, r in (('nsph_5_3', 40, 5, 3, math.pi / 2), ('nsph_6_2', 33, 6, 2, 1.1)):
        x = sphere_points(rng, n, D)
        axes = [torch.from_numpy(sphere_points(rng, 1, k)) for k in range(D, dl, -1)]
        dists = [r * torch.ones(1, 1, dtype=torch.float64) for _ in axes]
        levels = nsu.projection_from_sphere_to_subsphere(torch.from_numpy(x), axes, dists)
        out[name + '_x'] = x
        out[name + '_r'] = np.array(r)
        for lvl, a in enumerate(axes):
            out[name + '_axis%d' % lvl] = a.numpy()
        for lvl, y in enumerate(levels[1:]):
            out[name + '_y%d' % lvl] = y.numpy()

    # ---- reconstruction maps (appended after everything above) ---------------------------------
    # nested spheres: projection onto the nested sphere (:13-67) and the inverse chain (:149-213), same axes as above
    for name, n, D, dl, r in (('nsph_5_3', 40, 5, 3, math.pi / 2), ('nsph_6_2', 33, 6, 2, 1.1)):
        x = torch.from_numpy(out[name + '_x'])
        axes = [torch.from_numpy(out[name + '_axis%d' % lvl]) for lvl in range(D - dl)]
        dists = [r * torch.ones(1, 1, dtype=torch.float64) for _ in axes]
        out[name + '_ns0'] = nsu.projection_from_sphere_to_nested_sphere(x, axes[0], dists[0]).numpy()
        y_low = torch.from_numpy(out[name + '_y%d' % (D - dl - 1)])
        ups = nsu.projection_from_subsphere_to_sphere(y_low, axes, dists)
        for lvl, u in enumerate(ups[1:]):
            out[name + '_up%d' % lvl] = u.numpy()

    # sqrtm_torch (spd_utils_torch.py:33-50)
    for d in (3, 5):
        xs = spd_points(rng, 20, d)
        out['sqrtm%d_x' % d] = xs
        out['sqrtm%d_y' % d] = np.array([pt.sqrtm_torch(torch.from_numpy(m)).numpy() for m in xs])

    # projection_from_nested_spd_to_spd (nested_spd_utils.py:51-118)
    for name, n, D, d in (('recon_5_2', 12, 5, 2), ('recon_20_5', 16, 20, 5)):
        q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        w, v = q[:, :d].copy(), q[:, d:].copy()
        c = spd_points(rng, 1, D - d, max_cond=1e9)[0]
        k = rng.standard_normal((d, D - d))
        k = 0.7 * k / np.linalg.norm(k, 2)
        y = spd_points(rng, n, d)
        x = ref.nested_spd_utils.projection_from_nested_spd_to_spd(
            torch.from_numpy(y), torch.from_numpy(w), torch.from_numpy(v), torch.from_numpy(c), torch.from_numpy(k))
        out[name + '_w'], out[name + '_v'], out[name + '_c'], out[name + '_k'] = w, v, c, k
        out[name + '_y'], out[name + '_x'] = y, x.numpy()

    # ---- trust-region acquisition solver: the reference's own TrustRegions class (robust_trust_regions.py) with its
    # finite-difference Hessian (approximate_hessian.py) on the oracle's EI problem; own generator so that the arrays
    # above keep their values ----------------------------------------------------------------------------------------
    import types
    from oracle import gp as ogp, rtr as ortr, sphere as osph, spd as ospd
    TrustRegions, get_hessianfd = reference_loader.load_trust_regions()
    rng_tr = np.random.default_rng(SEED + 1)
    for name, manifold, dim, n, beta, noise, nstart in (('rtr_s2', 'sphere', 3, 12, 6.5 + LN2, 1e-2, 12),
                                                        ('rtr_s5', 'sphere', 6, 32, 1.0 + LN2, 1e-2, 12),
                                                        ('rtr_s5_noisy', 'sphere', 6, 32, 1.0 + LN2, 2.0, 6)):
        xt = osph.rand(rng_tr, n, dim)
        y = osph.ackley(xt)
        gp = ogp.make_gp(manifold, xt, y, beta=beta, noise=noise)
        man = ortr._Man(manifold, xt[0])
        cost, grad = ortr.ei_problem(gp)

        class Problem(object):     # the attributes TrustRegions.solve reads from a pymanopt Problem
            manifold = man
            verbosity = 0

            def precon(self, x, d):           # manifold_optimize.py:190-193
                if np.sum(d) == 0.:
                    d += 1e-30
                return d
        problem = Problem()
        problem.cost, problem.grad = cost, grad
        problem.hess = types.MethodType(get_hessianfd, problem)   # manifold_optimize.py:199-200 (sets _hess there)
        x0 = osph.rand(rng_tr, nstart, dim)
        xs, fs, its = [], [], []
        for i in range(nstart):
            solver = TrustRegions()
            x = solver.solve(problem, x=x0[i].copy())
            xs.append(x)
            fs.append(cost(x))
            its.append(solver._last_iter)
        out[name + '_xtrain'], out[name + '_y'] = xt, np.asarray(y)
        out[name + '_hyper'] = np.array([beta, noise])
        out[name + '_x0'], out[name + '_x'] = x0, np.array(xs)
        out[name + '_cost'], out[name + '_iters'] = np.array(fs), np.array(its)

    # ---- constrained trust regions: the reference's own ConstrainedTrustRegions class with its own Problem / PyTorch
    # autodiff backend for the constraint, in the configuration of gabo_spd.py (mingradnorm 1e-4, maxiter 100,
    # approx_hessian, one max-eigenvalue inequality constraint); own generator again -------------------------------
    import functools
    CTR, get_hessianfd_c, cons = reference_loader.load_constrained_trust_regions()
    rng_c = np.random.default_rng(SEED + 2)
    for name, d, n, max_eig, nstart in (('ctr_spd2', 2, 10, 3.0, 8), ('ctr_spd3', 3, 16, 4.0, 6),
                                            ('ctr_spd2_active', 2, 10, 2.0, 8)):
        xt = ospd.spd_sample(rng_c, n, d, max_cond=50.0)
        y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt)))
        beta = 0.5 + LN2
        gp = ogp.make_gp('spd', xt, y, beta=beta, noise=1e-2)
        man = ortr._Man('spd', xt[0])
        man.egrad2rgrad = ospd.egrad2rgrad          # used by the reference's Problem.grad for the constraint
        cost, grad = ortr.ei_problem(gp)

        class ProblemC(object):
            manifold = man
            verbosity = 0

            def precon(self, x, dd):
                if np.sum(dd) == 0.:
                    dd += 1e-30
                return dd
        problem = ProblemC()
        problem.cost, problem.grad = cost, grad
        problem.hess = types.MethodType(get_hessianfd_c, problem)
        constraint = functools.partial(cons.max_eigenvalue_constraint_torch, maximum_eigenvalue=max_eig)
        x0 = ospd.spd_sample(rng_c, nstart, d, min_eig=0.5, max_eig=0.9 * max_eig)      # feasible starts
        xs, fs, its = [], [], []
        for i in range(nstart):
            solver = CTR(mingradnorm=1e-4, maxiter=100)
            x = solver.solve(problem, x=x0[i].copy(), ineq_constraints=[constraint])
            xs.append(x)
            fs.append(cost(x))
            its.append(solver._last_iter)
        out[name + '_xtrain'], out[name + '_y'] = xt, np.asarray(y)
        out[name + '_hyper'] = np.array([beta, 1e-2, max_eig])
        out[name + '_x0'], out[name + '_x'] = x0, np.array(xs)
        out[name + '_cost'], out[name + '_iters'] = np.array(fs), np.array(its)

    # ---- strict variant: the reference's StrictConstrainedTrustRegions with the settings of hd_gabo_spd.py
    # (mingradnorm 2e-4, maxiter 100); own generator again ----------------------------------------------------------
    rng_s = np.random.default_rng(SEED + 3)
    for name, d, n, max_eig, nstart in (('sctr_spd2_active', 2, 10, 2.0, 8), ('sctr_spd3', 3, 16, 2.5, 6)):
        xt = ospd.spd_sample(rng_s, n, d, max_cond=50.0)
        y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt)))
        beta = 0.5 + LN2
        gp = ogp.make_gp('spd', xt, y, beta=beta, noise=1e-2)
        man = ortr._Man('spd', xt[0])
        man.egrad2rgrad = ospd.egrad2rgrad
        cost, grad = ortr.ei_problem(gp)

        class ProblemS(object):
            manifold = man
            verbosity = 0

            def precon(self, x, dd):
                if np.sum(dd) == 0.:
                    dd += 1e-30
                return dd
        problem = ProblemS()
        problem.cost, problem.grad = cost, grad
        problem.hess = types.MethodType(get_hessianfd_c, problem)
        constraint = functools.partial(cons.max_eigenvalue_constraint_torch, maximum_eigenvalue=max_eig)
        x0 = ospd.spd_sample(rng_s, nstart, d, min_eig=0.5, max_eig=0.9 * max_eig)
        xs, fs, its = [], [], []
        for i in range(nstart):
            solver = CTR.Strict(mingradnorm=2e-4, maxiter=100, minstepsize=1e-4)
            x = solver.solve(problem, x=x0[i].copy(), ineq_constraints=[constraint])
            xs.append(x)
            fs.append(cost(x))
            its.append(solver._last_iter)
        out[name + '_xtrain'], out[name + '_y'] = xt, np.asarray(y)
        out[name + '_hyper'] = np.array([beta, 1e-2, max_eig])
        out[name + '_x0'], out[name + '_x'] = x0, np.array(xs)
        out[name + '_cost'], out[name + '_iters'] = np.array(fs), np.array(its)

    # ---- ConstrainedTrustRegions on the sphere with the domain constraint of
    # examples/bo_sphere/constrained_benchmark_examples/gabo_sphere_inequality_constraints.py:109-122 (points within
    # pi/4 of a centre; solver ConstrainedTrustRegions(maxiter=200), :249) -------------------------------------------
    rng_d = np.random.default_rng(SEED + 4)
    dim = 3
    centre = np.zeros(dim)
    centre[0] = 1.0
    for name, angle in (('ctr_s2_domain', np.pi / 4.), ('ctr_s2_domain_active', 0.12)):   # pi/4: the example's value
        def domain_constraint(x, angle=angle):
            c = torch.Tensor(centre).type(x.dtype)
            in_prod = torch.mm(x[None], c[:, None])
            in_prod = torch.max(torch.min(in_prod, torch.ones(1, dtype=x.dtype)), -torch.ones(1, dtype=x.dtype))
            return angle - torch.acos(in_prod)[0, 0]
        xt = osph.rand(rng_d, 14, dim)
        y = osph.ackley(xt)
        beta = 6.5 + LN2
        gp = ogp.make_gp('sphere', xt, y, beta=beta, noise=1e-2)
        man = ortr._Man('sphere', xt[0])
        man.egrad2rgrad = osph.proj
        cost, grad = ortr.ei_problem(gp)

        class ProblemD(object):
            manifold = man
            verbosity = 0

            def precon(self, x, dd):
                if np.sum(dd) == 0.:
                    dd += 1e-30
                return dd
        problem = ProblemD()
        problem.cost, problem.grad = cost, grad
        problem.hess = types.MethodType(get_hessianfd_c, problem)
        x0 = []
        while len(x0) < 10:                               # starts inside the domain (sample_sphere_constrained, :125)
            p = osph.rand(rng_d, 1, dim)[0]
            if np.arccos(np.clip(p @ centre, -1, 1)) < 0.9 * angle:
                x0.append(p)
        x0 = np.array(x0)
        xs, fs, its = [], [], []
        for i in range(len(x0)):
            solver = CTR(maxiter=200)
            x = solver.solve(problem, x=x0[i].copy(), ineq_constraints=[domain_constraint])
            xs.append(x)
            fs.append(cost(x))
            its.append(solver._last_iter)
        out[name + '_xtrain'], out[name + '_y'] = xt, np.asarray(y)
        out[name + '_hyper'] = np.array([beta, 1e-2, angle])
        out[name + '_x0'], out[name + '_x'] = x0, np.array(xs)
        out[name + '_cost'], out[name + '_iters'] = np.array(fs), np.array(its)

    # ---- equality constraint: the great circle x[1] = 0 of gabo_sphere_equality_constraints.py:104-109 with
    # ConstrainedTrustRegions(maxiter=200) (:197), starts on the circle (sample_sphere_constrained, :112-118) ----------
    rng_e = np.random.default_rng(SEED + 5)

    def y_great_circle(x):
        return x[1] - 0.
    xt = osph.rand(rng_e, 14, 3)
    y = osph.ackley(xt)
    beta = 6.5 + LN2
    gp = ogp.make_gp('sphere', xt, y, beta=beta, noise=1e-2)
    man = ortr._Man('sphere', xt[0])
    man.egrad2rgrad = osph.proj
    cost, grad = ortr.ei_problem(gp)

    class ProblemE(object):
        manifold = man
        verbosity = 0

        def precon(self, x, dd):
            if np.sum(dd) == 0.:
                dd += 1e-30
            return dd
    problem = ProblemE()
    problem.cost, problem.grad = cost, grad
    problem.hess = types.MethodType(get_hessianfd_c, problem)
    x0 = rng_e.standard_normal((8, 3))
    x0[:, 1] = 0.
    x0 /= np.linalg.norm(x0, axis=-1, keepdims=True)
    xs, fs, its = [], [], []
    for i in range(len(x0)):
        solver = CTR(maxiter=200)
        x = solver.solve(problem, x=x0[i].copy(), eq_constraints=[y_great_circle])
        xs.append(x)
        fs.append(cost(x))
        its.append(solver._last_iter)
    out['ctr_s2_circle_xtrain'], out['ctr_s2_circle_y'] = xt, np.asarray(y)
    out['ctr_s2_circle_hyper'] = np.array([beta, 1e-2])
    out['ctr_s2_circle_x0'], out['ctr_s2_circle_x'] = x0, np.array(xs)
    out['ctr_s2_circle_cost'], out['ctr_s2_circle_iters'] = np.array(fs), np.array(its)

    # ---- augmented Lagrangian: the reference's own AugmentedLagrangeMethod around its own TrustRegions, as in
    # gabo_sphere_inequality_constraints.py:251-256 / gabo_sphere_equality_constraints.py:199-203 (gammas_fact=0.05);
    # iteration limits reduced (30 outer, 50 inner) to keep the fixture generation short ---------------------------
    from oracle import alm as oalm  # noqa: F401
    ALM = reference_loader.load_alm()
    rng_a = np.random.default_rng(SEED + 6)
    for name, kind in (('alm_s2_domain', 'ineq'), ('alm_s2_circle', 'eq')):
        angle = 0.12

        def domain_constraint(x, angle=angle):
            c = torch.Tensor(centre).type(x.dtype)
            in_prod = torch.mm(x[None], c[:, None])
            in_prod = torch.max(torch.min(in_prod, torch.ones(1, dtype=x.dtype)), -torch.ones(1, dtype=x.dtype))
            return angle - torch.acos(in_prod)[0, 0]

        def great_circle(x):
            return x[1] - 0.
        xt = osph.rand(rng_a, 14, 3)
        y = osph.ackley(xt)
        beta = 6.5 + LN2
        gp = ogp.make_gp('sphere', xt, y, beta=beta, noise=1e-2)
        man = ortr._Man('sphere', xt[0])
        man.egrad2rgrad = osph.proj
        cost, grad = ortr.ei_problem(gp)

        class ProblemA(object):
            manifold = man
            verbosity = 0

            def precon(self, x, dd):
                if np.sum(dd) == 0.:
                    dd += 1e-30
                return dd
        problem = ProblemA()
        problem.cost, problem.grad = cost, grad
        if kind == 'ineq':
            x0 = []
            while len(x0) < 6:
                p = osph.rand(rng_a, 1, 3)[0]
                if np.arccos(np.clip(p @ centre, -1, 1)) < 0.9 * angle:
                    x0.append(p)
            x0 = np.array(x0)
        else:
            x0 = rng_a.standard_normal((6, 3))
            x0[:, 1] = 0.
            x0 /= np.linalg.norm(x0, axis=-1, keepdims=True)
        xs, its = [], []
        for i in range(len(x0)):
            solver = ALM(inner_solver=TrustRegions(maxiter=50), maxiter=30, gammas_fact=0.05)
            if kind == 'ineq':
                x = solver.solve(problem, x=x0[i].copy(), ineq_constraints=[domain_constraint])
            else:
                x = solver.solve(problem, x=x0[i].copy(), eq_constraints=[great_circle])
            xs.append(x)
            its.append(solver._last_iter)
        out[name + '_xtrain'], out[name + '_y'] = xt, np.asarray(y)
        out[name + '_hyper'] = np.array([beta, 1e-2, angle])
        out[name + '_x0'], out[name + '_x'], out[name + '_iters'] = x0, np.array(xs), np.array(its)

    path = os.path.join(HERE, 'reference_vectors.npz')
    np.savez_compressed(path, **out)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
