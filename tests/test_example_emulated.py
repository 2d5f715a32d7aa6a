"""The BO loop of examples/gabo_sphere.py (the reference's gabo_sphere.py written against the drop-in modules) run on
CPU with every CUDA entry point replaced by an oracle-backed stand-in: pins the HOST wiring of a whole iteration -- GP fit
(L-BFGS over the marginal likelihood), Expected Improvement, raw-sample screening, multi-start trust regions, candidate
selection, data update -- without a device.  The kernels themselves are covered by the -m gpu tests."""
import importlib.util
import os

import numpy as np
import torch

from gabotorch_b200 import _lib, ops
from oracle import gp as ogp
from oracle import rtr as ortr
from oracle import spd as ospd
from oracle import sphere as osph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_gp(dgp):
    d = dgp.desc
    if dgp.manifold == _lib.SPD:      # the emulated "factor records" are the flattened matrices themselves
        xt = dgp.x_train.numpy().reshape(-1, dgp.dim, dgp.dim)
        return ogp.GPData('spd', xt, None, d.mean, d.outputscale, 0.0, d.beta, d.best_f, alpha=dgp.alpha.numpy(),
                          minv=dgp.minv.numpy(), kxx=d.kxx)
    return ogp.GPData('sphere', dgp.x_train.numpy(), None, d.mean, d.outputscale, 0.0, d.beta, d.best_f,
                      alpha=dgp.alpha.numpy(), minv=dgp.minv.numpy(), kxx=d.kxx)


def _install(monkeypatch):
    t64 = lambda x: torch.as_tensor(x, dtype=torch.float64)  # noqa: E731
    monkeypatch.setattr(ops, 'device', lambda: torch.device('cpu'))
    monkeypatch.setattr(ops, 'to_dev64', lambda x: t64(x).detach().contiguous())

    def sphere_gram(x1, x2, param=0.0, kind=_lib.KIND_GAUSS, diag=False, out_dtype=torch.float64):
        d = osph.sphere_distance(t64(x1), t64(x2), diag=diag)
        if kind == _lib.KIND_DIST:
            return d
        return torch.exp(-param * d * d) if kind == _lib.KIND_GAUSS else torch.exp(-param * d)
    monkeypatch.setattr(ops, 'sphere_gram', sphere_gram)

    def gp_mll(dmat, y, theta, want_grad=True, want_factors=False):
        theta = t64(theta).reshape(-1, 4).numpy()
        ll, gr = zip(*[ogp.exact_log_likelihood(t64(dmat).numpy(), t64(y).numpy(), th) for th in theta])
        return (torch.tensor(ll, dtype=torch.float64), torch.tensor(np.array(gr)) if want_grad else None, None, None,
                torch.zeros(len(theta), dtype=torch.int32))
    monkeypatch.setattr(ops, 'gp_mll', gp_mll)

    def gp_fit(dmat, y, raw0, beta_min, noise_min, priors, fixed, maxiter=15000, pgtol=1e-5, ftol=2.220446049250313e-09):
        # stand-in of the one-launch BFGS fit: L-BFGS-B over the same objective (the emulated gp_mll above), per start
        from scipy.optimize import minimize
        from gabotorch_b200.gp_fit import MarginalLogLikelihood
        pri = [(priors[2 * i], priors[2 * i + 1]) if priors[2 * i] > 0 else None for i in range(3)]
        obj = MarginalLogLikelihood(dmat, y, beta_min, noise_min, beta_prior=pri[0], outputscale_prior=pri[1],
                                    noise_prior=pri[2])
        raws, fs, info = [], [], []
        for start in t64(raw0).reshape(-1, 4).numpy():
            bounds = [(v, v) if f else (None, None) for v, f in zip(start, fixed)]
            res = minimize(obj, start, jac=True, method='L-BFGS-B', bounds=bounds,
                           options={'maxiter': maxiter, 'gtol': pgtol, 'ftol': ftol})
            raws.append(res.x), fs.append(res.fun), info.append([0 if res.success else 2, res.nit, res.nfev])
        return np.array(raws), np.array(fs), np.array(info, dtype=int)
    monkeypatch.setattr(ops, 'gp_fit', gp_fit)

    def gp_factor(kmat, y, outputscale, noise, mean):
        k = outputscale * t64(kmat).numpy() + noise * np.eye(len(y))
        k = np.tril(k) + np.tril(k, -1).T
        kin = np.linalg.inv(k)
        return torch.from_numpy(kin @ (t64(y).numpy() - mean)), torch.from_numpy(0.5 * (kin + kin.T))
    monkeypatch.setattr(ops, 'gp_factor', gp_factor)

    def ei_eval(gp, x, want_grad=False):
        og = _oracle_gp(gp)
        out = []
        for p in t64(x).numpy():
            try:
                out.append(ogp.ei_and_grad(og, p, want_grad=True))
            except np.linalg.LinAlgError:                       # not SPD: NaN, as the device entry reports it
                out.append((np.nan, np.full(p.shape, np.nan)))
        ei, gr = torch.tensor([o[0] for o in out], dtype=torch.float64), torch.tensor(np.array([o[1] for o in out]))
        return (ei, gr) if want_grad else ei
    monkeypatch.setattr(ops, 'ei_eval', ei_eval)

    def acq_rtr(gp, x0, maxiter=1000, mingradnorm=1e-6, kappa=0.1, theta=1.0, rho_prime=0.1, rho_regularization=1e3,
                mininner=1, maxinner=None, delta_bar=None, delta0=None):
        og = _oracle_gp(gp)
        opts = ortr.TROptions(maxiter=maxiter, mingradnorm=mingradnorm, kappa=kappa, theta=theta, rho_prime=rho_prime,
                              rho_regularization=rho_regularization, mininner=mininner, maxinner=maxinner,
                              delta_bar=delta_bar, delta0=delta0)
        xs, vals, its = ortr.gen_candidates(og, t64(x0).numpy(), opts)
        return (torch.from_numpy(xs), torch.from_numpy(vals), torch.from_numpy(its.astype(np.int32)),
                torch.full((len(its),), 2, dtype=torch.int32))
    monkeypatch.setattr(ops, 'acq_rtr', acq_rtr)

    def argmax_records(values, gidx=None):
        v = torch.nan_to_num(t64(values).reshape(-1), nan=-float('inf'))
        slot = torch.argmax(v).reshape(1)
        return slot, v[slot]
    monkeypatch.setattr(ops, 'argmax_records', argmax_records)

    def sphere_op(op, a, b, c=None):
        assert op == _lib.OP_LOG
        a, b = np.atleast_2d(t64(a).numpy()), np.atleast_2d(t64(b).numpy())
        return torch.from_numpy(osph.log(np.broadcast_to(a, b.shape), b))
    monkeypatch.setattr(ops, 'sphere_op', sphere_op)


def test_gabo_sphere_example_loop_with_emulated_kernels(monkeypatch):
    _install(monkeypatch)
    spec = importlib.util.spec_from_file_location('gabo_sphere_example', os.path.join(ROOT, 'examples', 'gabo_sphere.py'))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    x, y, best = ex.run(dim=3, n_iters=3, num_restarts=3, raw_samples=24, nb_data_init=5, seed=7, verbose=False)
    assert tuple(x.shape) == (8, 3) and tuple(y.shape) == (8,) and len(best) == 4
    np.testing.assert_allclose(x.norm(dim=-1).numpy(), 1.0, atol=1e-12)                 # candidates stay on the sphere
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and best[-1] == float(y.min())
    # the objective of the example equals the oracle's restatement of test_functions_sphere.py:34-65
    np.testing.assert_allclose(y.numpy(), osph.ackley(x.numpy()), rtol=1e-12)


def _install_spd(monkeypatch):
    t64 = lambda x: torch.as_tensor(x, dtype=torch.float64)  # noqa: E731
    unpack = lambda v: ospd.vector_to_symmetric_matrix_mandel(t64(v))  # noqa: E731
    monkeypatch.setattr(ops, 'mandel_unpack', unpack)
    monkeypatch.setattr(ops, 'mandel_pack', lambda m: ospd.symmetric_matrix_to_vector_mandel(t64(m)))

    def spd_factor(x, d, is_mandel, check=True, flags=None):
        m = unpack(x) if is_mandel else t64(x)
        if float(torch.linalg.eigvalsh(m).min()) <= 0:
            raise ops.NotPositiveDefiniteError('not positive definite')
        return m.reshape(m.shape[0], d * d).contiguous()
    monkeypatch.setattr(ops, 'spd_factor', spd_factor)

    def from_factors(fac1, fac2, d, param=0.0, kind=_lib.KIND_GAUSS, compute=_lib.GABO_F32, symmetric=False,
                     out_dtype=torch.float64, out=None):
        dist = ospd.affine_invariant_distance(fac1.reshape(-1, d, d), fac2.reshape(-1, d, d), exact=True)
        if kind == _lib.KIND_DIST:
            return dist
        return torch.exp(-param * dist * dist) if kind == _lib.KIND_GAUSS else torch.exp(-param * dist)
    monkeypatch.setattr(ops, 'spd_ai_gram_from_factors', from_factors)

    def spd_ai_gram(x1, x2, param=0.0, kind=_lib.KIND_GAUSS, is_mandel=True, compute=_lib.GABO_F32, **kw):
        m1, m2 = (unpack(x1), unpack(x2)) if is_mandel else (t64(x1), t64(x2))
        d = m1.shape[-1]
        return from_factors(m1.reshape(-1, d * d), m2.reshape(-1, d * d), d, param, kind)
    monkeypatch.setattr(ops, 'spd_ai_gram', spd_ai_gram)

    def spd_scalar(what, x, b, c=None):
        x, b = t64(x).numpy(), t64(b).numpy()
        if what == 0:
            return torch.tensor(np.array([ospd.dist(p, q) for p, q in zip(x, b)]))
        if what == 1:
            return torch.tensor(np.array([ospd.norm(p, u) for p, u in zip(x, b)]))
        return torch.tensor(np.array([ospd.inner(p, u, v) for p, u, v in zip(x, b, t64(c).numpy())]))
    monkeypatch.setattr(ops, 'spd_scalar', spd_scalar)

    def spd_op(op, a, b, c=None):
        fn = {_lib.OP_RETR: ospd.retr, _lib.OP_EXP: ospd.exp, _lib.OP_EGRAD2RGRAD: ospd.egrad2rgrad,
              _lib.OP_LOG: ospd.log}[op]
        a, b = t64(a).numpy(), t64(b).numpy()
        shape = b.shape
        d = shape[-1]
        a, b = np.broadcast_to(a, shape).reshape(-1, d, d), b.reshape(-1, d, d)
        return torch.from_numpy(np.array([fn(p, u) for p, u in zip(a, b)]).reshape(shape))
    monkeypatch.setattr(ops, 'spd_op', spd_op)


    def acq_ctr(gp, x0, constraints=(), strict=False, delta_cons=1e-6, maxiter=1000, mingradnorm=1e-6, kappa=0.1,
                theta=1.0, rho_prime=0.1, rho_regularization=1e3, mininner=1, maxinner=None, delta_bar=None, delta0=None):
        # stand-in of gabo_acq_ctr: the oracle restatement of the reference's [Strict]ConstrainedTrustRegions
        from oracle import ctr as octr
        og = _oracle_gp(gp)
        opts = ortr.TROptions(maxiter=maxiter, mingradnorm=mingradnorm, kappa=kappa, theta=theta, rho_prime=rho_prime,
                              rho_regularization=rho_regularization, mininner=mininner, maxinner=maxinner,
                              delta_bar=delta_bar, delta0=delta0)
        cons = [octr.max_eigenvalue_constraint(b) if k == 'max' else octr.min_eigenvalue_constraint(b)
                for k, b in constraints]
        xs, vals, its = [], [], []
        for p in t64(x0).numpy():
            x, c, k = octr.solve_ctr(og, p, ineq_constraints=cons, opts=opts, delta_cons=delta_cons, strict=strict)
            xs.append(x)
            vals.append(-c)
            its.append(k)
        return (torch.from_numpy(np.array(xs)), torch.tensor(vals, dtype=torch.float64),
                torch.tensor(its, dtype=torch.int32), torch.full((len(its),), 2, dtype=torch.int32))
    monkeypatch.setattr(ops, 'acq_ctr', acq_ctr)


def test_gabo_spd_example_loop_with_emulated_kernels(monkeypatch):
    # gabo_spd.py's loop: Mandel inputs, ConstrainedTrustRegions with the max-eigenvalue constraint (routed to the
    # one-launch kernel entry gabo_acq_ctr, emulated here by the oracle), GP fit on the affine-invariant squared distances
    _install(monkeypatch)
    _install_spd(monkeypatch)
    spec = importlib.util.spec_from_file_location('gabo_spd_example', os.path.join(ROOT, 'examples', 'gabo_spd.py'))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    x, y, best = ex.run(dim=2, n_iters=2, num_restarts=3, raw_samples=16, nb_data_init=5, seed=3, verbose=False)
    assert tuple(x.shape) == (7, 3) and tuple(y.shape) == (7,) and len(best) == 3
    mats = ospd.vector_to_symmetric_matrix_mandel(x).numpy()
    assert np.linalg.eigvalsh(mats).min() > 0                                         # candidates are SPD matrices
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and best[-1] == float(y.min())
    np.testing.assert_allclose(y.numpy(), ospd.ackley(x), rtol=1e-10)                 # test_functions_spd.py:34-69
