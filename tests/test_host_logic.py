"""Host-side logic that needs no GPU: parameter transforms of the kernel classes, solver / manifold recognition,
restart sharding and the one-all-gather record exchange (gloo, world_size 2)."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import gabotorch_b200 as g
from gabotorch_b200 import manifold_optimization as mo
from oracle import rcg as orcg


def test_beta_parameterisation_matches_gpytorch_greater_than():
    k = g.SphereGaussianKernel(beta_min=6.5)
    assert abs(float(k.beta.detach()) - (6.5 + math.log(2.0))) < 1e-6      # raw_beta = 0 -> beta_min + softplus(0)
    k.beta = 9.25
    assert abs(float(k.beta.detach()) - 9.25) < 1e-5
    assert tuple(k.raw_beta.shape) == (1, 1)
    s = g.SpdAffineInvariantGaussianKernel(beta_min=0.5)
    s.beta = 0.75
    assert abs(float(s.beta.detach()) - 0.75) < 1e-6
    sk = g.ScaleKernel(k)
    sk.outputscale = 2.5
    assert abs(float(sk.outputscale.detach()) - 2.5) < 1e-6
    lap = g.SphereLaplaceKernel()
    lap.lengthscale = 0.3
    assert abs(float(lap.lengthscale.detach()) - 0.3) < 1e-6


def test_spd_kernel_diagonal_shortcut_needs_no_gpu():
    k = g.SpdAffineInvariantGaussianKernel(beta_min=0.5)
    x = torch.randn(7, 6, dtype=torch.float64)
    out = k.forward(x, x, diagonal_distance=True)
    assert tuple(out.shape) == (7, 1) and float((out - 1).abs().max()) == 0.0
    # gpytorch passes diag=True; the SPD kernels' keyword is diagonal_distance, so diag lands in **params
    # (kernels_spd.py:72) and the full path is taken -> needs the device
    if not torch.cuda.is_available():
        with torch.no_grad(), pytest.raises(g.GaboError):
            k.forward(x, x, diag=True)


def test_unsupported_triples_raise():
    class TrustRegions:
        pass

    class Grassmann:
        pass
    with pytest.raises(NotImplementedError):
        mo._solver_options(TrustRegions())
    with pytest.raises(NotImplementedError):
        mo._manifold_kind(Grassmann())
    assert mo._manifold_kind(g.Sphere(4)) == 0 and mo._manifold_kind(g.PositiveDefinite(3)) == 1
    o = mo._solver_options(g.ConjugateGradient(maxiter=200, mingradnorm=1e-5))
    assert o['maxiter'] == 200 and o['mingradnorm'] == 1e-5 and o['ls_maxiter'] == 10 and o['contraction'] == 0.5
    with pytest.raises(NotImplementedError):
        g.PositiveDefinite(9)


def test_manifold_attributes():
    s = g.Sphere(6)
    assert s.dim == 5 and s._shape == (6,) and s.typicaldist == math.pi
    x = s.rand()
    assert x.shape == (6,) and abs(np.linalg.norm(x) - 1) < 1e-12
    p = g.PositiveDefinite(3)
    assert p.dim == 6 and p._n == 3 and abs(p.typicaldist - math.sqrt(6)) < 1e-12
    m = p.rand()
    assert np.all(np.linalg.eigvalsh(m) > 0.99)


def test_initialize_q_batch_nonneg_keeps_the_best_point():
    torch.manual_seed(0)
    X = torch.arange(50, dtype=torch.float64).reshape(50, 1, 1)
    Y = torch.rand(50, dtype=torch.float64)
    Y[17] = 5.0
    out = mo.initialize_q_batch_nonneg(X, Y, 8)
    assert out.shape == (8, 1, 1) and 17.0 in out.reshape(-1).tolist()
    with pytest.warns(mo.BadInitialCandidatesWarning):
        mo.initialize_q_batch_nonneg(X, torch.zeros(50, dtype=torch.float64), 8)


def test_shard_range_partitions_restarts():
    for num in (1, 5, 1024, 4097):
        for world in (1, 2, 3, 8):
            spans = [mo.shard_range(num, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == num
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # every rank "optimised" its shard of 10 restarts; values are a fixed array so the answer is known
        values = torch.tensor([0.3, 0.9, 0.1, 0.9, 0.2, float('nan'), 0.9, 0.0, 0.5, 0.4], dtype=torch.float64)
        cands = torch.arange(10, dtype=torch.float64).reshape(10, 1).repeat(1, 3) + 0.25
        lo, hi = mo.shard_range(10, rank, world)
        gidx = torch.arange(lo, hi)
        local = orcg.lexi_argmax_records(values[lo:hi].numpy(), gidx.numpy())   # selection rule (oracle = checker)
        v, gi, c = mo.allgather_records(values[lo:hi][local], gidx[local], cands[lo:hi][local])
        win = orcg.lexi_argmax_records(v.numpy(), gi.numpy())
        ret[rank] = (int(gi[win]), float(v[win]), c[win].tolist(), v.tolist(), gi.tolist())
    finally:
        dist.destroy_process_group()


def test_allgather_records_world_size_2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    a, b = ret[0], ret[1]
    assert a == b                                           # every rank reaches the same decision
    gi, v, c, allv, allg = a
    # single-process answer on the same value array: highest value, lowest global index on ties, NaN loses
    assert gi == 1 and v == 0.9 and c == [1.25, 1.25, 1.25]
    assert allg == [1, 6] and allv == [0.9, 0.9]


def test_gp_fit_transforms_priors_and_oracle_likelihood():
    # host side of gp_fit (no device call): gpytorch's softplus constraints round-trip, Gamma log-priors match
    # torch.distributions, and the oracle's likelihood equals the closed form -1/2 (r^T K^-1 r + log det K + n log 2 pi)
    import math
    from gabotorch_b200 import gp_fit
    from oracle import gp as ogp
    obj = gp_fit.MarginalLogLikelihood.__new__(gp_fit.MarginalLogLikelihood)
    obj.beta_min, obj.noise_min, obj.priors = 6.5, 1e-8, ((3.0, 2.0), (2.0, 0.15), (1.1, 0.05))
    theta = (7.25, 0.4, 2.0, -0.3)
    raw = obj.inverse_transform(theta)
    np.testing.assert_allclose(obj.transform(raw), theta, rtol=1e-12)
    assert abs(obj.transform([0.0, 0.0, 0.0, 1.0])[0] - (6.5 + math.log(2.0))) < 1e-15   # kernels_sphere.py:48-60
    lp, dlp = obj._prior_terms(theta)
    t64 = lambda v: torch.tensor(v, dtype=torch.float64)  # noqa: E731
    ref = sum(float(torch.distributions.Gamma(t64(c), t64(r)).log_prob(t64(v)))
              for (c, r), v in zip(obj.priors, theta[:3]))
    assert abs(lp - ref) < 1e-12
    assert abs(ogp.gamma_log_prob(0.4, 2.0, 0.15)
               - float(torch.distributions.Gamma(t64(2.0), t64(0.15)).log_prob(t64(0.4)))) < 1e-12
    for i, (c, r) in enumerate(obj.priors):
        assert abs(dlp[i] - ((c - 1.0) / theta[i] - r)) < 1e-15
    rng = np.random.default_rng(0)
    x = rng.standard_normal((9, 4))
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    dm = np.arccos(np.clip(x @ x.T, -1, 1)) ** 2
    y = rng.standard_normal(9)
    th = (2.5, 0.8, 0.05, 0.1)
    ll, grad = ogp.exact_log_likelihood(dm, y, th)
    k = th[1] * np.exp(-th[0] * dm) + th[2] * np.eye(9)
    r = y - th[3]
    closed = -0.5 * (r @ np.linalg.solve(k, r) + np.linalg.slogdet(k)[1] + 9 * math.log(2 * math.pi))
    assert abs(ll - closed) < 1e-10
    alpha = np.linalg.solve(k, r)
    w = np.outer(alpha, alpha) - np.linalg.inv(k)
    base = np.exp(-th[0] * dm)
    np.testing.assert_allclose(grad, [0.5 * np.sum(w * (-th[1] * dm * base)), 0.5 * np.sum(w * base),
                                      0.5 * np.trace(w), alpha.sum()], rtol=1e-8, atol=1e-10)


def test_trust_region_solver_surface_and_option_mapping():
    # constructor surface of the reference's TrustRegions (robust_trust_regions.py:91-108) and what reaches the kernel
    from gabotorch_b200 import manifold_optimization as mo
    s = mo.TrustRegions()
    assert (s.miniter, s.kappa, s.theta, s.rho_prime, s.use_rand, s.rho_regularization) == (3, 0.1, 1.0, 0.1, False, 1e3)
    assert (s._maxiter, s._mingradnorm, s._maxtime) == (1000, 1e-6, 1000)          # pymanopt Solver defaults
    o = mo._trust_region_options(mo.TrustRegions(kappa=0.2, maxiter=77, mingradnorm=1e-4, maxinner=3, Delta_bar=1.5))
    assert o == dict(maxiter=77, mingradnorm=1e-4, kappa=0.2, theta=1.0, rho_prime=0.1, rho_regularization=1e3,
                     mininner=1, maxinner=3, delta_bar=1.5, delta0=None)
    with pytest.raises(NotImplementedError):
        mo.TrustRegions(use_rand=True)

    class Foreign:            # a pymanopt-style solver object of another class with the same name and use_rand set
        use_rand = True
    Foreign.__name__ = 'TrustRegions'
    with pytest.raises(NotImplementedError):
        mo._trust_region_options(Foreign())
    with pytest.raises(NotImplementedError):
        mo._solver_options(type('SteepestDescent', (), {})())


class _OracleOps:
    """CPU stand-ins for the batched CUDA entry points the lock-step trust-region driver calls (same signatures as
    gabotorch_b200.ops), backed by the oracle: lets the masking / recurrence logic be checked without a GPU."""

    def __init__(self, gp):
        self.gp = gp

    def to_dev64(self, x):
        return torch.as_tensor(x, dtype=torch.float64)

    def ei_eval(self, gp, x, want_grad=False):
        from oracle import gp as ogp
        x = torch.as_tensor(x, dtype=torch.float64).numpy()
        ei, gr = [], []
        for p in x:
            try:
                e, g = ogp.ei_and_grad(self.gp, p, want_grad=True)
            except np.linalg.LinAlgError:
                e, g = np.nan, np.full(p.shape, np.nan)
            ei.append(e)
            gr.append(g)
        ei, gr = torch.tensor(np.array(ei)), torch.tensor(np.array(gr))
        return (ei, gr) if want_grad else ei

    def spd_scalar(self, what, x, b, c=None):
        from oracle import spd as ospd
        x, b = x.numpy(), b.numpy()
        if what == 1:
            return torch.tensor(np.array([ospd.norm(p, u) for p, u in zip(x, b)]))
        return torch.tensor(np.array([ospd.inner(p, u, v) for p, u, v in zip(x, b, c.numpy())]))

    def spd_op(self, op, a, b, c=None):
        from gabotorch_b200 import _lib
        from oracle import spd as ospd
        fn = {_lib.OP_RETR: ospd.retr, _lib.OP_EGRAD2RGRAD: ospd.egrad2rgrad}[op]
        return torch.tensor(np.array([fn(p, u) for p, u in zip(a.numpy(), b.numpy())]))

    def sym_eig(self, mat, vectors=True):
        # stand-in of the batched eigensolver kernel (gabo_sym_eig)
        lam, vec = torch.linalg.eigh(torch.as_tensor(mat, dtype=torch.float64))
        return lam, (vec if vectors else None), torch.zeros(1, dtype=torch.int32)

    def acq_ctr(self, gp, x0, constraints=(), strict=False, delta_cons=1e-6, maxiter=1000, mingradnorm=1e-6, kappa=0.1,
                theta=1.0, rho_prime=0.1, rho_regularization=1e3, mininner=1, maxinner=None, delta_bar=None,
                delta0=None):
        # stand-in of the one-launch constrained kernel (gabo_acq_ctr): the oracle's serial solve per restart
        from oracle import ctr as octr, rtr as ortr
        opts = ortr.TROptions(maxiter=maxiter, mingradnorm=mingradnorm, kappa=kappa, theta=theta, rho_prime=rho_prime,
                              rho_regularization=rho_regularization, mininner=mininner, maxinner=maxinner,
                              delta_bar=delta_bar, delta0=delta0)
        cons = [octr.max_eigenvalue_constraint(b) if k == 'max' else octr.min_eigenvalue_constraint(b)
                for k, b in constraints]
        res = [octr.solve_ctr(self.gp, p, ineq_constraints=cons, opts=opts, delta_cons=delta_cons, strict=strict)
               for p in torch.as_tensor(x0, dtype=torch.float64).numpy()]
        return (torch.from_numpy(np.array([r[0] for r in res])), torch.tensor([-r[1] for r in res], dtype=torch.float64),
                torch.tensor([r[2] for r in res], dtype=torch.int32), torch.full((len(res),), 2, dtype=torch.int32))


@pytest.mark.parametrize('manifold,dim,n,R', [('spd', 2, 10, 6), ('spd', 3, 12, 5), ('sphere', 20, 16, 6)])
def test_lockstep_trust_regions_follow_the_serial_solver(monkeypatch, manifold, dim, n, R):
    # batched_trust_regions advances all restarts together with masks; each restart must follow the serial algorithm
    # of the reference's TrustRegions (oracle/rtr.py, pinned on the reference's own class) iteration for iteration
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops
    from oracle import gp as ogp, rtr as ortr, sphere as osph, spd as ospd
    rng = np.random.default_rng(dim + n)
    if manifold == 'spd':
        xt = ospd.spd_sample(rng, n, dim, max_cond=50.0)
        y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt)))
        x0 = ospd.spd_sample(rng, R, dim, max_cond=50.0)
        beta = 0.5 + math.log(2.0)
    else:
        xt = osph.rand(rng, n, dim)
        y = osph.ackley(xt)
        x0 = osph.rand(rng, R, dim)
        beta = 0.35 + math.log(2.0)
    gp = ogp.make_gp(manifold, xt, y, beta=beta, noise=1e-2)
    fake = _OracleOps(gp)
    for name in ('to_dev64', 'ei_eval', 'spd_scalar', 'spd_op', 'sym_eig'):
        monkeypatch.setattr(ops, name, getattr(fake, name))
    handle = type('GP', (), {'manifold': _lib.SPD if manifold == 'spd' else _lib.SPHERE, 'dim': dim, 'n_train': n})()
    assert mo._rtr_kernel_covers(handle) == (manifold == 'spd')       # SPD(d) has its own one-launch kernel now
    X, val, iters, reason = mo.batched_trust_regions(handle, x0, maxiter=12)
    opts = ortr.TROptions(maxiter=12)
    for i in range(R):
        xi, ci, ki = ortr.solve_tr(gp, x0[i], opts)
        assert int(iters[i]) == ki
        np.testing.assert_allclose(X[i].numpy(), xi, rtol=0, atol=1e-9)
        assert abs(float(val[i]) + ci) <= 1e-10 * max(1.0, abs(ci))
        assert int(reason[i]) == (1 if ki >= 12 else 2)
    small = type('GP', (), {'manifold': _lib.SPHERE, 'dim': 6, 'n_train': 32})()
    assert mo._rtr_kernel_covers(small)


def test_trust_region_dispatch_chooses_kernel_or_lockstep_in_fp64(monkeypatch):
    # gen_candidates_manifold: small spheres and SPD(d) -> gabo_acq_rtr (one launch); eigenvalue-constrained SPD solves
    # -> gabo_acq_ctr (one launch); large spheres and generic constraint callables -> lock-step driver on an fp64
    # evaluator (no device needed: the solve entry points are replaced by recorders)
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops
    calls = []

    class FakeGP:
        def __init__(self, manifold, dim, n, compute=_lib.GABO_F32):
            self.manifold, self.dim, self.n_train, self.compute = manifold, dim, n, compute

        def with_compute(self, compute):
            return FakeGP(self.manifold, self.dim, self.n_train, compute)

    def recorder(name):
        def solve(gp, pts, **kw):
            calls.append((name, gp.compute, sorted(kw)))
            r = pts.shape[0]
            return pts, torch.zeros(r, dtype=torch.float64), torch.zeros(r, dtype=torch.int32), torch.zeros(r)
        return solve
    monkeypatch.setattr(ops, 'acq_rcg', recorder('rcg'))
    monkeypatch.setattr(ops, 'acq_rtr', recorder('rtr'))
    monkeypatch.setattr(ops, 'acq_ctr', recorder('ctr'))
    monkeypatch.setattr(mo, 'batched_trust_regions', recorder('lockstep'))
    monkeypatch.setattr(ops, 'to_dev64', lambda x: torch.as_tensor(x, dtype=torch.float64))

    def acq_for(gp):
        a = mo.ExpectedImprovement.__new__(mo.ExpectedImprovement)
        a._gp = gp
        return a
    sphere_x0 = torch.nn.functional.normalize(torch.randn(5, 1, 6, dtype=torch.float64), dim=-1)
    spd_x0 = torch.eye(3, dtype=torch.float64).expand(4, 1, 3, 3).clone()
    mo.gen_candidates_manifold(sphere_x0, acq_for(FakeGP(_lib.SPHERE, 6, 32)), g.Sphere(6), mo.TrustRegions())
    mo.gen_candidates_manifold(spd_x0, acq_for(FakeGP(_lib.SPD, 3, 32)), g.PositiveDefinite(3), mo.TrustRegions())
    mo.gen_candidates_manifold(spd_x0, acq_for(FakeGP(_lib.SPD, 3, 32)), g.PositiveDefinite(3), mo.ConjugateGradient())
    big = torch.nn.functional.normalize(torch.randn(3, 1, 40, dtype=torch.float64), dim=-1)
    mo.gen_candidates_manifold(big, acq_for(FakeGP(_lib.SPHERE, 40, 16)), g.Sphere(40), mo.TrustRegions())
    assert [(c[0], c[1]) for c in calls] == [('rtr', _lib.GABO_F32), ('rtr', _lib.GABO_F32),
                                             ('rcg', _lib.GABO_F32), ('lockstep', _lib.GABO_F64)]
    assert 'kappa' in calls[0][2] and 'kappa' in calls[1][2] and 'contraction' in calls[2][2]
    # constraints: ConstrainedTrustRegions + eigenvalue constraints on SPD -> the constrained kernel; any other callable
    # (here on a small sphere, and a lambda on SPD) -> lock-step driver with the constrained tCG (fp64); without
    # constraints it is the plain solver; everything else is refused
    import functools
    from gabotorch_b200 import riemannian_utils as ru
    cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=3.0)]
    calls.clear()
    mo.gen_candidates_manifold(spd_x0, acq_for(FakeGP(_lib.SPD, 3, 32)), g.PositiveDefinite(3),
                               mo.ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100), approx_hessian=True,
                               inequality_constraints=cons)
    mo.gen_candidates_manifold(sphere_x0, acq_for(FakeGP(_lib.SPHERE, 6, 32)), g.Sphere(6),
                               mo.ConstrainedTrustRegions(), inequality_constraints=[lambda x: x[0]])
    mo.gen_candidates_manifold(sphere_x0, acq_for(FakeGP(_lib.SPHERE, 6, 32)), g.Sphere(6), mo.ConstrainedTrustRegions())
    mo.gen_candidates_manifold(spd_x0, acq_for(FakeGP(_lib.SPD, 3, 32)), g.PositiveDefinite(3),
                               mo.ConstrainedTrustRegions(), inequality_constraints=[lambda x: 3.0 - x[0, 0]])
    assert [(c[0], c[1]) for c in calls] == [('ctr', _lib.GABO_F32), ('lockstep', _lib.GABO_F64),
                                             ('rtr', _lib.GABO_F32), ('lockstep', _lib.GABO_F64)]
    assert 'constraints' in calls[0][2] and 'delta_cons' in calls[0][2] and 'strict' in calls[0][2]
    assert 'ineq_constraints' in calls[1][2] and 'ineq_constraints' not in calls[2][2]
    assert mo.eigenvalue_constraint_specs(cons) == [('max', 3.0)]
    assert mo.eigenvalue_constraint_specs(cons + [functools.partial(ru.min_eigenvalue_constraint_torch, 0.01)]) == \
        [('max', 3.0), ('min', 0.01)]
    assert mo.eigenvalue_constraint_specs(cons * 3) is None and mo.eigenvalue_constraint_specs([len]) is None
    calls.clear()
    mo.gen_candidates_manifold(spd_x0, acq_for(FakeGP(_lib.SPD, 3, 32)), g.PositiveDefinite(3),
                               mo.StrictConstrainedTrustRegions(mingradnorm=2e-4, maxiter=100, minstepsize=1e-4),
                               approx_hessian=True, inequality_constraints=cons)
    assert calls[0][0] == 'ctr' and 'strict' in calls[0][2]
    with pytest.raises(NotImplementedError):
        mo.gen_candidates_manifold(sphere_x0, acq_for(FakeGP(_lib.SPHERE, 6, 32)), g.Sphere(6), mo.TrustRegions(),
                                   inequality_constraints=cons)
    calls.clear()
    mo.gen_candidates_manifold(sphere_x0, acq_for(FakeGP(_lib.SPHERE, 6, 32)), g.Sphere(6),
                               mo.ConstrainedTrustRegions(), equality_constraints=[lambda x: x[1]])
    assert calls[0][:2] == ('lockstep', _lib.GABO_F64) and 'eq_constraints' in calls[0][2]
    with pytest.raises(NotImplementedError):
        mo.gen_candidates_manifold(sphere_x0, acq_for(FakeGP(_lib.SPHERE, 6, 32)), g.Sphere(6),
                                   mo.ConstrainedTrustRegions(), equality_constraints=[lambda x: x[1]],
                                   inequality_constraints=cons)


def _screen_worker(rank, world, port, ret):
    import torch.distributed as dist
    from gabotorch_b200 import ops
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        ops.to_dev64 = lambda x: torch.as_tensor(x, dtype=torch.float64)       # CPU plumbing test: no device
        X = torch.linspace(-1.0, 2.0, 11 * 3, dtype=torch.float64).reshape(11, 1, 3)   # identical on every rank
        seen = []

        def acq(x):
            seen.append(x.shape[0])
            return (x.reshape(x.shape[0], -1) ** 2).sum(-1)
        vals = mo.sharded_acq_values(acq, X)
        ret[rank] = (vals.tolist(), seen)
    finally:
        dist.destroy_process_group()


def test_sharded_raw_sample_screening_world_size_2_gloo():
    # SURVEY 8e, raw-sample screening: every rank evaluates only its block, one all-gather returns the whole vector
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_screen_worker, args=(world, port, ret), nprocs=world, join=True)
    X = torch.linspace(-1.0, 2.0, 11 * 3, dtype=torch.float64).reshape(11, 3)
    full = (X ** 2).sum(-1).tolist()
    assert ret[0][0] == full and ret[1][0] == full
    assert ret[0][1] == [5] and ret[1][1] == [6]                # 11 samples: blocks [0, 5) and [5, 11)


@pytest.mark.parametrize('name', ['ctr_spd2_active', 'ctr_spd2', 'ctr_spd3', 'sctr_spd2_active', 'sctr_spd3'])
@pytest.mark.parametrize('closed_form', [True, False])
def test_lockstep_constrained_trust_regions_reproduce_the_reference_solver(monkeypatch, golden, name, closed_form):
    # the golden arrays come from the reference's OWN ConstrainedTrustRegions class in gabo_spd.py's configuration
    # (tests/golden/make_golden.py); the lock-step driver must reproduce its iteration counts and candidates, with the
    # constraint batched in closed form and through the generic autograd route
    import functools
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops, riemannian_utils as ru
    from oracle import gp as ogp
    beta, noise, max_eig = (float(v) for v in golden[name + '_hyper'])
    xt = golden[name + '_xtrain']
    gp = ogp.make_gp('spd', xt, golden[name + '_y'], beta=beta, noise=noise)
    fake = _OracleOps(gp)
    for attr in ('to_dev64', 'ei_eval', 'spd_scalar', 'spd_op', 'sym_eig'):
        monkeypatch.setattr(ops, attr, getattr(fake, attr))
    if closed_form:
        cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=max_eig)]
    else:
        cons = [lambda x: max_eig - torch.linalg.eigvalsh(x)[-1]]
    handle = type('GP', (), {'manifold': _lib.SPD, 'dim': xt.shape[-1], 'n_train': xt.shape[0]})()
    strict = name.startswith('sctr')          # StrictConstrainedTrustRegions with hd_gabo_spd.py's mingradnorm
    X, val, iters, reason = mo.batched_trust_regions(handle, golden[name + '_x0'], maxiter=100,
                                                     mingradnorm=2e-4 if strict else 1e-4, strict=strict,
                                                     ineq_constraints=mo.batched_constraints(cons, _lib.SPD))
    np.testing.assert_array_equal(iters.numpy(), golden[name + '_iters'])
    np.testing.assert_allclose(X.numpy(), golden[name + '_x'], rtol=0, atol=1e-8)
    np.testing.assert_allclose(-val.numpy(), golden[name + '_cost'], rtol=1e-9, atol=1e-13)


def test_gabo_spd_iteration_through_the_public_api_with_emulated_kernels(monkeypatch):
    # The acquisition step of examples/bo_spd/benchmark_examples/gabo_spd.py:136-203 through the drop-in surface:
    # joint_optimize_manifold(EI, PositiveDefinite, ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100), Mandel
    # pre / post processing, approx_hessian=True, one max-eigenvalue inequality constraint).  The CUDA entry points
    # are replaced by oracle-backed stand-ins, so this pins the HOST wiring (shapes through pre / post processing,
    # solver dispatch, candidate selection) without a device; the kernels themselves are covered by the -m gpu tests.
    import functools
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops, riemannian_utils as ru
    from oracle import ctr as octr, gp as ogp, rtr as ortr, spd as ospd
    d, n, max_eig = 2, 10, 2.0
    rng = np.random.default_rng(17)
    xt = ospd.spd_sample(rng, n, d, max_cond=50.0)
    y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt)))
    gp = ogp.make_gp('spd', xt, y, beta=0.5 + math.log(2.0), noise=1e-2)
    fake = _OracleOps(gp)
    for attr in ('to_dev64', 'ei_eval', 'spd_scalar', 'spd_op', 'acq_ctr'):
        monkeypatch.setattr(ops, attr, getattr(fake, attr))
    monkeypatch.setattr(ops, 'device', lambda: torch.device('cpu'))
    monkeypatch.setattr(ops, 'mandel_unpack', lambda v: ospd.vector_to_symmetric_matrix_mandel(torch.as_tensor(v)))
    monkeypatch.setattr(ops, 'mandel_pack', lambda m: ospd.symmetric_matrix_to_vector_mandel(torch.as_tensor(m)))

    def argmax_records(values, gidx=None):
        v = torch.nan_to_num(torch.as_tensor(values, dtype=torch.float64).reshape(-1), nan=-float('inf'))
        slot = torch.argmax(v).reshape(1)
        return slot, v[slot]
    monkeypatch.setattr(ops, 'argmax_records', argmax_records)

    class Handle:                                   # what ExpectedImprovement.device_gp() hands to the solvers
        manifold, dim, n_train = _lib.SPD, d, n

        def with_compute(self, compute):
            return self
    acq = mo.ExpectedImprovement.__new__(mo.ExpectedImprovement)
    acq._gp = Handle()
    man = g.PositiveDefinite(d)
    man.min_eig, man.max_eig = 0.5, 0.9 * max_eig                      # feasible raw samples (spd_sample law)
    cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=max_eig)]
    solver = mo.ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=30)
    kw = dict(pre_processing_manifold=ru.vector_to_symmetric_matrix_mandel_torch,
              post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch, approx_hessian=True,
              inequality_constraints=cons)
    ics = mo.gen_batch_initial_conditions_manifold(acq, man, None, 1, 4, 24, options={'seed': 5},
                                                   post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch)
    assert tuple(ics.shape) == (4, 1, 3)
    cand, vals = mo.gen_candidates_manifold(ics, acq, man, solver, **kw)
    assert tuple(cand.shape) == (4, 1, 3) and tuple(vals.shape) == (4,)
    opts = ortr.TROptions(mingradnorm=1e-4, maxiter=30)
    mats0 = ospd.vector_to_symmetric_matrix_mandel(ics[:, 0]).numpy()
    for i in range(4):                              # every restart equals the serial oracle solve from the same start
        xi, ci, _ = octr.solve_ctr(gp, mats0[i], ineq_constraints=[octr.max_eigenvalue_constraint(max_eig)], opts=opts)
        np.testing.assert_allclose(ospd.vector_to_symmetric_matrix_mandel(cand[i]).numpy()[0], xi, rtol=0, atol=1e-8)
        assert abs(float(vals[i]) + ci) <= 1e-9 * max(1.0, abs(ci))
    best = mo.joint_optimize_manifold(acq, man, solver, q=1, num_restarts=4, raw_samples=24, options={'seed': 5}, **kw)
    assert tuple(best.shape) == (1, 3)
    np.testing.assert_allclose(best.numpy(), cand[int(torch.argmax(vals))].numpy(), rtol=0, atol=1e-12)
    # (feasibility of the result is NOT asserted: the non-strict solver only keeps the LINEARISED constraints, and the
    # reference's own class leaves the feasible set in the same way -- golden set ctr_spd2_active)


@pytest.mark.parametrize('name', ['ctr_s2_domain', 'ctr_s2_domain_active'])
def test_lockstep_constrained_trust_regions_on_the_sphere_reproduce_the_reference_solver(monkeypatch, golden, name):
    # sphere + a user-supplied torch callable (the domain constraint of gabo_sphere_inequality_constraints.py:113-120):
    # the generic route of batched_constraints (torch.autograd per restart + tangent projection) in the lock-step driver
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops
    from oracle import gp as ogp
    beta, noise, angle = (float(v) for v in golden[name + '_hyper'])
    xt = golden[name + '_xtrain']
    gp = ogp.make_gp('sphere', xt, golden[name + '_y'], beta=beta, noise=noise)
    fake = _OracleOps(gp)
    for attr in ('to_dev64', 'ei_eval'):
        monkeypatch.setattr(ops, attr, getattr(fake, attr))

    def domain_constraint(x):
        centre = torch.zeros(3, dtype=x.dtype)
        centre[0] = 1
        in_prod = torch.mm(x[None], centre[:, None])
        in_prod = torch.max(torch.min(in_prod, torch.ones(1, dtype=x.dtype)), -torch.ones(1, dtype=x.dtype))
        return angle - torch.acos(in_prod)[0, 0]
    handle = type('GP', (), {'manifold': _lib.SPHERE, 'dim': 3, 'n_train': xt.shape[0]})()
    X, val, iters, _ = mo.batched_trust_regions(handle, golden[name + '_x0'][:5], maxiter=200,
                                                ineq_constraints=mo.batched_constraints([domain_constraint], _lib.SPHERE))
    np.testing.assert_array_equal(iters.numpy(), golden[name + '_iters'][:5])
    np.testing.assert_allclose(X.numpy(), golden[name + '_x'][:5], rtol=0, atol=1e-8)
    np.testing.assert_allclose(-val.numpy(), golden[name + '_cost'][:5], rtol=1e-9, atol=1e-13)


def test_lockstep_equality_constrained_trust_regions_reproduce_the_reference_solver(monkeypatch, golden):
    # the great-circle equality constraint of gabo_sphere_equality_constraints.py:104-109 as a plain torch callable
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops
    from oracle import gp as ogp
    name = 'ctr_s2_circle'
    beta, noise = (float(v) for v in golden[name + '_hyper'])
    xt = golden[name + '_xtrain']
    gp = ogp.make_gp('sphere', xt, golden[name + '_y'], beta=beta, noise=noise)
    fake = _OracleOps(gp)
    for attr in ('to_dev64', 'ei_eval'):
        monkeypatch.setattr(ops, attr, getattr(fake, attr))
    handle = type('GP', (), {'manifold': _lib.SPHERE, 'dim': 3, 'n_train': xt.shape[0]})()
    cons = mo.batched_constraints([lambda x: x[1] - 0.], _lib.SPHERE)
    X, val, iters, _ = mo.batched_trust_regions(handle, golden[name + '_x0'], maxiter=200, eq_constraints=cons)
    np.testing.assert_array_equal(iters.numpy(), golden[name + '_iters'])
    np.testing.assert_allclose(X.numpy(), golden[name + '_x'], rtol=0, atol=1e-8)
    np.testing.assert_allclose(-val.numpy(), golden[name + '_cost'], rtol=1e-9, atol=1e-13)
    with pytest.raises(NotImplementedError):
        mo.batched_trust_regions(handle, golden[name + '_x0'], eq_constraints=cons, ineq_constraints=cons)


@pytest.mark.parametrize('name,kind', [('alm_s2_domain', 'ineq'), ('alm_s2_circle', 'eq')])
def test_lockstep_augmented_lagrangian_reproduces_the_reference_solver(monkeypatch, golden, name, kind):
    # golden arrays: the reference's own AugmentedLagrangeMethod class around its own TrustRegions (make_golden.py);
    # here through gen_candidates_manifold with the solver objects a user of the reference would build
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops
    from oracle import gp as ogp
    beta, noise, angle = (float(v) for v in golden[name + '_hyper'])
    xt = golden[name + '_xtrain']
    gp = ogp.make_gp('sphere', xt, golden[name + '_y'], beta=beta, noise=noise)
    fake = _OracleOps(gp)
    for attr in ('to_dev64', 'ei_eval'):
        monkeypatch.setattr(ops, attr, getattr(fake, attr))

    def domain_constraint(x):
        centre = torch.zeros(3, dtype=x.dtype)
        centre[0] = 1
        in_prod = torch.mm(x[None], centre[:, None])
        in_prod = torch.max(torch.min(in_prod, torch.ones(1, dtype=x.dtype)), -torch.ones(1, dtype=x.dtype))
        return angle - torch.acos(in_prod)[0, 0]

    class Handle:
        manifold, dim, n_train = _lib.SPHERE, 3, xt.shape[0]

        def with_compute(self, compute):
            return self
    acq = mo.ExpectedImprovement.__new__(mo.ExpectedImprovement)
    acq._gp = Handle()
    solver = mo.AugmentedLagrangeMethod(inner_solver=mo.TrustRegions(maxiter=50), maxiter=30, gammas_fact=0.05)
    kw = ({'inequality_constraints': [domain_constraint]} if kind == 'ineq'
          else {'equality_constraints': [lambda x: x[1] - 0.]})
    x0 = torch.from_numpy(golden[name + '_x0'])[:, None, :]
    cand, vals, info = mo.gen_candidates_manifold(x0, acq, g.Sphere(3), solver, return_info=True, **kw)
    # the step-size rule dist(x_k, x_{k-1}) < 1e-10 fires only when two consecutive outer iterates agree to the last
    # bits (acos(1 - 1.1e-16) is already 1.5e-8): with the equality constraint the reference's own run stops there for
    # some starts, and a last-bit difference moves that stop by many iterations without moving the candidate
    same_iters = info['iters'].numpy() == golden[name + '_iters']
    assert same_iters.all() if kind == 'ineq' else same_iters.mean() >= 0.5
    # 30 outer iterations multiply the penalty parameter up to 0.3^-30: rounding differences between the batched torch
    # expressions and the reference's numpy ones are amplified to ~1e-7 by then (they agree to 1e-14 after 20
    # iterations), hence the looser tolerance here and the tight comparison on a 12-iteration run below
    np.testing.assert_allclose(cand[:, 0].numpy()[same_iters], golden[name + '_x'][same_iters], rtol=0, atol=2e-6)
    if kind == 'eq':    # where the stop moved, the run went on along the circle: still feasible, still on the sphere
        assert np.abs(cand[:, 0].numpy()[:, 1]).max() <= 1e-4
    np.testing.assert_allclose(np.linalg.norm(cand[:, 0].numpy(), axis=-1), 1.0, atol=1e-12)
    assert tuple(vals.shape) == (len(golden[name + '_x0']),)
    from oracle import alm as oalm, ctr as octr, sphere as osph
    short = mo.AugmentedLagrangeMethod(inner_solver=mo.TrustRegions(maxiter=50), maxiter=12, gammas_fact=0.05)
    cand12, _, info12 = mo.gen_candidates_manifold(x0, acq, g.Sphere(3), short, return_info=True, **kw)
    e1 = np.array([0.0, 1.0, 0.0])
    okw = ({'ineq_constraints': [octr.sphere_domain_constraint([1.0, 0.0, 0.0], angle)]} if kind == 'ineq'
           else {'eq_constraints': [(lambda x: x[1], lambda x: osph.proj(x, e1))]})
    for i in range(x0.shape[0]):
        xo, ko = oalm.solve_alm(gp, golden[name + '_x0'][i], maxiter=12, inner_opts={'maxiter': 50}, gammas_fact=0.05,
                                **okw)
        if int(info12['iters'][i]) == ko:
            np.testing.assert_allclose(cand12[i, 0].numpy(), xo, rtol=0, atol=1e-10)
        else:
            assert kind == 'eq' and abs(float(cand12[i, 0, 1])) <= 1e-4
    with pytest.raises(NotImplementedError):
        mo.AugmentedLagrangeMethod(inner_solver=object())


def test_batched_constraints_closed_form_equals_autograd(monkeypatch):
    # max / min eigenvalue constraints: the closed-form batch evaluation (extreme eigenpair) must equal what
    # torch.autograd gives one point at a time (the reference's route, pymanopt_addons/problem.py:118-137), values and
    # Riemannian gradients X sym(G) X
    import functools
    from gabotorch_b200 import _lib, manifold_optimization as mo, ops, riemannian_utils as ru
    from oracle import spd as ospd
    monkeypatch.setattr(ops, 'spd_op', _OracleOps(None).spd_op)
    monkeypatch.setattr(ops, 'sym_eig', _cpu_sym_eig)            # stand-in of the batched eigensolver kernel
    X = torch.from_numpy(ospd.spd_sample(np.random.default_rng(1), 9, 4, max_cond=50.0))
    closed = mo.batched_constraints([functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=3.5),
                                     functools.partial(ru.min_eigenvalue_constraint_torch, minimum_eigenvalue=0.2)],
                                    _lib.SPD)
    generic = mo.batched_constraints([lambda x: 3.5 - torch.linalg.eigvalsh(x)[-1],
                                      lambda x: torch.linalg.eigvalsh(x)[0] - 0.2], _lib.SPD)
    fc, gc = closed(X)
    fg, gg = generic(X)
    assert tuple(fc.shape) == (9, 2) and len(gc) == 2
    np.testing.assert_allclose(fc.numpy(), fg.numpy(), rtol=0, atol=1e-13)
    for a, b in zip(gc, gg):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=0, atol=1e-11)
        np.testing.assert_allclose(a.numpy(), np.swapaxes(a.numpy(), -1, -2), rtol=0, atol=1e-12)
    lam = np.linalg.eigvalsh(X.numpy())
    np.testing.assert_allclose(fc.numpy(), np.stack([3.5 - lam[:, -1], lam[:, 0] - 0.2], -1), rtol=0, atol=1e-13)


def test_riemannian_cg_on_a_product_manifold_finds_the_known_minimiser():
    """Host logic of fit_gpytorch_manifold (manifold_gp_fit.py:54-222): pymanopt-style CG on Product(Euclidean, Grassmann,
    Sphere) with numpy points.  Problem with a closed-form answer: minimise (t - 2)^2 - tr(X^T A X) - v^T B v, whose
    minimum is -(sum of the two largest eigenvalues of A) - (largest eigenvalue of B) at t = 2."""
    from gabotorch_b200 import manifold_gp_fit as mgf
    rng = np.random.default_rng(0)
    np.random.seed(0)
    A = rng.standard_normal((6, 6)); A = A @ A.T
    B = rng.standard_normal((4, 4)); B = B @ B.T
    man = mgf.ProductParam([mgf.EuclideanParam(1), mgf.GrassmannParam(6, 2), mgf.SphereParam(4)])

    def cost(x):
        t, X, v = x
        return float((t[0] - 2.0) ** 2 - np.trace(X.T @ A @ X) - v @ B @ v)

    def cost_grad(x):
        t, X, v = x
        return cost(x), [np.array([2.0 * (t[0] - 2.0)]), -2.0 * A @ X, -2.0 * B @ v]

    x0 = man.rand()
    x, log = mgf.riemannian_cg(man, cost, cost_grad, x0, g.ConjugateGradient(maxiter=400, mingradnorm=1e-8))
    want = -np.sort(np.linalg.eigvalsh(A))[-2:].sum() - np.linalg.eigvalsh(B)[-1]
    assert abs(log['cost'] - want) <= 1e-8 * abs(want), (log, want)
    assert abs(x[0][0] - 2.0) < 1e-6
    assert np.abs(x[1].T @ x[1] - np.eye(2)).max() < 1e-12 and abs(np.linalg.norm(x[2]) - 1.0) < 1e-12
    assert log['stop'] in ('mingradnorm', 'minstepsize') and log['iterations'] < 400
    # the retraction / projection pairs satisfy the manifold identities the solver relies on
    u = man.proj(x, man.rand())
    assert abs(np.trace(x[1].T @ u[1])) < 1e-12 and abs(x[2] @ u[2]) < 1e-12
    assert mgf.host_manifold(mgf.GrassmannParam(5, 2))._n == 5


def test_riemannian_trust_regions_on_a_product_manifold():
    """Host trust-region solver behind optimize_reconstruction_parameters_nested_sphere (pymanopt TrustRegions semantics,
    finite-difference Hessian): same closed-form problem as the CG test."""
    from gabotorch_b200 import manifold_gp_fit as mgf
    rng = np.random.default_rng(1)
    np.random.seed(1)
    A = rng.standard_normal((6, 6)); A = A @ A.T
    B = rng.standard_normal((4, 4)); B = B @ B.T
    man = mgf.ProductParam([mgf.EuclideanParam(1), mgf.GrassmannParam(6, 2), mgf.SphereParam(4)])
    assert man.dim == 1 + 8 + 3

    def cost(x):
        t, X, v = x
        return float((t[0] - 2.0) ** 2 - np.trace(X.T @ A @ X) - v @ B @ v)

    def cost_grad(x):
        t, X, v = x
        return cost(x), [np.array([2.0 * (t[0] - 2.0)]), -2.0 * A @ X, -2.0 * B @ v]

    x, log = mgf.solve_on_manifold(man, cost, cost_grad, man.rand(), g.TrustRegions(maxiter=200, mingradnorm=1e-7))
    want = -np.sort(np.linalg.eigvalsh(A))[-2:].sum() - np.linalg.eigvalsh(B)[-1]
    assert log['stop'] == 'mingradnorm' and log['iterations'] < 200, log
    assert abs(log['cost'] - want) <= 1e-9 * abs(want), (log, want)
    assert np.abs(x[1].T @ x[1] - np.eye(2)).max() < 1e-12 and abs(np.linalg.norm(x[2]) - 1.0) < 1e-12
    with pytest.raises(NotImplementedError):
        mgf.solve_on_manifold(man, cost, cost_grad, man.rand(), object())


def test_nested_sphere_reconstruction_cost_and_fit_host_logic(monkeypatch):
    """nested_spheres_optimization.py:20-98: the differentiable inverse chain equals the oracle's (pinned on the reference's
    own functions), its gradient matches finite differences, and the fit recovers the distances that generated the data.
    The tensor code is device-agnostic; here it runs on CPU tensors (ops.to_dev64 replaced), the device run is in
    tests/test_nested_gpu.py."""
    from gabotorch_b200 import nested_optimization as nopt
    from oracle import nested_sphere as onsph, sphere as osph
    monkeypatch.setattr(nopt.ops, 'to_dev64', lambda x: torch.as_tensor(x, dtype=torch.float64))
    monkeypatch.setattr(nopt, '_dev64_keep_grad', lambda x: torch.as_tensor(x, dtype=torch.float64))
    rng = np.random.default_rng(5)
    np.random.seed(5)
    D, d, n = 6, 3, 40
    axes = [torch.from_numpy(osph.rand(rng, 1, k)) for k in range(D, d, -1)]
    true_r = [1.1, 0.7, 1.9]
    xs = torch.from_numpy(osph.rand(rng, n, d))
    dists = [torch.tensor([[r]], dtype=torch.float64) for r in true_r]
    xd = onsph.projection_from_subsphere_to_sphere(xs, axes, dists)[-1]
    ours = nopt._reconstruct(xs, [a.reshape(-1) for a in axes], torch.tensor(true_r, dtype=torch.float64))
    assert float((ours - xd).abs().max()) < 1e-13
    assert float(nopt.min_error_reconstruction_cost(xd, xs, axes, dists)) < 1e-12
    # gradient of the cost with respect to the distances vs central differences
    p = torch.tensor([0.9, 1.0, 1.5], dtype=torch.float64, requires_grad=True)
    f = nopt._cost_from_distances(xd, xs, [a.reshape(-1) for a in axes], p)
    f.backward()
    for i in range(3):
        e = torch.zeros(3, dtype=torch.float64); e[i] = 1e-6
        fd = (nopt._cost_from_distances(xd, xs, [a.reshape(-1) for a in axes], p.detach() + e)
              - nopt._cost_from_distances(xd, xs, [a.reshape(-1) for a in axes], p.detach() - e)) / 2e-6
        assert abs(float(fd) - float(p.grad[i])) <= 1e-5 * max(1.0, abs(float(fd)))
    for solver in (g.TrustRegions(), g.ConjugateGradient(maxiter=300)):
        out = g.optimize_reconstruction_parameters_nested_sphere(xd, xs, axes, solver, nb_init_candidates=50)
        assert len(out) == 3 and all(tuple(o.shape) == (1,) and o.dtype == torch.float32 for o in out)
        got = np.array([float(o) for o in out])
        assert np.abs(got - np.array(true_r)).max() < 2e-3, (got, true_r)


def _cpu_sym_eig(mat, vectors=True):
    """Test stand-in of ops.sym_eig (the device kernel) for the CPU runs of the tensor code around it."""
    lam, vec = torch.linalg.eigh(torch.as_tensor(mat, dtype=torch.float64))
    return lam, (vec if vectors else None), torch.zeros(1, dtype=torch.int32)


def _nested_spd_on_cpu(monkeypatch):
    from gabotorch_b200 import nested_optimization as nopt
    monkeypatch.setattr(nopt.ops, 'to_dev64', lambda x: torch.as_tensor(x).detach().to(torch.float64).contiguous())
    monkeypatch.setattr(nopt, '_dev64_keep_grad', lambda x: torch.as_tensor(x, dtype=torch.float64))
    monkeypatch.setattr(nopt.ops, 'sym_eig', _cpu_sym_eig)
    return nopt


def test_nested_spd_reconstruction_costs_match_reference(monkeypatch):
    """nested_spd_optimization.py:22-92: both costs against values produced by the reference's own functions
    (tests/golden/make_golden_recon_cost.py; the reference accumulates in float32, hence 5e-6), and their gradients with
    respect to (V, C, K) against central differences.  Host run of the tensor code (the eigensolver kernel replaced by a
    stand-in); the device run is tests/test_nested_gpu.py."""
    nopt = _nested_spd_on_cpu(monkeypatch)
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'recon_cost_vectors.npz'))
    for name in ('rc_6_2', 'rc_10_3', 'rc_20_5'):
        args = [torch.from_numpy(gold[name + '_' + k]) for k in ('x', 'y', 'w', 'v', 'c', 'k')]
        ai = float(nopt.min_affine_invariant_distance_reconstruction_cost(*args))
        le = float(nopt.min_log_euclidean_distance_reconstruction_cost(*args))
        assert abs(ai - float(gold[name + '_ai'])) <= 5e-6 * abs(ai), (name, ai, float(gold[name + '_ai']))
        assert abs(le - float(gold[name + '_le'])) <= 5e-6 * abs(le), (name, le, float(gold[name + '_le']))
    rng = np.random.default_rng(3)
    name = 'rc_6_2'
    x, y, w = (torch.from_numpy(gold[name + '_' + k]) for k in ('x', 'y', 'w'))
    for kind in ('affine_invariant', 'log_euclidean'):
        fn = nopt._SpdReconstructionCost(x, y, w, kind)
        params = [torch.from_numpy(gold[name + '_' + k]).clone().requires_grad_(True) for k in ('v', 'c', 'k')]
        fn(*params).backward()
        for i, p in enumerate(params):
            direction = torch.from_numpy(rng.standard_normal(tuple(p.shape)))
            if i == 1:
                direction = 0.5 * (direction + direction.T)           # C stays symmetric
            h = 1e-6
            plus = [q.detach() + (h * direction if j == i else 0.0) for j, q in enumerate(params)]
            minus = [q.detach() - (h * direction if j == i else 0.0) for j, q in enumerate(params)]
            fd = (float(fn(*plus)) - float(fn(*minus))) / (2 * h)
            an = float((p.grad * direction).sum())
            assert abs(fd - an) <= 2e-6 * max(1.0, abs(fd)), (kind, i, fd, an)
    # candidate batches: a leading dimension gives the same values as one call per candidate
    fn = nopt._SpdReconstructionCost(x, y, w, 'affine_invariant')
    vs = torch.stack([torch.from_numpy(gold[name + '_v']), torch.from_numpy(np.linalg.qr(rng.standard_normal((6, 4)))[0])])
    cs = torch.stack([torch.from_numpy(gold[name + '_c']), torch.eye(4, dtype=torch.float64) * 1.5])
    ks = torch.stack([torch.from_numpy(gold[name + '_k']), torch.zeros(2, 4, dtype=torch.float64)])
    both = fn(vs, cs, ks)
    for i in range(2):
        assert abs(float(both[i]) - float(fn(vs[i], cs[i], ks[i]))) <= 1e-12 * abs(float(both[i]))


def test_riemannian_alm_and_spd_parameter_manifold():
    """Host ALM (augmented_Lagrange_method.py:66-328) on product manifolds, and the PositiveDefinite parameter manifold."""
    from gabotorch_b200 import manifold_gp_fit as mgf
    np.random.seed(11)
    # 1. max <a, x> on S^2 subject to x_1 = 0  ->  x = (a_0, 0, a_2) / |.|; a scalar rides along (t - 1)^2
    a = np.array([0.3, 0.8, -0.5])
    man = mgf.ProductParam([mgf.SphereParam(3), mgf.EuclideanParam(1)])

    def cost(p):
        return float(-a @ p[0] + (p[1][0] - 1.0) ** 2)

    def cost_grad(p):
        return cost(p), [-a, np.array([2.0 * (p[1][0] - 1.0)])]

    def plane(p):
        return float(p[0][1]), [np.array([0.0, 1.0, 0.0]), np.zeros(1)]

    want = np.array([a[0], 0.0, a[2]]) / np.hypot(a[0], a[2])
    for inner in (g.TrustRegions(maxiter=50), g.ConjugateGradient(maxiter=100)):
        # tight inner tolerances: with the defaults (1e-3 -> 1e-6) the loop ends as soon as an inner solve starts below its
        # tolerance and returns its start (step size 0 < minstepsize), exactly like the reference's loop
        solver = g.AugmentedLagrangeMethod(inner_solver=inner, maxiter=40, starting_tolgradnorm=1e-6,
                                           ending_tolgradnorm=1e-9)
        x, log = mgf.riemannian_alm(man, cost, cost_grad, man.rand(), solver, eq_constraints=plane)
        # pymanopt's adaptive line search restarts every inner solve from a unit-length step and gives up after 10
        # contractions: close to the solution the CG inner solver stalls ('minstepsize'), TrustRegions does not
        tol = 1e-4 if type(inner).__name__ == 'TrustRegions' else 5e-3
        assert np.abs(x[0] - want).max() < tol and abs(x[1][0] - 1.0) < tol, (x, log)
        assert log['violation'] < tol and log['iterations'] <= 40
        loose, llog = mgf.riemannian_alm(man, cost, cost_grad, man.rand(),
                                         g.AugmentedLagrangeMethod(inner_solver=inner, maxiter=40), eq_constraints=plane)
        assert np.abs(loose[0] - want).max() < 5e-3 and llog['stop'] in ('minstepsize', 'mingradnorm', 'maxiter'), llog
        assert inner._mingradnorm == 1e-6                      # the caller's solver object is not modified
    # inequality: the same objective with x_1 <= 0.2  <=>  0.2 - x_1 >= 0 (active at the solution)
    def cap(p):
        return float(0.2 - p[0][1]), [np.array([0.0, -1.0, 0.0]), np.zeros(1)]
    solver = g.AugmentedLagrangeMethod(inner_solver=g.TrustRegions(maxiter=50), maxiter=40, starting_tolgradnorm=1e-6,
                                       ending_tolgradnorm=1e-9)
    x, log = mgf.riemannian_alm(man, cost, cost_grad, man.rand(), solver, ineq_constraints=[cap])
    # the reference's multiplier update for inequalities is lambda + rho * g(x) (augmented_Lagrange_method.py:156-159),
    # which also grows on the feasible side and so keeps the iterate strictly inside: feasibility is what it guarantees
    assert x[0][1] <= 0.2 + 1e-3 and abs(np.linalg.norm(x[0]) - 1.0) < 1e-12 and log['lambdas'][0] >= 0.0, (x, log)
    # 2. SPD parameter manifold: retraction = exponential map, metric, gradient conversion; CG and TR find C0
    spd = mgf.SpdParam(4)
    c0 = spd.rand() * 3.0
    pman = mgf.ProductParam([spd])

    def cost2(p):
        return float(((p[0] - c0) ** 2).sum())

    def cost2_grad(p):
        return cost2(p), [2.0 * (p[0] - c0)]
    for solver in (g.TrustRegions(maxiter=200), g.ConjugateGradient(maxiter=500)):
        x, log = mgf.solve_on_manifold(pman, cost2, cost2_grad, pman.rand(), solver)
        assert np.abs(x[0] - c0).max() < 1e-5, log
        assert np.linalg.eigvalsh(x[0]).min() > 0
    xa, u = spd.rand(), spd.proj(None, np.random.randn(4, 4))
    assert abs(spd.dist(xa, spd.retr(xa, 0.2 * u)) - 0.2 * spd.norm(xa, u)) < 1e-12
    assert abs(spd.inner(xa, u, u) - spd.norm(xa, u) ** 2) < 1e-12
    assert mgf.host_manifold(type('PositiveDefinite', (), {'_n': 3})()).dim == 6


def test_nested_spd_reconstruction_fit_host_logic(monkeypatch):
    """optimize_reconstruction_parameters_nested_spd (nested_spd_optimization.py:95-186) on data that the mapping can
    reproduce exactly: the fit drives the cost from the best random candidate down by orders of magnitude, keeps
    W^T V = 0, returns an SPD bottom block and a contraction."""
    nopt = _nested_spd_on_cpu(monkeypatch)
    rng = np.random.default_rng(8)
    np.random.seed(8)
    D, d, n = 5, 2, 10
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    w, v = q[:, :d], q[:, d:]
    c = np.diag([1.0, 1.5, 2.0])
    k = rng.standard_normal((d, D - d))
    k = 0.5 * k / np.linalg.norm(k)
    ys = []
    for _ in range(n):
        b = rng.standard_normal((d, d))
        ys.append(b @ b.T + 0.5 * np.eye(d))
    y = torch.from_numpy(np.array(ys))
    x = nopt._reconstruct_spd(y, nopt._SpectralFn.apply(y, 1), torch.from_numpy(w), torch.from_numpy(v),
                              torch.from_numpy(c), torch.from_numpy(k))
    for cost_fn in (g.min_affine_invariant_distance_reconstruction_cost,
                    g.min_log_euclidean_distance_reconstruction_cost):
        assert float(cost_fn(x, y, torch.from_numpy(w), torch.from_numpy(v), torch.from_numpy(c),
                             torch.from_numpy(k))) < 1e-10
    vo, co, ko = g.optimize_reconstruction_parameters_nested_spd(
        x, y, torch.from_numpy(w), g.ConjugateGradient(maxiter=100),
        cost_function=g.min_log_euclidean_distance_reconstruction_cost, nb_init_candidates=30, maxiter=30)
    log = g.optimize_reconstruction_parameters_nested_spd.last_log
    assert tuple(vo.shape) == (D, D - d) and tuple(co.shape) == (D - d, D - d) and tuple(ko.shape) == (d, D - d)
    assert vo.dtype == co.dtype == ko.dtype == torch.float64
    assert log['cost'] < 0.05 * log['start_cost'], log
    assert float(torch.linalg.norm(vo.T @ torch.from_numpy(w))) < 5e-3, log
    assert float(torch.linalg.eigvalsh(co).min()) > 0 and float(torch.linalg.norm(ko)) < 1.0
    with pytest.raises(NotImplementedError):
        g.optimize_reconstruction_parameters_nested_spd(x, y, torch.from_numpy(w), g.ConjugateGradient(),
                                                        cost_function=lambda *a: 0.0)


def test_nested_spd_eigenvalue_constraints(monkeypatch):
    """nested_spd_constraints_utils.py:13-97: eigenvalue constraints of the ambient matrix evaluated on the latent one.
    Values against the oracle's reconstruction (pinned on the reference) + LAPACK, and against the reference's own functions
    when the tree is mounted; gradients against central differences; the batched evaluation of
    ``batched_constraints`` (one autograd pass over all restarts) equals the one-point-at-a-time route."""
    import functools
    from gabotorch_b200 import _lib, manifold_optimization as mo, nested_mappings as nmap, ops
    from oracle import nested as onest, reference_loader, spd as ospd
    nopt = _nested_spd_on_cpu(monkeypatch)
    monkeypatch.setattr(nmap.ops, 'sym_eig', _cpu_sym_eig)
    monkeypatch.setattr(ops, 'spd_op', _OracleOps(None).spd_op)
    import gabotorch_b200.kernel_utils as ku
    monkeypatch.setattr(ku, '_dev64_keep_grad', lambda x: torch.as_tensor(x, dtype=torch.float64))
    rng = np.random.default_rng(17)
    D, d, R = 7, 3, 6
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    w, v = q[:, :d].copy(), q[:, d:].copy()
    c = ospd.spd_sample(rng, 1, D - d, max_cond=20.0)[0]
    k = rng.standard_normal((d, D - d))
    k = 0.6 * k / np.linalg.norm(k)
    y = ospd.spd_sample(rng, R, d, max_cond=50.0)
    args = tuple(torch.from_numpy(t) for t in (w, v, c, k))
    xa = onest.projection_from_nested_spd_to_spd(y, w, v, c, k).numpy()
    lam = np.linalg.eigvalsh(xa)
    fmax = g.max_eigenvalue_nested_spd_constraint(torch.from_numpy(y), 4.0, *args)
    fmin = g.min_eigenvalue_nested_spd_constraint(torch.from_numpy(y), 0.05, *args)
    np.testing.assert_allclose(fmax.numpy(), 4.0 - lam[:, -1], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fmin.numpy(), lam[:, 0] - 0.05, rtol=0, atol=1e-12)
    one = g.max_eigenvalue_nested_spd_constraint(torch.from_numpy(y[2]), 4.0, *args)
    assert one.dim() == 0 and abs(float(one) - float(fmax[2])) < 1e-13
    if reference_loader.available():
        reference_loader.load()
        import importlib
        ref = importlib.import_module('BoManifolds.nested_mappings.nested_spd_constraints_utils')
        for i in range(R):
            r1 = ref.max_eigenvalue_nested_spd_constraint(torch.from_numpy(y[i]), 4.0, *args)
            r2 = ref.min_eigenvalue_nested_spd_constraint(torch.from_numpy(y[i]), 0.05, *args)
            assert abs(float(r1) - float(fmax[i])) < 1e-10 and abs(float(r2) - float(fmin[i])) < 1e-10
    # gradient with respect to the latent matrix
    yt = torch.from_numpy(y[1]).clone().requires_grad_(True)
    g.max_eigenvalue_nested_spd_constraint(yt, 4.0, *args).backward()
    direction = rng.standard_normal((d, d)); direction = torch.from_numpy(0.5 * (direction + direction.T))
    h = 1e-6
    fd = (float(g.max_eigenvalue_nested_spd_constraint(torch.from_numpy(y[1]) + h * direction, 4.0, *args))
          - float(g.max_eigenvalue_nested_spd_constraint(torch.from_numpy(y[1]) - h * direction, 4.0, *args))) / (2 * h)
    assert abs(fd - float((yt.grad * direction).sum())) < 1e-6 * max(1.0, abs(fd))
    # batched evaluation (supports_batch) == one point at a time
    cons = [functools.partial(g.max_eigenvalue_nested_spd_constraint, maximum_eigenvalue=4.0, projection_matrix=args[0],
                              projection_complement_matrix=args[1], bottom_spd_matrix=args[2], contraction_matrix=args[3])]
    per_point = [lambda x: cons[0](x)]                       # a plain callable: the reference's one-at-a-time route
    fb, gb = mo.batched_constraints(cons, _lib.SPD)(torch.from_numpy(y))
    fp, gp_ = mo.batched_constraints(per_point, _lib.SPD)(torch.from_numpy(y))
    np.testing.assert_allclose(fb.numpy(), fp.numpy(), rtol=0, atol=1e-13)
    np.testing.assert_allclose(gb[0].numpy(), gp_[0].numpy(), rtol=0, atol=1e-12)
    # random latent points: a projected ambient sample, numpy like the reference
    monkeypatch.setattr(nmap, 'projection_from_spd_to_nested_spd',
                        lambda x, p: torch.as_tensor(p).T @ torch.as_tensor(x) @ torch.as_tensor(p))
    sample = g.random_nested_spd_with_spd_eigenvalue_constraints(None, lambda: xa[0], args[0])
    assert isinstance(sample, np.ndarray) and np.abs(sample - w.T @ xa[0] @ w).max() < 1e-12


def test_numpy_host_helpers_of_spd_utils():
    """symmetric_matrix_to_vector_mandel / vector_to_symmetric_matrix_mandel / spd_sample (spd_utils.py:57-101, 290-306): the
    numpy helpers the examples import, against the oracle's (reference-pinned) Mandel functions and the sampling law."""
    import types
    from gabotorch_b200 import riemannian_utils as ru
    from oracle import spd as ospd
    rng = np.random.default_rng(2)
    for d in (1, 2, 3, 5, 8):
        a = rng.standard_normal((d, d)); m = a + a.T
        v = ru.symmetric_matrix_to_vector_mandel(m)
        want = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(m[None]))[0].numpy()
        np.testing.assert_allclose(v, want, rtol=0, atol=1e-15)
        np.testing.assert_allclose(ru.vector_to_symmetric_matrix_mandel(v), m, rtol=0, atol=1e-15)
    man = types.SimpleNamespace(_n=4, min_eig=0.5, max_eig=3.0)
    np.random.seed(3)
    x = ru.spd_sample(man)
    lam = np.linalg.eigvalsh(x)
    assert x.shape == (4, 4) and np.abs(x - x.T).max() < 1e-14 and lam.min() >= 0.5 - 1e-12 and lam.max() <= 3.0 + 1e-12
    man.rand = types.MethodType(ru.spd_sample, man)           # the binding of gabo_spd.py:102
    assert man.rand().shape == (4, 4)


def test_tensor_gp_expected_improvement_host_logic(monkeypatch):
    """ops.TensorGP (EI of the log-Euclidean latent GP of hd_gabo_spd.py as differentiable tensor code): values against a
    plain fp64 restatement of botorch's analytic EI, Riemannian gradient against central differences.  Host run with the
    logm kernels replaced by an eigh-based stand-in; the device run is tests/test_lockstep_gpu.py."""
    from gabotorch_b200 import _lib, kernel_utils as ku, ops
    from oracle import spd as ospd

    class _Logm:
        @staticmethod
        def apply(m):
            lam, q = torch.linalg.eigh(m)
            return (q * torch.log(lam).unsqueeze(-2)) @ q.transpose(-1, -2)
    monkeypatch.setattr(ku, '_SpdLogm', _Logm)
    monkeypatch.setattr(ops, 'to_dev64', lambda x: torch.as_tensor(x, dtype=torch.float64))
    monkeypatch.setattr(ops, 'spd_op', _OracleOps(None).spd_op)
    rng = np.random.default_rng(6)
    d, n, r = 3, 9, 4
    xt = ospd.spd_sample(rng, n, d, max_cond=30.0)
    y = rng.standard_normal(n)
    xs = ospd.spd_sample(rng, r, d, max_cond=30.0)
    scale, ls, noise, mean = 1.3, 1.7, 0.05, 0.1

    def logm(m):
        lam, q = np.linalg.eigh(m)
        return (q * np.log(lam)) @ q.T
    st = np.array([logm(m) for m in xt])
    kk = lambda a, b: scale * math.exp(-np.sum((a - b + 1e-15) ** 2) / ls ** 2)   # noqa: E731
    K = np.array([[kk(a, b) for b in st] for a in st]) + noise * np.eye(n)
    kinv = np.linalg.inv(K)
    gp = ops.TensorGP(d, torch.from_numpy(st), torch.from_numpy(kinv @ (y - mean)), torch.from_numpy(kinv), mean, scale,
                      1.0 / ls ** 2, float(y.min()), use_log=True)

    def ei_ref(x):
        kx = np.array([kk(logm(x), b) for b in st])
        mu = mean + kx @ kinv @ (y - mean)
        s = math.sqrt(max(scale - kx @ kinv @ kx, 1e-9))
        u = (y.min() - mu) / s
        return s * (math.exp(-0.5 * u * u) / math.sqrt(2 * math.pi) + u * 0.5 * math.erfc(-u / math.sqrt(2)))
    ei, grad = ops.ei_eval(gp, torch.from_numpy(xs), want_grad=True)
    np.testing.assert_allclose(ei.numpy(), [ei_ref(x) for x in xs], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(ops.ei_eval(gp, torch.from_numpy(xs)).numpy(), ei.numpy(), rtol=0, atol=0)
    xi = rng.standard_normal((d, d)); xi = 0.5 * (xi + xi.T)
    for i in range(r):
        fd = (ei_ref(xs[i] + 1e-6 * xi) - ei_ref(xs[i] - 1e-6 * xi)) / 2e-6
        xinv = np.linalg.inv(xs[i])
        assert abs(fd - np.trace(xinv @ grad[i].numpy() @ xinv @ xi)) <= 1e-5 * abs(fd) + 1e-10
    assert gp.manifold == _lib.SPD and gp.with_compute(_lib.GABO_F64) is gp and gp.point_shape == (d, d)
