"""Device tests of the constrained trust-region / ALM solvers and of the two example BO loops (green on the B200 since
the round-1 suite; no xfail marks: a regression must fail).  The host logic they exercise is also pinned on CPU against
the reference's own solver classes (tests/test_host_logic.py, tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from gabotorch_b200 import _lib
from oracle import gp as ogp
from test_acq_gpu import device_gp

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['ctr_spd2_active', 'ctr_spd3'])
def test_lockstep_constrained_trust_regions_on_device_vs_reference_solver(golden, name):
    # gabo_spd.py's configuration: ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100), approx_hessian, one
    # max-eigenvalue constraint; golden arrays from the reference's own class (tests/golden/make_golden.py)
    import functools
    from gabotorch_b200 import manifold_optimization as mo, riemannian_utils as ru
    beta, noise, max_eig = (float(v) for v in golden[name + '_hyper'])
    gp = ogp.make_gp('spd', golden[name + '_xtrain'], golden[name + '_y'], beta=beta, noise=noise)
    cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=max_eig)]
    X, val, iters, _ = mo.batched_trust_regions(device_gp(gp, _lib.GABO_F64), golden[name + '_x0'], maxiter=100,
                                                mingradnorm=1e-4,
                                                ineq_constraints=mo.batched_constraints(cons, _lib.SPD))
    same = iters.cpu().numpy() == golden[name + '_iters']
    assert same.mean() >= 0.7
    np.testing.assert_allclose(X.cpu().numpy()[same], golden[name + '_x'][same], rtol=0, atol=1e-5)
    np.testing.assert_allclose(-val.cpu().numpy()[same], golden[name + '_cost'][same], rtol=1e-5, atol=1e-9)


def _domain_constraint(angle):
    def constraint(x):
        centre = torch.zeros(3, dtype=x.dtype, device=x.device)
        centre[0] = 1
        in_prod = torch.mm(x[None], centre[:, None])
        one = torch.ones(1, dtype=x.dtype, device=x.device)
        in_prod = torch.max(torch.min(in_prod, one), -one)
        return angle - torch.acos(in_prod)[0, 0]
    return constraint


def test_lockstep_constrained_trust_regions_on_the_sphere_on_device(golden):
    # ConstrainedTrustRegions(maxiter=200) with the domain constraint of gabo_sphere_inequality_constraints.py as a
    # user-supplied torch callable (autograd per restart on the device); golden: the reference's own class
    from gabotorch_b200 import manifold_optimization as mo
    name = 'ctr_s2_domain'
    beta, noise, angle = (float(v) for v in golden[name + '_hyper'])
    gp = ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=beta, noise=noise)
    cons = mo.batched_constraints([_domain_constraint(angle)], _lib.SPHERE)
    X, val, iters, _ = mo.batched_trust_regions(device_gp(gp, _lib.GABO_F64), golden[name + '_x0'], maxiter=200,
                                                ineq_constraints=cons)
    same = iters.cpu().numpy() == golden[name + '_iters']
    assert same.mean() >= 0.7
    np.testing.assert_allclose(X.cpu().numpy()[same], golden[name + '_x'][same], rtol=0, atol=1e-5)


def test_lockstep_augmented_lagrangian_on_device(golden):
    # AugmentedLagrangeMethod(inner_solver=TrustRegions(maxiter=50), maxiter=12) against the oracle restatement of the
    # reference's class (oracle/alm.py, itself pinned on the class through the alm_* golden arrays)
    from gabotorch_b200 import manifold_optimization as mo
    from oracle import alm as oalm, ctr as octr
    name = 'alm_s2_domain'
    beta, noise, angle = (float(v) for v in golden[name + '_hyper'])
    gp = ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=beta, noise=noise)
    cons = mo.batched_constraints([_domain_constraint(angle)], _lib.SPHERE)
    X, _, iters, _ = mo.batched_alm(device_gp(gp, _lib.GABO_F64), golden[name + '_x0'], dict(maxiter=50),
                                    ineq_constraints=cons, maxiter=12)
    for i, x0 in enumerate(golden[name + '_x0']):
        xo, ko = oalm.solve_alm(gp, x0, ineq_constraints=[octr.sphere_domain_constraint([1.0, 0.0, 0.0], angle)],
                                maxiter=12, inner_opts={'maxiter': 50})
        assert int(iters[i]) == ko
        np.testing.assert_allclose(X[i].cpu().numpy(), xo, rtol=0, atol=1e-5)


def _load_example(name):
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(name, os.path.join(root, 'examples', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_gabo_sphere_example_runs_on_device():
    # examples/gabo_sphere.py: GP fit + EI + TrustRegions (gabo_acq_rtr) per iteration; host wiring pinned on CPU by
    # tests/test_example_emulated.py
    from oracle import sphere as osph
    x, y, best = _load_example('gabo_sphere').run(dim=3, n_iters=3, num_restarts=5, raw_samples=100, seed=11,
                                                  verbose=False)
    assert tuple(x.shape) == (8, 3) and len(best) == 4
    np.testing.assert_allclose(x.norm(dim=-1).numpy(), 1.0, atol=1e-12)
    np.testing.assert_allclose(y.numpy(), osph.ackley(x.numpy()), rtol=1e-9)
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:]))


def test_gabo_spd_example_runs_on_device():
    # examples/gabo_spd.py: ConstrainedTrustRegions + max-eigenvalue constraint through the lock-step driver
    from oracle import spd as ospd
    x, y, best = _load_example('gabo_spd').run(dim=2, n_iters=2, num_restarts=5, raw_samples=50, seed=11, verbose=False)
    assert tuple(x.shape) == (7, 3) and len(best) == 3
    assert np.linalg.eigvalsh(ospd.vector_to_symmetric_matrix_mandel(x).numpy()).min() > 0
    np.testing.assert_allclose(y.numpy(), ospd.ackley(x), rtol=1e-8)
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:]))


def test_strict_constrained_trust_regions_with_nested_spd_constraints_on_device(golden):
    # the acquisition step of hd_gabo_spd.py (:242-268): StrictConstrainedTrustRegions(mingradnorm=2e-4, maxiter=100) on the
    # LATENT manifold SPD(3) with the eigenvalue constraints of the AMBIENT matrix (nested_spd_constraints_utils.py:13-78,
    # reconstructed with projection_from_nested_spd_to_spd).  The constraints are batch-capable callables: one
    # reconstruction + one gabo_sym_eig launch + one autograd pass for all restarts.  The strict variant never leaves the
    # feasible set; the iterates end feasible with an acquisition value no worse than at the feasible starts.
    import functools
    import gabotorch_b200 as g
    from gabotorch_b200 import manifold_optimization as mo
    from oracle import nested as onest, spd as ospd
    name = 'ctr_spd3'
    beta, noise, _ = (float(v) for v in golden[name + '_hyper'])
    gp = ogp.make_gp('spd', golden[name + '_xtrain'], golden[name + '_y'], beta=beta, noise=noise)
    dgp = device_gp(gp, _lib.GABO_F64)
    rng = np.random.default_rng(4)
    D, d = 6, 3
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    w, v = q[:, :d].copy(), q[:, d:].copy()
    c = np.diag([0.8, 1.0, 1.2])
    k = rng.standard_normal((d, D - d))
    k = 0.4 * k / np.linalg.norm(k)
    args = dict(projection_matrix=torch.from_numpy(w), projection_complement_matrix=torch.from_numpy(v),
                bottom_spd_matrix=torch.from_numpy(c), contraction_matrix=torch.from_numpy(k))
    x0 = golden[name + '_x0']
    amb = np.linalg.eigvalsh(onest.projection_from_nested_spd_to_spd(x0, w, v, c, k).numpy())
    max_eig, min_eig = float(amb[:, -1].max() * 1.05), float(amb[:, 0].min() * 0.5)      # every start strictly feasible
    cons = [functools.partial(g.max_eigenvalue_nested_spd_constraint, maximum_eigenvalue=max_eig, **args),
            functools.partial(g.min_eigenvalue_nested_spd_constraint, minimum_eigenvalue=min_eig, **args)]
    batched = mo.batched_constraints(cons, _lib.SPD)
    f0, _ = batched(torch.from_numpy(x0).cuda())
    assert float(f0.min()) > 0
    # the device values equal the oracle composition (reconstruction pinned on the reference + LAPACK)
    np.testing.assert_allclose(f0.cpu().numpy(), np.stack([max_eig - amb[:, -1], amb[:, 0] - min_eig], -1), rtol=0, atol=1e-10)
    from gabotorch_b200 import ops
    ei0 = ops.ei_eval(dgp, torch.from_numpy(x0).cuda())
    X, val, iters, _ = mo.batched_trust_regions(dgp, x0, maxiter=100, mingradnorm=2e-4, ineq_constraints=batched,
                                                strict=True)
    f1, _ = batched(X)
    assert float(f1.min()) >= -1e-6, f1.min()
    assert bool((val >= ei0 - 1e-9).all())
    assert np.linalg.eigvalsh(X.cpu().numpy()).min() > 0


def test_hd_gabo_spd_example_loop_on_device():
    # the whole HD-GaBO iteration of hd_gabo_spd.py on the device: manifold GP fit (Grassmann projection), latent
    # log-Euclidean GP, reconstruction fit, strict constrained trust regions on the latent manifold with ambient eigenvalue
    # constraints, reconstruction of the candidate
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('hd_gabo_spd_example', os.path.join(root, 'examples', 'hd_gabo_spd.py'))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    x, y, best = ex.run(dim=5, latent_dim=2, n_iters=3, num_restarts=3, raw_samples=40, nb_data_init=5, seed=3, verbose=False)
    assert tuple(x.shape) == (8, 15) and tuple(y.shape) == (8,) and len(best) == 4
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and best[-1] == float(y.min())
    from gabotorch_b200 import riemannian_utils as ru
    lam = torch.linalg.eigvalsh(ru.vector_to_symmetric_matrix_mandel_torch(x).cpu())
    assert float(lam.min()) > 0 and float(lam[5:].max()) <= 5.0 + 1e-6          # candidates respect the ambient eigenvalue bound


def test_tensor_gp_expected_improvement_vs_oracle_formulas():
    # EI of the log-Euclidean latent GP (ops.TensorGP: device tensor code + autograd) against a plain fp64 restatement
    # (kernel exp(-||logm X - logm X_i + 1e-15||_F^2 / l^2), botorch analytic EI), gradient against central differences
    import math
    import gabotorch_b200 as g
    from gabotorch_b200 import ops, riemannian_utils as ru
    from oracle import spd as ospd
    rng = np.random.default_rng(6)
    d, n, r = 3, 9, 5
    xt = ospd.spd_sample(rng, n, d, max_cond=30.0)
    y = rng.standard_normal(n)
    xs = ospd.spd_sample(rng, r, d, max_cond=30.0)
    cov = g.ScaleKernel(g.SpdLogEuclideanGaussianKernel())
    cov.outputscale = 1.3
    cov.base_kernel.lengthscale = 1.7
    xt_vec = ru.symmetric_matrix_to_vector_mandel_torch(torch.from_numpy(xt)).cpu()
    model = g.ManifoldGP(xt_vec, torch.from_numpy(y), cov, noise=0.05, mean=0.1)
    acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
    gp = acq.device_gp()
    assert getattr(gp, 'is_tensor_gp', False)
    ei, grad = ops.ei_eval(gp, torch.from_numpy(xs), want_grad=True)

    def logm(m):
        lam, q = np.linalg.eigh(m)
        return (q * np.log(lam)) @ q.T

    def ei_ref(x):
        st = np.array([logm(m) for m in xt])
        kk = lambda a, b: 1.3 * math.exp(-np.sum((a - b + 1e-15) ** 2) / 1.7 ** 2)
        K = np.array([[kk(a, b) for b in st] for a in st]) + 0.05 * np.eye(n)
        kx = np.array([kk(logm(x), b) for b in st])
        mu = 0.1 + kx @ np.linalg.solve(K, y - 0.1)
        var = max(1.3 - kx @ np.linalg.solve(K, kx), 1e-9)
        s = math.sqrt(var)
        u = (y.min() - mu) / s
        return s * (math.exp(-0.5 * u * u) / math.sqrt(2 * math.pi) + u * 0.5 * math.erfc(-u / math.sqrt(2)))
    want = np.array([ei_ref(x) for x in xs])
    # the training Gram comes from gabo_frobenius_gram (exp on the fp32 MUFU path, ~1e-7): 1e-5 like every kernel value
    np.testing.assert_allclose(ei.cpu().numpy(), want, rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(acq(ru.symmetric_matrix_to_vector_mandel_torch(torch.from_numpy(xs)).cpu()[:, None]).numpy(),
                               want, rtol=1e-5, atol=1e-12)
    # Riemannian gradient X sym(E) X: <grad, xi>_X = tr(X^-1 grad X^-1 xi) = dEI(xi)
    xi = rng.standard_normal((d, d)); xi = 0.5 * (xi + xi.T)
    h = 1e-6
    for i in range(r):
        fd = (ei_ref(xs[i] + h * xi) - ei_ref(xs[i] - h * xi)) / (2 * h)
        xinv = np.linalg.inv(xs[i])
        an = np.trace(xinv @ grad[i].cpu().numpy() @ xinv @ xi)
        assert abs(fd - an) <= 1e-4 * abs(fd) + 1e-9, (i, fd, an)


def test_hd_gabo_sphere_example_loop_on_device():
    # the whole HD-GaBO iteration of hd_gabo_sphere.py on the device: manifold GP fit (sphere-valued axes), latent sphere GP,
    # reconstruction fit of the distances to axis, multi-start trust regions on the latent sphere, lift of the candidate
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('hd_gabo_sphere_example', os.path.join(root, 'examples', 'hd_gabo_sphere.py'))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    x, y, best = ex.run(dim=5, latent_dim=3, n_iters=3, num_restarts=3, raw_samples=40, nb_data_init=5, seed=5, verbose=False)
    assert tuple(x.shape) == (8, 5) and tuple(y.shape) == (8,) and len(best) == 4
    np.testing.assert_allclose(x.norm(dim=-1).numpy(), 1.0, atol=1e-6)      # the fitted distances are float32, as in the reference
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and best[-1] == float(y.min())


@pytest.mark.parametrize('constraint,solver', [('bound', 'ctr'), ('inequality', 'ctr'), ('inequality', 'alm'),
                                               ('equality', 'ctr'), ('equality', 'alm')])
def test_constrained_sphere_example_loops_on_device(constraint, solver):
    # the three constrained examples of the reference (bo_sphere/constrained_benchmark_examples/): user-supplied torch
    # callables as constraints, the feasible sampler re-bound on the manifold object, both solver choices; the proposed
    # candidates stay on the sphere and inside the domain (to the solvers' tolerance)
    import importlib.util, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('gabo_sphere_constrained_example',
                                                  os.path.join(root, 'examples', 'gabo_sphere_constrained.py'))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    x, y, best, worst = ex.run(constraint, solver, n_iters=3, num_restarts=3, raw_samples=40, seed=11, verbose=False)
    assert tuple(x.shape) == (8, 3) and len(best) == 4
    np.testing.assert_allclose(x.norm(dim=-1).numpy(), 1.0, atol=1e-9)
    assert worst > -1e-3, worst
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:]))
