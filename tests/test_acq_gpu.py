"""Acquisition path (A1-A4) through the C ABI against the oracle: EI values and closed-form Riemannian gradients,
multi-start conjugate gradient trajectories (fp64 arithmetic follows the oracle step for step), fp32 solves reach the
same optima, candidate selection is bit-exact.  The third-party parts of the oracle (pymanopt CG, botorch EI) are
restated from the published algorithms: parity unpinned, see oracle/__init__.py."""
import math

import numpy as np
import pytest
import torch

from gabotorch_b200 import _lib, ops
from oracle import gp as ogp
from oracle import rcg as orcg
from oracle import spd as ospd
from oracle import sphere as osph

pytestmark = pytest.mark.gpu


def sphere_problem(D, n, beta, noise, seed):
    rng = np.random.default_rng(seed)
    xt = osph.rand(rng, n, D)
    gp = ogp.make_gp('sphere', xt, osph.ackley(xt), beta=beta, noise=noise)
    return rng, gp


def spd_problem(d, n, beta, noise, seed):
    rng = np.random.default_rng(seed)
    xt = ospd.spd_sample(rng, n, d, max_cond=100.0)
    y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt)))
    gp = ogp.make_gp('spd', xt, y, beta=beta, noise=noise)
    return rng, gp


def device_gp(gp, compute):
    if gp.manifold == 'sphere':
        xdev = ops.to_dev64(gp.x_train)
        return ops.DeviceGP(_lib.SPHERE, gp.x_train.shape[-1], xdev, gp.alpha, gp.minv, gp.mean, gp.outputscale,
                            gp.beta, gp.best_f, gp.kxx, compute)
    d = gp.x_train.shape[-1]
    fac = ops.spd_factor(gp.x_train, d, False)
    return ops.DeviceGP(_lib.SPD, d, fac, gp.alpha, gp.minv, gp.mean, gp.outputscale, gp.beta, gp.best_f, gp.kxx,
                        compute)


@pytest.mark.parametrize('D,n', [(3, 5), (6, 32), (4, 33), (9, 100), (16, 128)])
def test_sphere_ei_and_gradient(D, n):
    rng, gp = sphere_problem(D, n, beta=1.0 + math.log(2.0), noise=1e-2, seed=D)
    x = osph.rand(rng, 64, D)
    x[0] = gp.x_train[0]                                      # on a training point (distance clamp path)
    ref = [ogp.ei_and_grad(gp, xi) for xi in x]
    ei_ref = np.array([r[0] for r in ref])
    g_ref = np.array([r[1] for r in ref])
    ei64, g64 = ops.ei_eval(device_gp(gp, _lib.GABO_F64), x, want_grad=True)
    np.testing.assert_allclose(ei64.cpu().numpy(), ei_ref, rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(g64.cpu().numpy(), g_ref, rtol=1e-7, atol=1e-10 * max(1.0, np.abs(g_ref).max()))
    ei32, g32 = ops.ei_eval(device_gp(gp, _lib.GABO_F32), x, want_grad=True)
    np.testing.assert_allclose(ei32.cpu().numpy(), ei_ref, rtol=2e-4, atol=1e-6 * ei_ref.max())
    np.testing.assert_allclose(g32.cpu().numpy(), g_ref, rtol=0, atol=5e-4 * np.abs(g_ref).max())


@pytest.mark.parametrize('d,n', [(2, 7), (3, 32), (5, 20), (8, 32), (3, 70)])
def test_spd_ei_and_gradient(d, n):
    rng, gp = spd_problem(d, n, beta=0.3 + math.log(2.0), noise=1e-2, seed=10 + d)
    x = ospd.spd_sample(rng, 40, d, max_cond=100.0)
    x[0] = gp.x_train[0]
    ref = [ogp.ei_and_grad(gp, xi) for xi in x]
    ei_ref = np.array([r[0] for r in ref])
    g_ref = np.array([r[1] for r in ref])
    ei64, g64 = ops.ei_eval(device_gp(gp, _lib.GABO_F64), x, want_grad=True)
    np.testing.assert_allclose(ei64.cpu().numpy(), ei_ref, rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(g64.cpu().numpy(), g_ref, rtol=1e-6, atol=1e-9 * max(1.0, np.abs(g_ref).max()))
    ei32, g32 = ops.ei_eval(device_gp(gp, _lib.GABO_F32), x, want_grad=True)
    np.testing.assert_allclose(ei32.cpu().numpy(), ei_ref, rtol=1e-3, atol=1e-5 * ei_ref.max())
    np.testing.assert_allclose(g32.cpu().numpy(), g_ref, rtol=0, atol=2e-3 * np.abs(g_ref).max())
    bad = x.copy()
    bad[5] = -bad[5]
    ei_bad = ops.ei_eval(device_gp(gp, _lib.GABO_F64), bad).cpu().numpy()
    assert np.isnan(ei_bad[5]) and not np.isnan(np.delete(ei_bad, 5)).any()


@pytest.mark.parametrize('manifold,dim,n', [('sphere', 6, 32), ('sphere', 3, 12), ('spd', 3, 16), ('spd', 8, 32),
                                            ('spd', 2, 40),
                                            # kernel variants: padded dimension 16 with two / four points per lane, the
                                            # shared-memory fallback (dim > 8 with n > 64), odd SPD sizes (dummy column
                                            # of the looped Jacobi, warp-cooperative direction solve), n > 32 on SPD
                                            ('sphere', 9, 40), ('sphere', 4, 70), ('sphere', 12, 100), ('spd', 5, 20),
                                            ('spd', 7, 16), ('spd', 4, 40)])
def test_rcg_f64_follows_the_oracle_step_for_step(manifold, dim, n):
    steps = 25
    if manifold == 'sphere':
        rng, gp = sphere_problem(dim, n, beta=1.0 + math.log(2.0), noise=1e-2, seed=77 + dim)
        x0 = osph.rand(rng, 12, dim)
    else:
        rng, gp = spd_problem(dim, n, beta=0.3 + math.log(2.0), noise=1e-2, seed=88 + dim)
        x0 = ospd.spd_sample(rng, 12, dim, min_eig=0.5, max_eig=3.0)
    opts = orcg.CGOptions(maxiter=steps)
    ref = [orcg.solve_cg(gp, xi, opts) for xi in x0]
    cand, val, iters, reason = ops.acq_rcg(device_gp(gp, _lib.GABO_F64), x0, maxiter=steps)
    cand, val, iters, reason = cand.cpu().numpy(), val.cpu().numpy(), iters.cpu().numpy(), reason.cpu().numpy()
    agree = 0
    for i, (x, cost, it, why) in enumerate(ref):
        same_path = (it == iters[i]) and (why == reason[i])
        if same_path and np.allclose(cand[i], x, rtol=1e-6, atol=1e-8):
            agree += 1
            assert abs(val[i] + cost) <= 1e-7 * max(abs(cost), 1e-12) + 1e-13
        # in every case the solve must not end below the oracle's optimum by more than rounding noise
        assert val[i] >= -cost * (1 - 1e-3) - 1e-12 or val[i] >= ogp.ei_and_grad(gp, x0[i], False)[0]
    assert agree >= len(ref) - 2, 'only %d of %d restarts follow the oracle trajectory' % (agree, len(ref))


@pytest.mark.parametrize('manifold,dim,n,R', [('sphere', 6, 32, 256), ('spd', 3, 24, 128), ('spd', 8, 32, 64)])
def test_rcg_f32_improves_ei_and_stays_on_the_manifold(manifold, dim, n, R):
    if manifold == 'sphere':
        rng, gp = sphere_problem(dim, n, beta=1.0 + math.log(2.0), noise=1e-2, seed=5)
        x0 = osph.rand(rng, R, dim)
    else:
        rng, gp = spd_problem(dim, n, beta=0.3 + math.log(2.0), noise=1e-2, seed=6)
        x0 = ospd.spd_sample(rng, R, dim, min_eig=0.5, max_eig=3.0)
    dgp = device_gp(gp, _lib.GABO_F32)
    ei0 = ops.ei_eval(dgp, x0).cpu().numpy()
    cand, val, iters, reason = ops.acq_rcg(dgp, x0, maxiter=60)
    cand, val = cand.cpu().numpy(), val.cpu().numpy()
    assert np.all(val >= ei0 * (1 - 1e-5) - 1e-12)                 # line search never accepts an increase of the cost
    assert np.all(np.isin(reason.cpu().numpy(), [1, 2, 3]))
    # the reported value is the acquisition at the returned candidate (manifold_optimize.py:227), checked by the oracle
    ref_val = np.array([ogp.ei_and_grad(gp, c, want_grad=False)[0] for c in cand[:32]])
    np.testing.assert_allclose(val[:32], ref_val, rtol=2e-3, atol=1e-5 * ref_val.max())
    if manifold == 'sphere':
        np.testing.assert_allclose(np.linalg.norm(cand, axis=-1), 1.0, atol=1e-12)
    else:
        assert np.abs(cand - np.swapaxes(cand, -1, -2)).max() == 0.0
        assert np.linalg.eigvalsh(cand).min() > 0
    # the fp32 solves find the same best acquisition value as the fp64 solves (same algorithm, different rounding)
    _, val64, _, _ = ops.acq_rcg(device_gp(gp, _lib.GABO_F64), x0, maxiter=60)
    assert abs(val.max() - float(val64.max())) <= 2e-2 * float(val64.max())


def test_rcg_stopping_criteria():
    rng, gp = sphere_problem(4, 10, beta=2.0, noise=1e-2, seed=1)
    x0 = osph.rand(rng, 8, 4)
    dgp = device_gp(gp, _lib.GABO_F64)
    _, _, it1, why1 = ops.acq_rcg(dgp, x0, maxiter=1)
    assert torch.all(it1 == 0) and torch.all(why1 == 1)            # pymanopt: stops when iter + 1 >= maxiter
    _, _, it2, why2 = ops.acq_rcg(dgp, x0, maxiter=1000, mingradnorm=1e30)
    assert torch.all(it2 == 0) and torch.all(why2 == 2)
    _, _, it3, why3 = ops.acq_rcg(dgp, x0, maxiter=1000)
    assert torch.all(why3 >= 2) and int(it3.max()) < 999           # converges long before maxiter


@pytest.mark.parametrize('compute', [_lib.GABO_F32, _lib.GABO_F64])
def test_sphere_rcg_speculative_line_search_is_exactly_the_sequential_one(compute):
    # The launcher picks the speculation width from the number of restarts (4 warps per restart up to 148 restarts,
    # 2 up to 1216, plain sequential search beyond).  Speculation evaluates several trial steps of pymanopt's
    # backtracking at once and keeps the first acceptable one, so iterates, values and iteration counts must be
    # BIT-IDENTICAL to the sequential search whatever the width.
    rng, gp = sphere_problem(6, 32, beta=1.0 + math.log(2.0), noise=1e-2, seed=77)
    x0 = osph.rand(rng, 1300, 6)
    dgp = device_gp(gp, compute)
    seq = ops.acq_rcg(dgp, x0, maxiter=40)               # width 1
    two = ops.acq_rcg(dgp, x0[:600], maxiter=40)         # width 2
    four = ops.acq_rcg(dgp, x0[:100], maxiter=40)        # width 4
    for part, m in ((two, 600), (four, 100)):
        assert torch.equal(part[0], seq[0][:m])          # candidates
        assert torch.equal(part[1], seq[1][:m])          # EI values
        assert torch.equal(part[2], seq[2][:m])          # iterations
        assert torch.equal(part[3], seq[3][:m])          # stopping reasons
    assert int(seq[2].max()) > 3


@pytest.mark.parametrize('d', [3, 8])
def test_spd_rcg_speculative_line_search_is_exactly_the_sequential_one(d, monkeypatch):
    # same property on SPD(d): the CTA-per-restart kernel with 1 or 2 warps per restart (GABO_ACQ_SPEC forces the
    # width, the launcher otherwise picks it from the restart count and the occupancy) must give bit-identical solves
    rng, gp = spd_problem(d, 24, beta=0.3 + math.log(2.0), noise=1e-2, seed=90 + d)
    x0 = ospd.spd_sample(rng, 48, d, max_cond=100.0)
    dgp = device_gp(gp, _lib.GABO_F32)
    runs = {}
    for width in ('1', '2', '4'):          # 4 exists for d <= 5 only; at d = 8 the request falls through to 2
        monkeypatch.setenv('GABO_ACQ_SPEC', width)
        runs[width] = ops.acq_rcg(dgp, x0, maxiter=15)
    monkeypatch.delenv('GABO_ACQ_SPEC')
    auto = ops.acq_rcg(dgp, x0, maxiter=15)
    for other in (runs['2'], runs['4'], auto):
        for a, b in zip(runs['1'], other):
            assert torch.equal(a, b)
    assert int(runs['1'][2].max()) > 2


def test_argmax_records_is_exact_and_sharding_invariant():
    rng = np.random.default_rng(0)
    v = rng.standard_normal(5000).astype(np.float32).astype(np.float64)
    v[[17, 901, 4000]] = v.max() + 1.0                            # three-way tie
    v[5] = np.nan
    slot, best = ops.argmax_records(v)
    assert int(slot) == orcg.best_candidate(v) == 17 and float(best) == v[17]
    # shards of any size, gathered in any order, select the same global index
    for world in (2, 4, 8):
        recs_v, recs_g = [], []
        for r in range(world):
            lo, hi = (r * 5000) // world, ((r + 1) * 5000) // world
            gidx = torch.arange(lo, hi)
            s, b = ops.argmax_records(v[lo:hi], gidx)
            recs_v.append(float(b))
            recs_g.append(int(gidx[int(s)]))
        order = rng.permutation(world)
        w, _ = ops.argmax_records(np.array(recs_v)[order], torch.tensor(np.array(recs_g)[order]))
        assert np.array(recs_g)[order][int(w)] == 17
    allnan = np.full(7, np.nan)
    assert int(ops.argmax_records(allnan)[0]) == 0


# ---- trust-region solver: pinned on the reference's own TrustRegions class -----------------------

def _gp_from_golden(golden, name):
    beta, noise = golden[name + '_hyper']
    return ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))


@pytest.mark.parametrize('name', ['rtr_s2', 'rtr_s5', 'rtr_s5_noisy'])
def test_rtr_f64_reproduces_the_reference_solver(golden, name):
    # tests/golden/make_golden.py ran robust_trust_regions.TrustRegions + approximate_hessian.get_hessianfd (the
    # reference's own code) from these starts: same iteration counts, same candidates, same costs
    gp = _gp_from_golden(golden, name)
    x, val, iters, reason = ops.acq_rtr(device_gp(gp, _lib.GABO_F64), golden[name + '_x0'])
    np.testing.assert_array_equal(iters.cpu().numpy(), golden[name + '_iters'])
    np.testing.assert_allclose(x.cpu().numpy(), golden[name + '_x'], rtol=0, atol=1e-7)
    np.testing.assert_allclose(-val.cpu().numpy(), golden[name + '_cost'], rtol=1e-8, atol=1e-12)
    assert (reason.cpu().numpy() == 2).all()                       # every solve ended on the gradient norm


@pytest.mark.parametrize('D,n,R', [(3, 5, 33), (6, 32, 256), (4, 40, 50), (8, 100, 40), (12, 64, 30)])
def test_rtr_vs_oracle_and_f32_reaches_the_same_optima(D, n, R):
    from oracle import rtr as ortr
    rng, gp = sphere_problem(D, n, beta=1.0 + math.log(2.0), noise=1e-2, seed=100 + D)
    x0 = osph.rand(rng, R, D)
    opts = ortr.TROptions(maxiter=60)
    ref_x, ref_v, ref_it = ortr.gen_candidates(gp, x0[:12], opts)
    x64, v64, it64, _ = ops.acq_rtr(device_gp(gp, _lib.GABO_F64), x0, maxiter=60)
    same = it64.cpu().numpy()[:12] == ref_it
    assert same.sum() >= 10                                        # a flipped accept/reject test changes the path
    np.testing.assert_allclose(x64.cpu().numpy()[:12][same], ref_x[same], rtol=0, atol=1e-6)
    np.testing.assert_allclose(v64.cpu().numpy()[:12], ref_v, rtol=1e-6, atol=1e-10)
    ei0 = np.array([ogp.ei_and_grad(gp, xi, want_grad=False)[0] for xi in x0])
    x32, v32, it32, _ = ops.acq_rtr(device_gp(gp, _lib.GABO_F32), x0, maxiter=60)
    for xs, vs in ((x64, v64), (x32, v32)):
        xs, vs = xs.cpu().numpy(), vs.cpu().numpy()
        np.testing.assert_allclose(np.linalg.norm(xs, axis=-1), 1.0, atol=1e-12)
        assert (vs >= ei0 - 1e-6 * max(1.0, ei0.max())).all()      # trust regions never accept an increase of the cost
        check = np.array([ogp.ei_and_grad(gp, xi, want_grad=False)[0] for xi in xs[:16]])
        np.testing.assert_allclose(vs[:16], check, rtol=2e-4, atol=1e-6 * max(check.max(), 1e-30))
    # fp32 and fp64 solves end in the same local optima for (almost) every start
    close = np.abs(v32.cpu().numpy() - v64.cpu().numpy()) <= 1e-3 * np.abs(v64.cpu().numpy()).max()
    assert close.mean() >= 0.9


def test_rtr_through_the_reference_api_and_argument_errors():
    import gabotorch_b200 as g
    rng = np.random.default_rng(3)
    xt = osph.rand(rng, 20, 3)
    y = osph.ackley(xt)
    model = g.ManifoldGP(xt, y, g.ScaleKernel(g.SphereGaussianKernel(beta_min=6.5)), noise=1e-2)
    ei = g.ExpectedImprovement(model, best_f=float(np.min(y)), maximize=False, compute='f64')
    x0 = torch.from_numpy(osph.rand(rng, 9, 3))[:, None, :]
    cand, vals = g.gen_candidates_manifold(x0, ei, g.Sphere(3), g.TrustRegions(), approx_hessian=True)
    assert tuple(cand.shape) == (9, 1, 3) and tuple(vals.shape) == (9,)
    np.testing.assert_allclose(ei(cand).numpy(), vals.numpy(), rtol=1e-9, atol=1e-14)
    assert (vals >= ei(x0) - 1e-12).all()
    best = g.joint_optimize_manifold(ei, g.Sphere(3), g.TrustRegions(maxiter=50), q=1, num_restarts=8, raw_samples=64,
                                     approx_hessian=True)
    assert tuple(best.shape) == (1, 3) and abs(float(best.norm()) - 1.0) < 1e-12
    with pytest.raises(NotImplementedError):
        g.TrustRegions(use_rand=True)
    with pytest.raises(ValueError):                                            # GP on the sphere, solver manifold SPD
        g.gen_candidates_manifold(x0, ei, g.PositiveDefinite(2), g.TrustRegions())
    rng2, gp = sphere_problem(20, 16, beta=1.0, noise=1e-2, seed=1)
    with pytest.raises(_lib.GaboError):
        ops.acq_rtr(device_gp(gp, _lib.GABO_F64), osph.rand(rng2, 4, 20))      # dimension beyond the register kernel


@pytest.mark.parametrize('manifold,dim,n,R', [('spd', 3, 16, 24), ('sphere', 20, 32, 16)])
def test_lockstep_trust_regions_on_device_vs_oracle(manifold, dim, n, R):
    # SPD(d) and large spheres: the reference's TrustRegions for all restarts in lock-step over the batched kernels
    # (gabo_ei_eval, gabo_spd_op, gabo_spd_scalar); the host logic alone is checked on CPU in tests/test_host_logic.py
    from gabotorch_b200 import manifold_optimization as mo
    from oracle import rtr as ortr
    if manifold == 'spd':
        rng, gp = spd_problem(dim, n, beta=0.5 + math.log(2.0), noise=1e-2, seed=21)
        x0 = ospd.spd_sample(rng, R, dim, max_cond=50.0)
    else:
        rng, gp = sphere_problem(dim, n, beta=0.35 + math.log(2.0), noise=1e-2, seed=22)
        x0 = osph.rand(rng, R, dim)
    dgp = device_gp(gp, _lib.GABO_F64)
    assert mo._rtr_kernel_covers(dgp) == (manifold == 'spd')      # SPD(d) also has the one-launch kernel (tested below)
    X, val, iters, reason = mo.batched_trust_regions(dgp, x0, maxiter=15)
    X, val, iters = X.cpu().numpy(), val.cpu().numpy(), iters.cpu().numpy()
    ei0 = np.array([ogp.ei_and_grad(gp, xi, want_grad=False)[0] for xi in x0])
    assert (val >= ei0 - 1e-9 * max(1.0, ei0.max())).all()                     # no accepted step increases the cost
    assert set(np.unique(reason.cpu().numpy())) <= {1, 2} and (iters >= 1).all() and (iters <= 15).all()
    opts = ortr.TROptions(maxiter=15)
    same = 0
    for i in range(6):
        xi, ci, ki = ortr.solve_tr(gp, x0[i], opts)
        check = ogp.ei_and_grad(gp, X[i], want_grad=False)[0]
        assert abs(val[i] - check) <= 1e-6 * max(abs(check), 1e-12)
        if ki == iters[i]:
            same += 1
            np.testing.assert_allclose(X[i], xi, rtol=0, atol=1e-5)
            assert abs(val[i] + ci) <= 1e-6 * max(abs(ci), 1e-12)
    assert same >= 4
    if manifold == 'spd':
        assert np.linalg.eigvalsh(0.5 * (X + np.swapaxes(X, -1, -2))).min() > 0
    else:
        np.testing.assert_allclose(np.linalg.norm(X, axis=-1), 1.0, atol=1e-12)



# ---- trust regions on SPD(d) in one launch (gabo_acq_rtr / gabo_acq_ctr): pinned on the reference's own classes ------

def _spd_gp_from_golden(golden, name):
    beta, noise, bound = (float(v) for v in golden[name + '_hyper'])
    return ogp.make_gp('spd', golden[name + '_xtrain'], golden[name + '_y'], beta=beta, noise=noise), bound


@pytest.mark.parametrize('name', ['ctr_spd2_active', 'ctr_spd2', 'ctr_spd3', 'sctr_spd2_active', 'sctr_spd3'])
def test_spd_constrained_trust_region_kernel_reproduces_the_reference_solver(golden, name):
    # tests/golden/make_golden.py ran the reference's OWN ConstrainedTrustRegions / StrictConstrainedTrustRegions classes
    # (constrained_trust_regions.py) in gabo_spd.py's configuration -- mingradnorm 1e-4 (2e-4 strict, hd_gabo_spd.py:194),
    # maxiter 100, finite-difference Hessian, one max-eigenvalue inequality constraint -- from these starts
    from oracle import ctr as octr
    gp, max_eig = _spd_gp_from_golden(golden, name)
    strict = name.startswith('sctr')
    x, val, iters, reason = ops.acq_ctr(device_gp(gp, _lib.GABO_F64), golden[name + '_x0'], [('max', max_eig)],
                                        strict=strict, maxiter=100, mingradnorm=2e-4 if strict else 1e-4)
    x, val, iters = x.cpu().numpy(), val.cpu().numpy(), iters.cpu().numpy()
    same = iters == golden[name + '_iters']
    assert same.mean() >= 0.8, (iters, golden[name + '_iters'])
    np.testing.assert_allclose(x[same], golden[name + '_x'][same], rtol=0, atol=1e-6)
    np.testing.assert_allclose(-val[same], golden[name + '_cost'][same], rtol=1e-6, atol=1e-10)
    # restarts whose path forked on a rounding-level accept / reject decision: still a valid solve -- SPD, value equal to
    # the oracle's EI at the returned point, and no worse than the start
    cons = octr.max_eigenvalue_constraint(max_eig)[0]
    for i in range(len(iters)):
        assert np.linalg.eigvalsh(x[i]).min() > 0
        e_here = ogp.ei_and_grad(gp, x[i], want_grad=False)[0]
        e_start = ogp.ei_and_grad(gp, golden[name + '_x0'][i], want_grad=False)[0]
        assert abs(val[i] - e_here) <= 1e-7 * max(1.0, abs(e_here)) and val[i] >= e_start - 1e-12
        if strict:      # the strict solver never accepts an infeasible point
            assert cons(x[i]) >= -1e-9 or cons(golden[name + '_x0'][i]) < 0
    assert set(reason.cpu().numpy().tolist()) <= {1, 2}


@pytest.mark.parametrize('d,n,R', [(1, 6, 9), (2, 10, 33), (3, 12, 64), (4, 20, 16), (5, 33, 20), (8, 24, 10)])
def test_spd_trust_region_kernel_follows_the_serial_solver(d, n, R):
    # plain TrustRegions on SPD(d): the oracle (oracle/rtr.py, pinned on the reference's own class on the sphere) solved
    # serially from the same starts; pymanopt PositiveDefinite operations (exp retraction, identity transport)
    from oracle import rtr as ortr
    rng, gp = spd_problem(d, n, beta=0.5 + math.log(2.0), noise=1e-2, seed=300 + d)
    x0 = ospd.spd_sample(rng, R, d, max_cond=50.0)
    x, val, iters, _ = ops.acq_rtr(device_gp(gp, _lib.GABO_F32), x0, maxiter=15)     # fp64 whatever the descriptor says
    x, val, iters = x.cpu().numpy(), val.cpu().numpy(), iters.cpu().numpy()
    nref = min(R, 10)
    opts = ortr.TROptions(maxiter=15)
    same = 0
    for i in range(nref):
        xi, ci, ki = ortr.solve_tr(gp, x0[i], opts)
        if int(iters[i]) == ki:
            same += 1
            np.testing.assert_allclose(x[i], xi, rtol=0, atol=1e-6 * max(1.0, np.abs(xi).max()))
            assert abs(-val[i] - ci) <= 1e-7 * max(1.0, abs(ci))
    assert same >= nref - 2
    ei0 = np.array([ogp.ei_and_grad(gp, xi, want_grad=False)[0] for xi in x0])
    assert (val >= ei0 - 1e-12).all()                                  # trust regions never accept a worse point
    np.testing.assert_allclose(x, np.swapaxes(x, -1, -2), rtol=0, atol=1e-12)
    assert np.linalg.eigvalsh(x).min() > 0


def test_spd_constrained_kernel_agrees_with_the_lockstep_driver_and_handles_two_constraints():
    # two independent implementations of the reference's constrained solver on the device: the one-launch kernel and the
    # lock-step driver (host logic pinned on CPU against the reference's class, tests/test_host_logic.py)
    import functools
    from gabotorch_b200 import manifold_optimization as mo, riemannian_utils as ru
    rng, gp = spd_problem(3, 16, beta=0.5 + math.log(2.0), noise=1e-2, seed=77)
    x0 = ospd.spd_sample(rng, 48, 3, min_eig=0.3, max_eig=2.5, max_cond=50.0)
    dgp = device_gp(gp, _lib.GABO_F64)
    cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=3.0),
            functools.partial(ru.min_eigenvalue_constraint_torch, minimum_eigenvalue=0.2)]
    for strict in (False, True):
        xk, vk, ik, _ = ops.acq_ctr(dgp, x0, [('max', 3.0), ('min', 0.2)], strict=strict, maxiter=40, mingradnorm=1e-4)
        xl, vl, il, _ = mo.batched_trust_regions(dgp, x0, maxiter=40, mingradnorm=1e-4, strict=strict,
                                                 ineq_constraints=mo.batched_constraints(cons, _lib.SPD))
        same = (ik == il).cpu().numpy()
        assert same.mean() >= 0.85, (ik, il)
        np.testing.assert_allclose(xk.cpu().numpy()[same], xl.cpu().numpy()[same], rtol=0, atol=1e-6)
        np.testing.assert_allclose(vk.cpu().numpy()[same], vl.cpu().numpy()[same], rtol=1e-6, atol=1e-10)
        if strict:
            lam = np.linalg.eigvalsh(xk.cpu().numpy())
            assert lam.max() <= 3.0 + 1e-9 and lam.min() >= 0.2 - 1e-9
    # a start that is not positive definite is reported, not solved
    bad = x0[:3].copy()
    bad[1] = -bad[1]
    xb, vb, ib, rb = ops.acq_ctr(dgp, bad, [('max', 3.0)])
    assert np.isnan(float(vb[1])) and int(rb[1]) == -1 and np.isfinite(vb.cpu().numpy()[[0, 2]]).all()


def test_spd_trust_regions_through_the_reference_api():
    # gabo_spd.py:183,200-203 through the drop-in surface: the constrained solve is ONE launch of gabo_acq_ctr
    import functools
    import gabotorch_b200 as g
    from gabotorch_b200 import riemannian_utils as ru
    rng = np.random.default_rng(5)
    xt = ospd.spd_sample(rng, 12, 2, max_cond=50.0)
    xv = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt))
    y = ospd.ackley(xv)
    model = g.ManifoldGP(xv, torch.from_numpy(y), g.ScaleKernel(g.SpdAffineInvariantGaussianKernel(beta_min=0.6)),
                         noise=1e-2)
    model.covar_module.outputscale = 1.0
    acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
    man = g.PositiveDefinite(2)
    man.min_eig, man.max_eig = 0.5, 1.8
    cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=2.0)]
    kw = dict(pre_processing_manifold=ru.vector_to_symmetric_matrix_mandel_torch,
              post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch, approx_hessian=True)
    ics = g.gen_batch_initial_conditions_manifold(acq, man, None, 1, 8, 64, options={'seed': 3},
                                                  post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch)
    for solver, extra in ((g.TrustRegions(maxiter=50), {}),
                          (g.ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100), dict(inequality_constraints=cons)),
                          (g.StrictConstrainedTrustRegions(mingradnorm=2e-4, maxiter=100, minstepsize=1e-4),
                           dict(inequality_constraints=cons))):
        cand, vals, info = g.gen_candidates_manifold(ics, acq, man, solver, return_info=True, **kw, **extra)
        assert tuple(cand.shape) == (8, 1, 3) and tuple(vals.shape) == (8,)
        start_vals = acq(ics)
        assert bool((vals.cpu() >= start_vals.cpu() - 1e-12).all())
        mats = ospd.vector_to_symmetric_matrix_mandel(cand[:, 0].cpu()).numpy()
        assert np.linalg.eigvalsh(mats).min() > 0
        if type(solver).__name__ == 'StrictConstrainedTrustRegions':
            assert np.linalg.eigvalsh(mats).max() <= 2.0 + 1e-9
