"""Pin the CPU oracle (oracle/) on outputs of the reference's own code (tests/golden/reference_vectors.npz, generated
by tests/golden/make_golden.py from /root/reference) and, when the reference tree is present, on the live reference.
CPU only."""
import numpy as np
import pytest
import torch

from oracle import gp as ogp
from oracle import nested as onest
from oracle import rcg as orcg
from oracle import reference_loader
from oracle import spd as ospd
from oracle import sphere as osph

SPHERE_SETS = ['s2_n256', 's5_n1024', 's8_n96']
SPD_SETS = ['spd3_n128', 'spd8_n64', 'spd2_n40', 'spd5_n48']


@pytest.mark.parametrize('name', SPHERE_SETS)
def test_sphere_distance_and_kernel_match_reference(golden, name):
    x = golden[name + '_x']
    st = int(golden[name + '_stride'])
    d = osph.sphere_distance(x, x).numpy()
    k = osph.sphere_gaussian_kernel(x, x, float(golden[name + '_beta'])).numpy()
    # einsum vs the reference's bmm may differ in the last bit of the inner product; acos amplifies that near +-1
    np.testing.assert_allclose(d[::st, ::st], golden[name + '_d'], rtol=0, atol=3e-8)
    np.testing.assert_allclose(k[::st, ::st], golden[name + '_k'], rtol=1e-12, atol=1e-15)
    assert abs(k.sum() - float(golden[name + '_ksum'])) < 1e-9 * k.size
    np.testing.assert_allclose(osph.sphere_distance(x, x, diag=True).numpy(), golden[name + '_ddiag'], atol=3e-8)
    # the reference's own op sequence (cat + bmm) restated: identical numbers
    n = min(64, x.shape[0])
    np.testing.assert_allclose(osph.sphere_distance_loop(x[:n], x[:n]).numpy(), d[:n, :n], atol=3e-8)


def test_sphere_quirks(golden):
    x = golden['s2_n256_x']
    d = osph.sphere_distance(x[:4], x[:4]).numpy()
    assert np.allclose(np.diag(d), np.arccos(1 - 1e-15), atol=2e-8)      # d(x,x) = 4.47e-8, not 0
    a, b = golden['s3_rect_a'], golden['s3_rect_b']
    np.testing.assert_allclose(osph.sphere_distance(a, b).numpy(), golden['s3_rect_d'], atol=3e-8)
    assert abs(golden['s3_rect_d'][1, 1] - np.pi) < 1e-7                  # antipodal pair


@pytest.mark.parametrize('name', SPD_SETS)
def test_mandel_and_spd_distance_match_reference(golden, name):
    m, v = golden[name + '_mat'], golden[name + '_vec']
    np.testing.assert_allclose(ospd.symmetric_matrix_to_vector_mandel(m).numpy(), v, rtol=0, atol=1e-15)
    np.testing.assert_allclose(ospd.vector_to_symmetric_matrix_mandel(v).numpy(), golden[name + '_unpacked'],
                               rtol=0, atol=1e-15)
    back = golden[name + '_unpacked']
    d = ospd.affine_invariant_distance(back, back).numpy()
    # eigenvalues pass through float32 in the reference (spd_utils_torch.py:108): LAPACK driver differences
    # (single-matrix symeig vs batched eigh) show up at the float32 rounding level only
    np.testing.assert_allclose(d, golden[name + '_d'], rtol=3e-7, atol=2e-7)
    k = ospd.spd_affine_invariant_gaussian_kernel(v, v, float(golden[name + '_beta'])).numpy()
    np.testing.assert_allclose(k, golden[name + '_k'], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(ospd.frobenius_distance(back, back).numpy(), golden[name + '_frob'], rtol=1e-12,
                               atol=1e-14)
    np.testing.assert_allclose(ospd.logm(back).numpy(), golden[name + '_logm'], rtol=0, atol=1e-10)


def test_spd_quirks(golden):
    a, b = golden['spd3_rect_a'], golden['spd3_rect_b']
    np.testing.assert_allclose(ospd.affine_invariant_distance(a, b).numpy(), golden['spd3_rect_d'], rtol=3e-7,
                               atol=2e-7)
    d = golden['spd3_n128_d']
    assert np.allclose(np.diag(d), np.sqrt(1e-15), atol=1e-9)            # d(X,X) = 3.16e-8
    assert np.abs(d - d.T).max() < 5e-6                                   # symmetric only to float32 accuracy
    z = ospd.affine_invariant_distance(a, a, diagonal_distance=True)
    assert tuple(z.shape) == (a.shape[0], 1) and float(z.abs().max()) == 0.0
    # the reference's loop structure restated gives the same numbers as the vectorised oracle
    np.testing.assert_allclose(ospd.affine_invariant_distance_loop(a[:5], b[:6]).numpy(),
                               ospd.affine_invariant_distance(a[:5], b[:6]).numpy(), rtol=3e-7, atol=2e-7)


@pytest.mark.parametrize('name', ['proj_20_5', 'proj_5_2'])
def test_nested_projection_matches_reference(golden, name):
    x, w = golden[name + '_x'], golden[name + '_w']
    y = onest.projection_from_spd_to_nested_spd(x, w).numpy()
    np.testing.assert_allclose(y, golden[name + '_y'], rtol=1e-12, atol=1e-12)
    P = onest.mandel_projection_matrix(w)
    np.testing.assert_allclose(golden[name + '_xvec'] @ P.T, golden[name + '_yvec'], rtol=1e-11, atol=1e-11)


def test_sphere_manifold_formulas_match_reference_numpy(golden):
    x, y, v = golden['man_sphere_x'], golden['man_sphere_y'], golden['man_sphere_v']
    np.testing.assert_allclose(osph.log(x, y), golden['man_sphere_log'], atol=1e-12)
    np.testing.assert_allclose(osph.exp(x, 0.7 * golden['man_sphere_log']), golden['man_sphere_exp07'], atol=1e-12)
    np.testing.assert_allclose(osph.dist(x, y), golden['man_sphere_dist'], atol=1e-12)
    np.testing.assert_allclose(osph.parallel_transport(x, y, v), golden['man_sphere_pt'], atol=1e-12)


@pytest.mark.parametrize('d', [3, 8])
def test_spd_manifold_formulas_match_reference_numpy(golden, d):
    tag = 'man_spd%d_' % d
    x, y, u = golden[tag + 'x'], golden[tag + 'y'], golden[tag + 'u']
    np.testing.assert_allclose(ospd.log(x, y), golden[tag + 'log'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ospd.exp(x, 0.5 * golden[tag + 'log']), golden[tag + 'exp05'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ospd.dist(x, y), golden[tag + 'dist'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(ospd.parallel_transport(x, y, u), golden[tag + 'pt'], rtol=1e-8, atol=1e-8)


def test_manifold_identities():
    rng = np.random.default_rng(7)
    x, y = osph.rand(rng, 16, 6), osph.rand(rng, 16, 6)
    np.testing.assert_allclose(osph.exp(x, osph.log(x, y)), y, atol=1e-12)
    u = osph.proj(x, rng.standard_normal((16, 6)))
    assert np.abs(np.sum(osph.transp(x, y, u) * y, axis=-1)).max() < 1e-14       # transported vector is tangent at y
    np.testing.assert_allclose(np.linalg.norm(osph.parallel_transport(x, y, u), axis=-1),
                               np.linalg.norm(u, axis=-1), rtol=1e-12)
    X, Y = ospd.spd_sample(rng, 8, 4, max_cond=100), ospd.spd_sample(rng, 8, 4, max_cond=100)
    np.testing.assert_allclose(ospd.exp(X, ospd.log(X, Y)), Y, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(ospd.norm(X, ospd.log(X, Y)), ospd.dist(X, Y), rtol=1e-9)


@pytest.mark.parametrize('manifold', ['sphere', 'spd'])
def test_closed_form_ei_gradient_matches_autograd(manifold):
    rng = np.random.default_rng(11)
    if manifold == 'sphere':
        xt = osph.rand(rng, 12, 5)
        gp = ogp.make_gp('sphere', xt, osph.ackley(xt), beta=2.0, noise=1e-2)
        x = osph.rand(rng, 1, 5)[0]
        xx = torch.tensor(x, requires_grad=True)
        ogp.ei_torch(gp, xx).backward()
        rg = osph.proj(x, xx.grad.numpy())
    else:
        xt = ospd.spd_sample(rng, 10, 3, max_cond=100)
        y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(xt))
        gp = ogp.make_gp('spd', xt, y, beta=0.7, noise=1e-2)
        x = ospd.spd_sample(rng, 1, 3, max_cond=100)[0]
        xx = torch.tensor(x, requires_grad=True)
        ogp.ei_torch(gp, xx).backward()
        rg = ospd.egrad2rgrad(x, xx.grad.numpy())
    ei, g = ogp.ei_and_grad(gp, x)
    assert ei > 0
    np.testing.assert_allclose(g, rg, rtol=1e-7, atol=1e-12 + 1e-7 * np.abs(rg).max())


def test_oracle_cg_decreases_cost_and_selection_rule():
    rng = np.random.default_rng(3)
    xt = osph.rand(rng, 16, 4)
    gp = ogp.make_gp('sphere', xt, osph.ackley(xt), beta=2.0, noise=1e-2)
    x0 = osph.rand(rng, 4, 4)
    xs, vals, its = orcg.gen_candidates(gp, x0, orcg.CGOptions(maxiter=30))
    assert np.all(vals >= ogp.ei_batch(gp, x0) - 1e-12)
    assert np.allclose(np.linalg.norm(xs, axis=-1), 1.0, atol=1e-12)
    v = np.array([0.1, np.nan, 0.7, 0.7, -1.0])
    assert orcg.best_candidate(v) == 2
    assert orcg.lexi_argmax_records(v, [9, 3, 5, 4, 0]) == 3        # tie on value -> lower global index wins


@pytest.mark.skipif(not reference_loader.available(), reason='/root/reference not present (GPU box)')
def test_oracle_against_live_reference():
    ref = reference_loader.load()
    rng = np.random.default_rng(5)
    x = osph.rand(rng, 40, 4)
    np.testing.assert_allclose(osph.sphere_distance(x, x).numpy(),
                               ref.sphere_utils_torch.sphere_distance_torch(torch.from_numpy(x),
                                                                            torch.from_numpy(x)).numpy(), atol=3e-8)
    m = torch.from_numpy(ospd.spd_sample(rng, 12, 4, max_cond=100))
    np.testing.assert_allclose(ospd.affine_invariant_distance(m, m).numpy(),
                               ref.spd_utils_torch.affine_invariant_distance_torch(m, m).numpy(), rtol=3e-7, atol=2e-7)
    v = ref.spd_utils_torch.symmetric_matrix_to_vector_mandel_torch(m)
    np.testing.assert_allclose(ospd.symmetric_matrix_to_vector_mandel(m).numpy(), v.numpy(), atol=1e-15)


@pytest.mark.parametrize('name,D,dl', [('nsph_5_3', 5, 3), ('nsph_6_2', 6, 2)])
def test_nested_sphere_projection_matches_reference(golden, name, D, dl):
    # nested_spheres_utils.py:120-147 run from the reference's own code (tests/golden/make_golden.py)
    from oracle import nested_sphere as ons
    axes = [torch.from_numpy(golden[name + '_axis%d' % lvl]) for lvl in range(D - dl)]
    r = float(golden[name + '_r'])
    levels = ons.projection_from_sphere_to_subsphere(golden[name + '_x'], axes, [r] * len(axes))
    for lvl in range(D - dl):
        np.testing.assert_allclose(levels[lvl + 1].numpy(), golden[name + '_y%d' % lvl], rtol=0, atol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(levels[-1].numpy(), axis=-1), 1.0, atol=2e-6)   # the 1e-6 of :112


@pytest.mark.parametrize('name,D,dl', [('nsph_5_3', 5, 3), ('nsph_6_2', 6, 2)])
def test_nested_sphere_reconstruction_matches_reference(golden, name, D, dl):
    # nested_spheres_utils.py:13-67 and :149-213 run from the reference's own code
    from oracle import nested_sphere as ons
    axes = [torch.from_numpy(golden[name + '_axis%d' % lvl]) for lvl in range(D - dl)]
    r = float(golden[name + '_r'])
    ns = ons.projection_to_nested_sphere(golden[name + '_x'], axes[0], r)
    np.testing.assert_allclose(ns.numpy(), golden[name + '_ns0'], rtol=0, atol=1e-14)
    ups = ons.projection_from_subsphere_to_sphere(golden[name + '_y%d' % (D - dl - 1)], axes, [r] * len(axes))
    for lvl in range(D - dl):
        np.testing.assert_allclose(ups[lvl + 1].numpy(), golden[name + '_up%d' % lvl], rtol=0, atol=1e-14)
    # the reconstruction lands on the nested sphere of every level: distance r to that level's axis
    top = ups[-1].numpy()
    np.testing.assert_allclose(np.arccos(np.clip(top @ axes[0].numpy().reshape(-1), -1, 1)), r, atol=1e-9)


@pytest.mark.parametrize('d', [3, 5])
def test_sqrtm_matches_reference(golden, d):
    from oracle import spd as ospd_
    y = ospd_.sqrtm(torch.from_numpy(golden['sqrtm%d_x' % d])).numpy()
    np.testing.assert_allclose(y, golden['sqrtm%d_y' % d], rtol=0, atol=1e-12)
    np.testing.assert_allclose(y @ y, golden['sqrtm%d_x' % d], rtol=0, atol=1e-12)


@pytest.mark.parametrize('name', ['recon_5_2', 'recon_20_5'])
def test_nested_spd_reconstruction_matches_reference(golden, name):
    # nested_spd_utils.py:51-118 run from the reference's own code
    from oracle import nested as onest
    w, v, c, k = (golden[name + s] for s in ('_w', '_v', '_c', '_k'))
    x = onest.projection_from_nested_spd_to_spd(golden[name + '_y'], w, v, c, k).numpy()
    np.testing.assert_allclose(x, golden[name + '_x'], rtol=0, atol=1e-11)
    # right inverse of the projection: W^T X W = Y; and X is positive definite (|K| < 1)
    back = onest.projection_from_spd_to_nested_spd(x, w).numpy()
    np.testing.assert_allclose(back, golden[name + '_y'], rtol=0, atol=1e-11)
    assert np.linalg.eigvalsh(0.5 * (x + np.swapaxes(x, -1, -2))).min() > 0


@pytest.mark.parametrize('name', ['rtr_s2', 'rtr_s5', 'rtr_s5_noisy'])
def test_trust_region_oracle_matches_reference_solver(golden, name):
    # robust_trust_regions.py (the reference's own TrustRegions class) + approximate_hessian.py, run by make_golden.py
    from oracle import gp as ogp
    from oracle import rtr as ortr
    beta, noise = golden[name + '_hyper']
    gp = ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))
    xs, vals, its = ortr.gen_candidates(gp, golden[name + '_x0'])
    np.testing.assert_array_equal(its, golden[name + '_iters'])
    np.testing.assert_allclose(xs, golden[name + '_x'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(-vals, golden[name + '_cost'], rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize('name', ['ctr_spd2', 'ctr_spd3', 'ctr_spd2_active'])
def test_constrained_trust_region_oracle_matches_reference_solver(golden, name):
    # constrained_trust_regions.py (the reference's own ConstrainedTrustRegions class, its Problem + PyTorch autodiff
    # for the max-eigenvalue constraint) in the configuration of gabo_spd.py, run by make_golden.py
    from oracle import ctr as octr
    from oracle import gp as ogp
    from oracle import rtr as ortr
    beta, noise, max_eig = golden[name + '_hyper']
    gp = ogp.make_gp('spd', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))
    opts = ortr.TROptions(mingradnorm=1e-4, maxiter=100)
    cons = [octr.max_eigenvalue_constraint(float(max_eig))]
    for i, x0 in enumerate(golden[name + '_x0']):
        x, c, k = octr.solve_ctr(gp, x0, ineq_constraints=cons, opts=opts)
        assert k == int(golden[name + '_iters'][i])
        np.testing.assert_allclose(x, golden[name + '_x'][i], rtol=0, atol=1e-9)
        assert abs(c - golden[name + '_cost'][i]) <= 1e-10 * max(1.0, abs(c))


@pytest.mark.parametrize('name', ['sctr_spd2_active', 'sctr_spd3'])
def test_strict_constrained_trust_region_oracle_matches_reference_solver(golden, name):
    # StrictConstrainedTrustRegions (constrained_trust_regions.py:737-1415) with hd_gabo_spd.py's settings
    from oracle import ctr as octr
    from oracle import gp as ogp
    from oracle import rtr as ortr
    beta, noise, max_eig = golden[name + '_hyper']
    gp = ogp.make_gp('spd', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))
    opts = ortr.TROptions(mingradnorm=2e-4, maxiter=100)
    cons = [octr.max_eigenvalue_constraint(float(max_eig))]
    for i, x0 in enumerate(golden[name + '_x0']):
        x, c, k = octr.solve_ctr(gp, x0, ineq_constraints=cons, opts=opts, strict=True)
        assert k == int(golden[name + '_iters'][i])
        np.testing.assert_allclose(x, golden[name + '_x'][i], rtol=0, atol=1e-9)
        assert abs(c - golden[name + '_cost'][i]) <= 1e-10 * max(1.0, abs(c))
        assert np.linalg.eigvalsh(x)[-1] <= float(max_eig) + 1e-12            # strict: never leaves the feasible set


@pytest.mark.parametrize('name', ['ctr_s2_domain', 'ctr_s2_domain_active'])
def test_constrained_trust_region_oracle_on_the_sphere_matches_reference_solver(golden, name):
    # ConstrainedTrustRegions(maxiter=200) with the domain constraint of gabo_sphere_inequality_constraints.py
    from oracle import ctr as octr
    from oracle import gp as ogp
    from oracle import rtr as ortr
    beta, noise, angle = golden[name + '_hyper']
    gp = ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))
    cons = [octr.sphere_domain_constraint([1.0, 0.0, 0.0], float(angle))]
    opts = ortr.TROptions(maxiter=200)
    for i, x0 in enumerate(golden[name + '_x0'][:5]):
        x, c, k = octr.solve_ctr(gp, x0, ineq_constraints=cons, opts=opts)
        assert k == int(golden[name + '_iters'][i])
        np.testing.assert_allclose(x, golden[name + '_x'][i], rtol=0, atol=1e-9)
        assert abs(c - golden[name + '_cost'][i]) <= 1e-10 * max(1.0, abs(c))


def test_equality_constrained_trust_region_oracle_matches_reference_solver(golden):
    # the great-circle constraint x[1] = 0 of gabo_sphere_equality_constraints.py with ConstrainedTrustRegions(maxiter=200)
    from oracle import ctr as octr
    from oracle import gp as ogp
    from oracle import rtr as ortr
    from oracle import sphere as osph
    name = 'ctr_s2_circle'
    beta, noise = golden[name + '_hyper']
    gp = ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))
    e1 = np.array([0.0, 1.0, 0.0])
    cons = [(lambda x: x[1], lambda x: osph.proj(x, e1))]
    for i, x0 in enumerate(golden[name + '_x0']):
        x, c, k = octr.solve_ctr(gp, x0, eq_constraints=cons, opts=ortr.TROptions(maxiter=200))
        assert k == int(golden[name + '_iters'][i])
        np.testing.assert_allclose(x, golden[name + '_x'][i], rtol=0, atol=1e-9)
        assert abs(c - golden[name + '_cost'][i]) <= 1e-10 * max(1.0, abs(c))


@pytest.mark.parametrize('name,kind', [('alm_s2_domain', 'ineq'), ('alm_s2_circle', 'eq')])
def test_augmented_lagrangian_oracle_matches_reference_solver(golden, name, kind):
    # the reference's own AugmentedLagrangeMethod around its own TrustRegions (make_golden.py), constraints of the
    # constrained sphere examples
    from oracle import alm as oalm
    from oracle import ctr as octr
    from oracle import gp as ogp
    from oracle import sphere as osph
    beta, noise, angle = golden[name + '_hyper']
    gp = ogp.make_gp('sphere', golden[name + '_xtrain'], golden[name + '_y'], beta=float(beta), noise=float(noise))
    e1 = np.array([0.0, 1.0, 0.0])
    kw = ({'ineq_constraints': [octr.sphere_domain_constraint([1.0, 0.0, 0.0], float(angle))]} if kind == 'ineq'
          else {'eq_constraints': [(lambda x: x[1], lambda x: osph.proj(x, e1))]})
    for i, x0 in enumerate(golden[name + '_x0']):
        x, k = oalm.solve_alm(gp, x0, maxiter=30, inner_opts={'maxiter': 50}, gammas_fact=0.05, **kw)
        assert k == int(golden[name + '_iters'][i])
        np.testing.assert_allclose(x, golden[name + '_x'][i], rtol=0, atol=1e-9)


def test_nested_spd_reconstruction_costs_oracle_vs_reference():
    """oracle.nested.min_*_reconstruction_cost against values produced by the reference's own functions
    (nested_spd_optimization.py:22-92; tests/golden/make_golden_recon_cost.py).  Both sides accumulate in float32."""
    import os
    from oracle import nested as onest
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'recon_cost_vectors.npz'))
    for name in ('rc_6_2', 'rc_10_3', 'rc_20_5'):
        args = [gold[name + '_' + k] for k in ('x', 'y', 'w', 'v', 'c', 'k')]
        ai = float(onest.min_affine_invariant_distance_reconstruction_cost(*args))
        le = float(onest.min_log_euclidean_distance_reconstruction_cost(*args))
        assert abs(ai - float(gold[name + '_ai'])) <= 2e-6 * abs(ai), (name, ai, float(gold[name + '_ai']))
        assert abs(le - float(gold[name + '_le'])) <= 2e-6 * abs(le), (name, le, float(gold[name + '_le']))
