"""Nested SPD projection Y = W^T X W (P1): tensor-core Mandel contraction against the reference's bmm outputs."""
import numpy as np
import pytest
import torch

from gabotorch_b200 import nested_mappings as nm
from gabotorch_b200 import ops
from oracle import nested as onest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['proj_20_5', 'proj_5_2'])
def test_projection_golden(golden, name):
    x, w = golden[name + '_x'], golden[name + '_w']
    y = nm.projection_from_spd_to_nested_spd(torch.from_numpy(x), torch.from_numpy(w))
    assert y.dtype == torch.float64 and not y.is_cuda
    scale = np.abs(golden[name + '_y']).max()
    np.testing.assert_allclose(y.numpy(), golden[name + '_y'], rtol=0, atol=1e-5 * scale)
    ym = nm.projection_mandel(torch.from_numpy(golden[name + '_xvec']), w).cpu().numpy()
    np.testing.assert_allclose(ym, golden[name + '_yvec'], rtol=0, atol=1e-5 * np.abs(golden[name + '_yvec']).max())


@pytest.mark.parametrize('D,d,n', [(20, 5, 1), (20, 5, 63), (20, 5, 64), (20, 5, 65), (20, 5, 30000), (5, 2, 1001),
                                   (6, 3, 777), (12, 8, 500), (8, 8, 130), (3, 1, 50), (22, 4, 257)])
def test_projection_vs_oracle_ragged(D, d, n):
    rng = np.random.default_rng(D * 100 + d)
    dvh = D * (D + 1) // 2
    xv = rng.standard_normal((n, dvh))
    w = onest.grassmann_rand(rng, D, d)
    P = onest.mandel_projection_matrix(w)
    ref = xv.astype(np.float32).astype(np.float64) @ P.T
    got = nm.projection_mandel(torch.from_numpy(xv), w).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-6 * np.abs(ref).max())


def test_projection_linearity_and_spd_preservation():
    rng = np.random.default_rng(3)
    D, d, n = 20, 5, 4096
    proj = nm.NestedSpdProjection(onest.grassmann_rand(rng, D, d))
    a = torch.from_numpy(rng.standard_normal((n, 210))).float()
    b = torch.from_numpy(rng.standard_normal((n, 210))).float()
    lhs = proj.mandel(2.0 * a + b)
    rhs = 2.0 * proj.mandel(a) + proj.mandel(b)
    assert float((lhs - rhs).abs().max()) <= 2e-5 * float(rhs.abs().max())
    m = rng.standard_normal((256, D, D))
    x = m @ np.swapaxes(m, -1, -2) + 0.1 * np.eye(D)
    y = nm.projection_from_spd_to_nested_spd(torch.from_numpy(x), proj_matrix := onest.grassmann_rand(rng, D, d))
    assert np.linalg.eigvalsh(y.numpy()).min() > 0
    np.testing.assert_allclose(y.numpy(), onest.projection_from_spd_to_nested_spd(x, proj_matrix).numpy(), rtol=0,
                               atol=1e-5 * np.abs(x).max())


@pytest.mark.parametrize('name', ['proj_20_5', 'proj_5_2'])
def test_projection_f64_golden(golden, name):
    # reference-precision path used by the nested kernels: golden vectors generated from the reference's own bmm code
    ym = ops.nested_spd_project_f64(torch.from_numpy(golden[name + '_xvec']), torch.from_numpy(golden[name + '_w']))
    np.testing.assert_allclose(ym.cpu().numpy(), golden[name + '_yvec'], rtol=1e-12,
                               atol=1e-13 * np.abs(golden[name + '_yvec']).max())


@pytest.mark.parametrize('D,d,n1,n2', [(5, 2, 35, 20), (20, 5, 40, 33), (8, 3, 17, 17)])
def test_nested_spd_kernels_vs_oracle(D, d, n1, n2):
    # P2 of SURVEY section 8: kernels_nested_spd.py:104-136 (affine-invariant) and :191-246 (log-Euclidean); the oracle
    # composes the pieces pinned on the reference (Mandel unpack, W^T X W, AI distance / logm + Frobenius distance)
    import gabotorch_b200 as g
    from oracle import spd as ospd
    rng = np.random.default_rng(D * 10 + d)
    x1 = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, n1, D, max_cond=100.0)))
    x2 = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, n2, D, max_cond=100.0)))
    w = onest.grassmann_rand(rng, D, d)
    k = g.NestedSpdAffineInvariantGaussianKernel(D, d, beta_min=0.25)
    k.projection_matrix = torch.from_numpy(w)
    assert k.raw_projection_matrix.dtype == torch.float32         # the reference's parameter dtype
    w32 = k.raw_projection_matrix.detach().double().numpy()        # what both sides project with
    beta = float(k.beta.detach())
    with torch.no_grad():
        got = k.forward(x1, x2)
        sym = k.forward(x1, x1)
    assert got.dtype == torch.float64 and not got.is_cuda and tuple(got.shape) == (n1, n2)
    ref = onest.nested_spd_affine_invariant_gaussian_kernel(x1, x2, torch.from_numpy(w32), beta).numpy()
    m = ref >= 1e-6
    d_ref = np.sqrt(-np.log(np.maximum(ref, 1e-300)) / beta)
    tol = 1e-5 * np.maximum(1.0, 2 * beta * d_ref ** 2)            # bound implied by the distance tolerance (fp32 Jacobi)
    assert np.all(np.abs(got.numpy() - ref)[m] <= tol[m] * ref[m])
    np.testing.assert_allclose(sym.numpy(), onest.nested_spd_affine_invariant_gaussian_kernel(
        x1, x1, torch.from_numpy(w32), beta).numpy(), rtol=0, atol=2e-5)
    assert torch.all(k.forward(x1, x1, diagonal_distance=True) == 1) and k.forward(x1, x1, diagonal_distance=True).shape == (n1, 1)

    le = g.NestedSpdLogEuclideanGaussianKernel(D, d)
    le.projection_matrix = torch.from_numpy(w)
    le.lengthscale = 1.7
    ls = float(le.lengthscale.detach())
    with torch.no_grad():
        got_le = le.forward(x1, x2)
    ref_le = onest.nested_spd_log_euclidean_gaussian_kernel(x1, x2, torch.from_numpy(w32), ls).numpy()
    np.testing.assert_allclose(got_le.numpy(), ref_le, rtol=1e-7, atol=1e-12)
    # beta / lengthscale AND the Grassmann parameter are differentiable (fit_gpytorch_manifold); values checked in
    # tests/test_grad_gpu.py
    out = k.forward(x1, x2)
    out.sum().backward()
    assert k.raw_beta.grad is not None and torch.isfinite(k.raw_beta.grad).all()
    assert k.raw_projection_matrix.grad is not None and torch.isfinite(k.raw_projection_matrix.grad).all()


@pytest.mark.parametrize('name,D,dl', [('nsph_5_3', 5, 3), ('nsph_6_2', 6, 2)])
def test_nested_sphere_projection_golden(golden, name, D, dl):
    axes = [torch.from_numpy(golden[name + '_axis%d' % lvl]) for lvl in range(D - dl)]
    r = float(golden[name + '_r'])
    y = ops.nested_sphere_project(torch.from_numpy(golden[name + '_x']), axes, [torch.tensor([[r]])] * len(axes))
    np.testing.assert_allclose(y.cpu().numpy(), golden[name + '_y%d' % (D - dl - 1)], rtol=0, atol=1e-12)


def test_nested_sphere_kernel_vs_oracle():
    # kernels_nested_sphere.py:129-152 (hd_gabo_sphere.py:134 uses NestedSphereGaussianKernel(5 -> 3))
    import gabotorch_b200 as g
    from oracle import nested_sphere as ons
    from oracle import sphere as osph
    rng = np.random.default_rng(8)
    x1, x2 = osph.rand(rng, 45, 5), osph.rand(rng, 31, 5)
    k = g.NestedSphereGaussianKernel(5, 3, beta_min=2.0)
    axes = [a.detach().double() for a in k.axes]
    assert [tuple(a.shape) for a in axes] == [(1, 5), (1, 4)] and len(k.distances_to_axis) == 2
    beta = float(k.beta.detach())
    with torch.no_grad():
        got = k.forward(torch.from_numpy(x1), torch.from_numpy(x2))
        dg = k.forward(torch.from_numpy(x1), torch.from_numpy(x1), diag=True)
    ref = ons.nested_sphere_gaussian_kernel(x1, x2, axes, k.distances_to_axis, beta).numpy()
    m = ref >= 1e-6
    assert np.all(np.abs(got.numpy() - ref)[m] <= 1e-5 * ref[m]) and tuple(dg.shape) == (45, 1)
    new_axes = [osph.rand(rng, 1, 5), osph.rand(rng, 1, 4)]
    k.axes = new_axes
    ref2 = ons.nested_sphere_gaussian_kernel(x1, x2, [torch.from_numpy(a).float().double() for a in new_axes],
                                             k.distances_to_axis, beta).numpy()
    with torch.no_grad():
        got2 = k.forward(torch.from_numpy(x1), torch.from_numpy(x2)).numpy()
    m2 = ref2 >= 1e-6
    assert np.all(np.abs(got2 - ref2)[m2] <= 1e-5 * ref2[m2])


# ---- reconstruction maps -------------------------------------------------------------------------

@pytest.mark.parametrize('name,D,dl', [('nsph_5_3', 5, 3), ('nsph_6_2', 6, 2)])
def test_nested_sphere_chain_and_reconstruction_golden(golden, name, D, dl):
    # nested_spheres_utils.py:13-67, :117-147, :149-213 against the reference's own outputs
    axes = [torch.from_numpy(golden[name + '_axis%d' % lvl]) for lvl in range(D - dl)]
    r = float(golden[name + '_r'])
    dists = [torch.tensor([[r]], dtype=torch.float64)] * len(axes)
    x = torch.from_numpy(golden[name + '_x'])
    down = nm.projection_from_sphere_to_subsphere(x, axes, dists)
    assert len(down) == D - dl + 1 and down[0] is x
    for lvl in range(D - dl):
        assert down[lvl + 1].dtype == torch.float64 and not down[lvl + 1].is_cuda
        np.testing.assert_allclose(down[lvl + 1].numpy(), golden[name + '_y%d' % lvl], rtol=0, atol=1e-12)
    one = nm.projection_from_sphere_to_next_subsphere(x, axes[0], dists[0])
    np.testing.assert_allclose(one.numpy(), golden[name + '_y0'], rtol=0, atol=1e-12)
    ns = nm.projection_from_sphere_to_nested_sphere(x, axes[0], dists[0])
    np.testing.assert_allclose(ns.numpy(), golden[name + '_ns0'], rtol=0, atol=1e-12)
    y_low = torch.from_numpy(golden[name + '_y%d' % (D - dl - 1)])
    ups = nm.projection_from_subsphere_to_sphere(y_low, axes, dists)
    assert len(ups) == D - dl + 1
    for lvl in range(D - dl):
        np.testing.assert_allclose(ups[lvl + 1].numpy(), golden[name + '_up%d' % lvl], rtol=0, atol=1e-12)
    nxt = nm.projection_from_subsphere_to_next_sphere(y_low, axes[-1], dists[-1])
    np.testing.assert_allclose(nxt.numpy(), golden[name + '_up0'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('D,dl,n', [(3, 2, 1), (8, 3, 1000), (20, 5, 4097), (64, 60, 33)])
def test_nested_sphere_round_trip_vs_oracle(D, dl, n):
    # projecting the reconstruction returns the latent points (the projection is a left inverse up to the 1e-6
    # regularisers of nested_spheres_utils.py:58,107,112); both directions against the oracle on ragged sizes
    from oracle import nested_sphere as ons
    from oracle import sphere as osph
    rng = np.random.default_rng(D * 10 + dl)
    axes = [torch.from_numpy(osph.rand(rng, 1, k)) for k in range(D, dl, -1)]
    dists = [torch.tensor([[0.9 + 0.1 * i]], dtype=torch.float64) for i in range(len(axes))]
    y = torch.from_numpy(osph.rand(rng, n, dl))
    ups = ops.nested_sphere_reconstruct(y, axes, dists)
    ref = ons.projection_from_subsphere_to_sphere(y, axes, [float(r) for r in dists])
    for got, want in zip(ups, ref[1:]):
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(ups[-1].cpu().numpy(), axis=-1), 1.0, atol=1e-12)
    back = ops.nested_sphere_chain(ups[-1], axes, dists)
    ref_back = ons.projection_from_sphere_to_subsphere(ups[-1].cpu(), axes, [float(r) for r in dists])
    for got, want in zip(back, ref_back[1:]):
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=0, atol=1e-11)
    np.testing.assert_allclose(back[-1].cpu().numpy(), y.numpy(), rtol=0, atol=2e-5 * (D - dl))
    np.testing.assert_array_equal(ops.nested_sphere_project(ups[-1], axes, dists).cpu().numpy(),
                                  back[-1].cpu().numpy())


@pytest.mark.parametrize('d', [3, 5])
def test_sqrtm_golden(golden, d):
    y = ops.spd_sqrtm(torch.from_numpy(golden['sqrtm%d_x' % d])).cpu().numpy()
    np.testing.assert_allclose(y, golden['sqrtm%d_y' % d], rtol=0, atol=1e-12)


@pytest.mark.parametrize('d', [1, 2, 4, 6, 7, 8])
def test_sqrtm_squares_back(d):
    from oracle import spd as ospd
    x = ospd.spd_sample(np.random.default_rng(d), 300, d)
    y = ops.spd_sqrtm(torch.from_numpy(x)).cpu().numpy()
    np.testing.assert_allclose(y @ y, x, rtol=0, atol=1e-12 * np.abs(x).max())
    np.testing.assert_allclose(y, np.swapaxes(y, -1, -2), rtol=0, atol=1e-13)
    bad = x.copy()
    bad[7] = -bad[7]
    out = ops.spd_sqrtm(torch.from_numpy(bad)).cpu().numpy()
    assert np.isnan(out[7]).all() and not np.isnan(out[6]).any()


@pytest.mark.parametrize('name', ['recon_5_2', 'recon_20_5'])
def test_nested_spd_reconstruction_golden(golden, name):
    # nested_spd_utils.py:51-118 against the reference's own outputs, batch and single-matrix forms
    w, v, c, k = (torch.from_numpy(golden[name + s]) for s in ('_w', '_v', '_c', '_k'))
    y = torch.from_numpy(golden[name + '_y'])
    x = nm.projection_from_nested_spd_to_spd(y, w, v, c, k)
    assert x.dtype == torch.float64 and not x.is_cuda and tuple(x.shape) == golden[name + '_x'].shape
    scale = np.abs(golden[name + '_x']).max()
    np.testing.assert_allclose(x.numpy(), golden[name + '_x'], rtol=0, atol=1e-11 * scale)
    x0 = nm.projection_from_nested_spd_to_spd(y[0], w, v, c, k)
    np.testing.assert_allclose(x0.numpy(), golden[name + '_x'][0], rtol=0, atol=1e-11 * scale)


@pytest.mark.parametrize('D,d,n', [(20, 5, 3000), (9, 8, 257), (32, 3, 100), (2, 1, 5)])
def test_nested_spd_reconstruction_right_inverse(D, d, n):
    # W^T X W = Y (the reconstruction is a right inverse of P1), X positive definite for |K| < 1, oracle on a sample
    from oracle import spd as ospd
    rng = np.random.default_rng(D + d)
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    w, v = q[:, :d].copy(), q[:, d:].copy()
    c = ospd.spd_sample(rng, 1, D - d)[0]
    k = rng.standard_normal((d, D - d))
    k = 0.9 * k / np.linalg.norm(k, 2)
    y = ospd.spd_sample(rng, n, d)
    rec = nm.NestedSpdReconstruction(torch.from_numpy(w), torch.from_numpy(v), torch.from_numpy(c), torch.from_numpy(k))
    x = rec(torch.from_numpy(y)).cpu().numpy()
    ref = onest.projection_from_nested_spd_to_spd(y[:40], w, v, c, k).numpy()
    np.testing.assert_allclose(x[:40], ref, rtol=0, atol=1e-11 * np.abs(ref).max())
    np.testing.assert_allclose(np.swapaxes(w, 0, 1) @ x @ w, y, rtol=0, atol=1e-11 * np.abs(y).max())
    assert np.linalg.eigvalsh(0.5 * (x + np.swapaxes(x, -1, -2))).min() > 0
    with pytest.raises(ops.NotPositiveDefiniteError):
        nm.NestedSpdReconstruction(torch.from_numpy(w), torch.from_numpy(v), torch.from_numpy(-c), torch.from_numpy(k))


def test_nested_kernel_streams_raw_samples_through_the_tensor_core_projection():
    # hd_gabo_spd.py:239-256: the acquisition screens raw samples of SPD(20) against the training set through the nested
    # kernel; from TENSOR_CORE_ROWS rows the projection runs on the tensor cores (3xTF32, fp32) instead of the fp64 kernel
    import gabotorch_b200 as g
    from oracle import spd as ospd
    rng = np.random.default_rng(3)
    D, d, n_raw, n_train = 20, 5, 20000, 24
    raw = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, n_raw, D, max_cond=50.0)))
    train = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, n_train, D, max_cond=50.0)))
    k = g.NestedSpdAffineInvariantGaussianKernel(D, d, beta_min=0.25)
    k.projection_matrix = torch.from_numpy(onest.grassmann_rand(rng, D, d))
    assert n_raw >= k.TENSOR_CORE_ROWS
    with torch.no_grad():
        auto = k.forward(raw.cuda(), train.cuda())
        k.projection = 'f64'
        exact = k.forward(raw.cuda(), train.cuda())
        k.projection = 'tf32'
        forced = k.forward(raw[:100].cuda(), train.cuda())
    assert float((forced - auto[:100]).abs().max()) <= 5e-6          # same path; fp32 pair kernel, different tile shapes
    m = exact >= 1e-6
    rel = ((auto - exact).abs()[m] / exact[m]).max()
    assert float(rel) <= 2e-4, float(rel)          # fp32 projection: ~1e-6 on the entries, amplified by cond(Y) and beta d^2
    assert torch.equal(torch.argsort(auto.sum(1))[-50:].sort().values, torch.argsort(exact.sum(1))[-50:].sort().values) or \
        float((auto.sum(1) - exact.sum(1)).abs().max()) <= 1e-3


@pytest.mark.parametrize('mode', ['tc', 'bf16'])
def test_projection_opt_in_kernels_parity(mode):
    # the opt-in kernels (GABO_PROJECT_KERNEL, read once per process): 'tc' = tcgen05 / TMEM, same accuracy as the default
    # 3xTF32 kernel; 'bf16' = one TF32 product + ONE bf16 k16 product for both correction terms (1.2e-6 instead of 3e-7 of the
    # output scale, 0.80 instead of 0.74 of HBM); both against the fp64 operator product at 3e-6 of the output scale
    import os, subprocess, sys
    env = dict(os.environ, GABO_PROJECT_KERNEL=mode)
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), '..', 'scripts', 'dev_tc.py'), '--parity-only'],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert 'kernel: ' + mode in out.stdout and 'FAIL' not in out.stdout and out.stdout.count(' OK') >= 7, out.stdout


def test_optimize_reconstruction_parameters_nested_sphere_on_device():
    # nested_spheres_optimization.py:41-98 (hd_gabo_sphere.py:196-199): data generated with known distances-to-axis are
    # recovered; the cost equals the oracle's composition of the reference-pinned inverse chain and sphere distance
    import gabotorch_b200 as g
    from oracle import nested_sphere as onsph, sphere as osph
    rng = np.random.default_rng(12)
    np.random.seed(12)
    D, d, n = 7, 3, 64
    axes = [torch.from_numpy(osph.rand(rng, 1, k)) for k in range(D, d, -1)]
    true_r = [1.3, 0.8, 2.0, 1.1]
    xs = torch.from_numpy(osph.rand(rng, n, d))
    dists = [torch.tensor([[r]], dtype=torch.float64) for r in true_r]
    xd = onsph.projection_from_subsphere_to_sphere(xs, axes, dists)[-1]
    wrong = [torch.tensor([[r + 0.2]], dtype=torch.float64) for r in true_r]
    c_dev = float(g.min_error_reconstruction_cost(xd, xs, axes, wrong))
    xr = onsph.projection_from_subsphere_to_sphere(xs, axes, wrong)[-1]
    c_ref = float((osph.sphere_distance(xd, xr, diag=True) ** 2).sum())
    assert abs(c_dev - c_ref) <= 1e-10 * max(1.0, c_ref)
    out = g.optimize_reconstruction_parameters_nested_sphere(xd, xs, axes, g.TrustRegions(), nb_init_candidates=100)
    got = np.array([float(o) for o in out])
    assert np.abs(got - np.array(true_r)).max() < 2e-3, (got, true_r)
    assert g.optimize_reconstruction_parameters_nested_sphere.last_log['cost'] < 1e-6


@pytest.mark.parametrize('D,n', [(1, 3), (2, 5), (5, 130), (9, 33), (10, 64), (17, 7), (20, 200), (31, 5), (32, 9)])
def test_sym_eig_kernel(D, n):
    # gabo_sym_eig (the batched replacement of the per-matrix torch.symeig calls, spd_utils_torch.py:25,45,110) against
    # LAPACK: eigenvalues, orthonormal vectors, reconstruction; clustered and repeated eigenvalues; non-finite input flagged
    rng = np.random.default_rng(100 * D + n)
    a = rng.standard_normal((n, D, D))
    m = a @ a.transpose(0, 2, 1) + 0.1 * np.eye(D)
    if n > 2:
        m[1] = np.eye(D) * 2.5                                          # all eigenvalues equal
        q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        lam = np.concatenate([np.full(D - D // 2, 1.0), 1.0 + 1e-9 * np.arange(D // 2)])
        m[2] = (q * lam) @ q.T                                          # cluster
    lam_d, vec_d, flag = ops.sym_eig(torch.from_numpy(m))
    assert int(flag.item()) == 0
    lam_d, vec_d = lam_d.cpu().numpy(), vec_d.cpu().numpy()
    want = np.linalg.eigvalsh(m)
    scale = np.abs(want).max(axis=-1, keepdims=True)
    assert np.abs(np.sort(lam_d, axis=-1) - want).max() <= 1e-13 * scale.max()
    eye = np.eye(D)
    assert np.abs(vec_d.transpose(0, 2, 1) @ vec_d - eye).max() < 1e-13
    rec = (vec_d * lam_d[:, None, :]) @ vec_d.transpose(0, 2, 1)
    assert (np.abs(rec - m).max(axis=(1, 2)) <= 1e-13 * scale[:, 0]).all()
    only, none, _ = ops.sym_eig(torch.from_numpy(m), vectors=False)
    assert none is None and np.array_equal(only.cpu().numpy(), lam_d)
    bad = m.copy()
    bad[0, 0, 0] = np.nan
    lam_b, _, flag_b = ops.sym_eig(torch.from_numpy(bad))
    assert int(flag_b.item()) == 1 and np.isnan(lam_b[0].cpu().numpy()).all()
    if n > 1:
        assert np.abs(np.sort(lam_b[1:].cpu().numpy(), axis=-1) - want[1:]).max() <= 1e-13 * scale.max()


def test_nested_spd_reconstruction_costs_on_device():
    # nested_spd_optimization.py:22-92 on the device: values produced by the reference's own functions
    # (tests/golden/make_golden_recon_cost.py; float32 accumulation there), gradients against central differences
    import os
    import gabotorch_b200 as g
    from gabotorch_b200 import nested_optimization as nopt
    gold = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'recon_cost_vectors.npz'))
    for name in ('rc_6_2', 'rc_10_3', 'rc_20_5'):
        args = [torch.from_numpy(gold[name + '_' + k]) for k in ('x', 'y', 'w', 'v', 'c', 'k')]
        ai = g.min_affine_invariant_distance_reconstruction_cost(*args)
        le = g.min_log_euclidean_distance_reconstruction_cost(*args)
        assert ai.is_cuda and ai.dtype == torch.float64
        assert abs(float(ai) - float(gold[name + '_ai'])) <= 5e-6 * float(ai)
        assert abs(float(le) - float(gold[name + '_le'])) <= 5e-6 * float(le)
    rng = np.random.default_rng(3)
    name = 'rc_10_3'
    x, y, w = (torch.from_numpy(gold[name + '_' + k]) for k in ('x', 'y', 'w'))
    for kind in ('affine_invariant', 'log_euclidean'):
        fn = nopt._SpdReconstructionCost(x, y, w, kind)
        params = [torch.from_numpy(gold[name + '_' + k]).cuda().requires_grad_(True) for k in ('v', 'c', 'k')]
        fn(*params).backward()
        for i, p in enumerate(params):
            direction = torch.from_numpy(rng.standard_normal(tuple(p.shape))).cuda()
            if i == 1:
                direction = 0.5 * (direction + direction.T)
            h = 1e-6
            plus = [q.detach() + (h * direction if j == i else 0.0) for j, q in enumerate(params)]
            minus = [q.detach() - (h * direction if j == i else 0.0) for j, q in enumerate(params)]
            fd = (float(fn(*plus)) - float(fn(*minus))) / (2 * h)
            an = float((p.grad * direction).sum())
            assert abs(fd - an) <= 5e-6 * max(1.0, abs(fd)), (kind, i, fd, an)


def test_optimize_reconstruction_parameters_nested_spd_on_device():
    # nested_spd_optimization.py:95-186 (hd_gabo_spd.py:230-233, CG inner solver, log-Euclidean cost): data the mapping can
    # reproduce exactly; the fit ends far below the best random candidate with W^T V = 0, C SPD and |K| < 1
    import gabotorch_b200 as g
    from gabotorch_b200 import nested_optimization as nopt
    rng = np.random.default_rng(21)
    np.random.seed(21)
    D, d, n = 10, 3, 40
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    w, v = q[:, :d].copy(), q[:, d:].copy()
    c = np.diag(np.linspace(1.0, 2.0, D - d))
    k = rng.standard_normal((d, D - d))
    k = 0.5 * k / np.linalg.norm(k)
    ys = []
    for _ in range(n):
        b = rng.standard_normal((d, d))
        ys.append(b @ b.T + 0.5 * np.eye(d))
    y = torch.from_numpy(np.array(ys)).cuda()
    tw, tv, tc, tk = (torch.from_numpy(t).cuda() for t in (w, v, c, k))
    x = nopt._reconstruct_spd(y, nopt._SpectralFn.apply(y, 1), tw, tv, tc, tk)
    # the same map as the reconstruction kernel of nested_mappings (pinned on the reference, test_reconstruction_golden)
    from gabotorch_b200 import nested_mappings as nmap
    x_kernel = nmap.projection_from_nested_spd_to_spd(y, tw, tv, tc, tk)
    assert float((x - x_kernel.to(x.device)).abs().max()) < 1e-11
    for cost_fn in (g.min_affine_invariant_distance_reconstruction_cost, g.min_log_euclidean_distance_reconstruction_cost):
        assert float(cost_fn(x, y, tw, tv, tc, tk)) < 1e-9
        vo, co, ko = g.optimize_reconstruction_parameters_nested_spd(
            x, y, tw, g.ConjugateGradient(maxiter=100), cost_function=cost_fn, nb_init_candidates=100, maxiter=30)
        log = g.optimize_reconstruction_parameters_nested_spd.last_log
        assert not vo.is_cuda and vo.dtype == torch.float64 and tuple(ko.shape) == (d, D - d)
        assert log['cost'] < 0.05 * log['start_cost'], log
        assert float(torch.linalg.norm(vo.T @ torch.from_numpy(w))) < 1e-2, log
        assert float(torch.linalg.eigvalsh(co).min()) > 0 and float(torch.linalg.norm(ko)) < 1.0
        print('nested SPD reconstruction fit (%s): start %.3f -> %.5f in %.2f s, %d outer iterations'
              % (cost_fn.__name__, log['start_cost'], log['cost'], log['time'], log['iterations']))
