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
    # beta / lengthscale stay differentiable (GP hyper-parameter fitting), the projection matrix does not
    out = k.forward(x1, x2)
    out.sum().backward()
    assert k.raw_beta.grad is not None and torch.isfinite(k.raw_beta.grad).all()
    k.raw_projection_matrix.requires_grad_(True)
    with pytest.raises(NotImplementedError):
        k.forward(x1, x2)


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
