"""The reference-facing classes and functions (same names / signatures as BoManifolds) running on the device."""
import math

import numpy as np
import pytest
import torch

import gabotorch_b200 as g
from gabotorch_b200 import manifold_optimization as mo
from gabotorch_b200 import riemannian_utils as ru
from oracle import gp as ogp
from oracle import spd as ospd
from oracle import sphere as osph

pytestmark = pytest.mark.gpu


def test_sphere_kernel_class_like_the_reference_example(golden):
    # examples/kernels/sphere/sphere_gaussian_kernel_parameters.py:91-95
    x = torch.from_numpy(golden['s2_n256_x'])
    k = g.SphereGaussianKernel(beta_min=0.0)
    k.beta = float(golden['s2_n256_beta'])
    with torch.no_grad():
        K = k.forward(x, x)
    assert K.dtype == torch.float64 and not K.is_cuda and tuple(K.shape) == (256, 256)
    st = int(golden['s2_n256_stride'])
    np.testing.assert_allclose(K.numpy()[::st, ::st], golden['s2_n256_k'], rtol=1e-5, atol=1e-11)
    with torch.no_grad():
        Kd = k(x, x, diag=True)
    assert tuple(Kd.shape) == (256, 1)
    # device tensors stay on the device
    with torch.no_grad():
        Kc = k.forward(x.cuda(), x.cuda())
    assert Kc.is_cuda
    # autograd w.r.t. the kernel parameter (what GP fitting needs)
    K2 = k.forward(x[:32], x[:32])
    K2.sum().backward()
    d = osph.sphere_distance(x[:32], x[:32])
    beta = float(k.beta.detach())
    dK_dbeta = float((-(d * d) * torch.exp(-beta * d * d)).sum())
    sig = 1.0 / (1.0 + math.exp(-float(k.raw_beta.detach())))
    assert abs(float(k.raw_beta.grad) - dK_dbeta * sig) <= 1e-4 * abs(dK_dbeta * sig)
    lap = g.SphereLaplaceKernel()
    lap.lengthscale = 0.8
    with torch.no_grad():
        np.testing.assert_allclose(lap.forward(x[:50], x[:60]).numpy(),
                                   osph.sphere_laplace_kernel(x[:50], x[:60], float(lap.lengthscale.detach())).numpy(),
                                   rtol=1e-5)


def test_spd_kernel_classes(golden):
    v = torch.from_numpy(golden['spd3_n128_vec'])
    beta = float(golden['spd3_n128_beta'])
    k = g.SpdAffineInvariantGaussianKernel(beta_min=0.5, compute='f64')
    assert abs(float(k.beta.detach()) - beta) < 1e-6              # raw_beta = 0 -> beta_min + ln 2
    with torch.no_grad():
        K = k.forward(v, v)
        assert tuple(k.forward(v, v, diagonal_distance=True).shape) == (128, 1)
        sk = g.ScaleKernel(k)
        sk.outputscale = 2.0
        Ks = sk(v, v)
    np.testing.assert_allclose(K.numpy(), golden['spd3_n128_k'], rtol=2e-5, atol=1e-11)
    np.testing.assert_allclose(Ks.numpy(), 2.0 * K.numpy(), rtol=1e-6)
    with torch.no_grad():
        lap = g.SpdAffineInvariantLaplaceKernel(beta_min=0.5)
        np.testing.assert_allclose(lap.forward(v, v).numpy(), np.exp(-beta * golden['spd3_n128_d']), rtol=1e-5)
        fro = g.SpdFrobeniusGaussianKernel()
        fro.lengthscale = 1.3
        ls = float(fro.lengthscale.detach())                      # float32 parameter, .double()-ed like the reference
        np.testing.assert_allclose(fro.forward(v, v).numpy(), ospd.spd_frobenius_gaussian_kernel(v, v, ls).numpy(),
                                   rtol=1e-9)
        le = g.SpdLogEuclideanGaussianKernel()
        le.lengthscale = 1.3
        np.testing.assert_allclose(le.forward(v, v).numpy(),
                                   ospd.spd_log_euclidean_gaussian_kernel(v, v, ls).numpy(), rtol=1e-8, atol=1e-12)
    vg = v.clone().requires_grad_(True)                            # every kernel back-propagates to its inputs
    lap.forward(vg, v.clone()).sum().backward()                    # (values checked in tests/test_grad_gpu.py)
    assert vg.grad is not None and torch.isfinite(vg.grad).all()


def test_riemannian_utils_functions(golden):
    a, b = torch.from_numpy(golden['spd3_rect_a']), torch.from_numpy(golden['spd3_rect_b'])
    d = ru.affine_invariant_distance_torch(a, b)
    assert d.dtype == torch.float64 and tuple(d.shape) == (19, 27)
    np.testing.assert_allclose(d.numpy(), golden['spd3_rect_d'], rtol=1e-5, atol=1e-6)
    z = ru.affine_invariant_distance_torch(a, a, diagonal_distance=True)
    assert tuple(z.shape) == (19, 1) and float(z.abs().max()) == 0
    v = ru.symmetric_matrix_to_vector_mandel_torch(a)
    assert torch.equal(ru.vector_to_symmetric_matrix_mandel_torch(v), 0.5 * (a + a.transpose(-1, -2))) or \
        float((ru.vector_to_symmetric_matrix_mandel_torch(v) - a).abs().max()) < 1e-15
    s = torch.from_numpy(golden['s3_rect_a'])
    np.testing.assert_allclose(ru.sphere_distance_torch(s, torch.from_numpy(golden['s3_rect_b'])).numpy(),
                               golden['s3_rect_d'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ru.logm_torch(a[0]).numpy(), ospd.logm(a[0]).numpy(), atol=1e-10)


def _sphere_bo_setup(D=6, n=32, noise=1e-2, seed=1234):
    rng = np.random.default_rng(seed)
    xt = osph.rand(rng, n, D)
    y = osph.ackley(xt)
    base = g.SphereGaussianKernel(beta_min=1.0)
    model = g.ManifoldGP(torch.from_numpy(xt), torch.from_numpy(y), g.ScaleKernel(base), noise=noise)
    model.covar_module.outputscale = 1.0
    acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
    ogp_ = ogp.make_gp('sphere', xt, y, beta=float(base.beta.detach()), outputscale=1.0, noise=noise)
    return acq, ogp_


def test_joint_optimize_manifold_sphere():
    acq, gp = _sphere_bo_setup()
    man = g.Sphere(6)
    torch.manual_seed(0)
    ics = g.gen_batch_initial_conditions_manifold(acq, man, None, q=1, num_restarts=64, raw_samples=512,
                                                  options={'seed': 3})
    assert tuple(ics.shape) == (64, 1, 6)
    ei0 = acq(ics)
    np.testing.assert_allclose(ei0.cpu().numpy(), ogp.ei_batch(gp, ics[:, 0].cpu().numpy()), rtol=2e-4,
                               atol=1e-6 * float(ei0.max()))
    cands, vals = g.gen_candidates_manifold(ics, acq, man, g.ConjugateGradient(maxiter=100))
    assert tuple(cands.shape) == (64, 1, 6) and tuple(vals.shape) == (64,)
    assert torch.all(vals >= ei0 * (1 - 1e-5) - 1e-12)
    best = g.get_best_candidates(cands, vals)
    assert tuple(best.shape) == (1, 6) and float(acq(best[None])[0]) == pytest.approx(float(vals.max()), rel=1e-4)
    new_x = g.joint_optimize_manifold(acq, man, g.ConjugateGradient(maxiter=100), q=1, num_restarts=32,
                                      raw_samples=256, bounds=torch.stack([-torch.ones(6), torch.ones(6)]),
                                      options={'seed': 3})
    assert tuple(new_x.shape) == (1, 6) and abs(float(new_x.norm()) - 1) < 1e-12
    assert ogp.ei_and_grad(gp, new_x[0].cpu().numpy(), False)[0] >= 0.99 * float(ei0.max())
    with pytest.raises(NotImplementedError):
        g.gen_candidates_manifold(ics, acq, man, g.ConjugateGradient(), inequality_constraints=[lambda x: x])


def test_joint_optimize_manifold_spd_with_mandel_processing():
    # gabo_spd.py:200-203: candidates live in Mandel notation for the GP and as matrices for the solver
    rng = np.random.default_rng(4)
    d, n = 3, 24
    xt = ospd.spd_sample(rng, n, d, max_cond=100.0)
    xv = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt))
    y = ospd.ackley(xv)
    base = g.SpdAffineInvariantGaussianKernel(beta_min=0.5)
    model = g.ManifoldGP(xv, torch.from_numpy(y), g.ScaleKernel(base), noise=1e-2)
    model.covar_module.outputscale = 1.0
    acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
    man = g.PositiveDefinite(d)
    man.min_eig, man.max_eig = 0.001, 5.0
    new_x = g.joint_optimize_manifold(acq, man, g.ConjugateGradient(maxiter=50), q=1, num_restarts=16,
                                      raw_samples=128, bounds=None, options={'seed': 1},
                                      pre_processing_manifold=ru.vector_to_symmetric_matrix_mandel_torch,
                                      post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch)
    assert tuple(new_x.shape) == (1, 6)
    m = ospd.vector_to_symmetric_matrix_mandel(new_x.cpu()).numpy()[0]
    assert np.linalg.eigvalsh(m).min() > 0
    gp = ogp.make_gp('spd', xt, y, beta=float(base.beta.detach()), outputscale=1.0, noise=1e-2)
    ei_new = ogp.ei_and_grad(gp, m, False)[0]
    assert ei_new == pytest.approx(float(acq(new_x[None])[0]), rel=2e-3)
    raw = man.rand_batch(128)
    assert ei_new >= float(acq(ru.symmetric_matrix_to_vector_mandel_torch(raw)[:, None]).max()) * 0.5


@pytest.mark.parametrize('kernel', ['gauss', 'laplace'])
def test_sphere_kernels_backpropagate_to_their_inputs(kernel):
    # the reference differentiates acos(clamp(<x1, x2>)) with torch.autograd (sphere_utils_torch.py:29-55); the fused
    # path must give the same input gradients, including zero gradient where the clamp is active (identical points)
    rng = np.random.default_rng(11)
    a, b = osph.rand(rng, 17, 5), osph.rand(rng, 23, 5)
    b[0] = a[0]                                                  # clamp active for the pair (0, 0)
    wts = torch.from_numpy(rng.standard_normal((17, 23)))
    x1 = torch.from_numpy(a).clone().requires_grad_(True)
    x2 = torch.from_numpy(b).clone().requires_grad_(True)
    if kernel == 'gauss':
        k = g.SphereGaussianKernel(beta_min=1.0)
        ref_fn = lambda u, v: osph.sphere_gaussian_kernel(u, v, float(k.beta.detach()))
    else:
        k = g.SphereLaplaceKernel()
        k.lengthscale = 0.8
        ref_fn = lambda u, v: osph.sphere_laplace_kernel(u, v, 0.8)
    out = k.forward(x1, x2)
    (out * wts).sum().backward()
    r1 = torch.from_numpy(a).clone().requires_grad_(True)
    r2 = torch.from_numpy(b).clone().requires_grad_(True)
    ref = ref_fn(r1, r2)
    (ref * wts).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-7)
    scale = float(r1.grad.abs().max())
    np.testing.assert_allclose(x1.grad.numpy(), r1.grad.numpy(), rtol=0, atol=2e-5 * scale)
    np.testing.assert_allclose(x2.grad.numpy(), r2.grad.numpy(), rtol=0, atol=2e-5 * scale)


# f64: limited by the forward distance, whose eigenvalues pass through float32 like the reference's (spd_utils_torch.py:108)
@pytest.mark.parametrize('d,compute,tol', [(3, 'f64', 5e-6), (3, 'f32', 2e-4), (5, 'f64', 5e-6), (8, 'f32', 5e-4)])
def test_spd_kernel_backpropagates_to_its_inputs(d, compute, tol):
    # the reference differentiates cholesky / inverse / bmm / symeig with torch.autograd (spd_utils_torch.py:87-120);
    # here: exact-fp64 restatement of the same function under autograd vs the fused forward + log-map backward kernel
    rng = np.random.default_rng(40 + d)
    n1, n2 = 19, 27
    a = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, n1, d, max_cond=50.0)))
    b = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, n2, d, max_cond=50.0)))
    wts = torch.from_numpy(rng.standard_normal((n1, n2)))
    k = g.SpdAffineInvariantGaussianKernel(beta_min=0.3, compute=compute)
    beta = float(k.beta.detach())
    x1 = a.clone().requires_grad_(True)
    x2 = b.clone().requires_grad_(True)
    out = k.forward(x1, x2)
    (out * wts).sum().backward()
    r1 = a.clone().requires_grad_(True)
    r2 = b.clone().requires_grad_(True)
    dist = ospd.affine_invariant_distance(ospd.vector_to_symmetric_matrix_mandel(r1),
                                          ospd.vector_to_symmetric_matrix_mandel(r2), exact=True)
    ref = torch.exp(-dist * dist * beta)
    (ref * wts).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=max(tol, 1e-5), atol=1e-9)
    for got, want in ((x1.grad, r1.grad), (x2.grad, r2.grad)):
        scale = float(want.abs().max())
        np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=tol * scale)
