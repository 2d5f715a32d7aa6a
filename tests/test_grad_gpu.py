"""Device tests of the autograd surface (SURVEY 8f rank 1 + 2): every kernel class back-propagates to its inputs and to its
manifold-valued parameters like the reference does under torch.autograd (kernels_sphere.py:71-134, kernels_spd.py:72-313,
kernels_nested_spd.py:104-246, kernels_nested_sphere.py:129-152), checked against torch.autograd over the CPU oracle
(restatement of the reference's op sequences, fp64); ``fit_gpytorch_manifold`` (manifold_gp_fit.py:54-222)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gabotorch_b200 as g
from gabotorch_b200 import _lib, ops
from gabotorch_b200 import kernel_utils as ku
from gabotorch_b200 import manifold_gp_fit as mgf
from oracle import nested as onest
from oracle import nested_sphere as onsph
from oracle import spd as ospd
from oracle import sphere as osph


def _close(a, b, rtol, atol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max()
    assert err <= atol + rtol * np.abs(b).max(), 'max err %.3e vs scale %.3e' % (err, np.abs(b).max())


def _oracle_grads(fn, *xs):
    xs = [torch.tensor(np.asarray(x), dtype=torch.float64, requires_grad=True) for x in xs]
    out = fn(*xs)
    return out, xs


def _check_input_grads(kernel, oracle_fn, x1, x2, rtol=2e-5, atol=1e-8, seed=0):
    """sum(K * R) for a fixed random R: gradients with respect to both inputs, device vs oracle autograd."""
    rng = np.random.default_rng(seed)
    a = torch.tensor(x1, dtype=torch.float64, requires_grad=True)
    b = torch.tensor(x2, dtype=torch.float64, requires_grad=True)
    K = kernel.forward(a, b)
    R = torch.from_numpy(rng.standard_normal(tuple(K.shape)))
    (K * R.to(K.device)).sum().backward()
    Ko, (ao, bo) = _oracle_grads(oracle_fn, x1, x2)
    (Ko * R).sum().backward()
    _close(K.detach().cpu(), Ko.detach(), 2e-5, 1e-7)
    _close(a.grad, ao.grad, rtol, atol)
    _close(b.grad, bo.grad, rtol, atol)


def test_sphere_kernels_input_gradients():
    rng = np.random.default_rng(3)
    x, y = osph.rand(rng, 37, 4), osph.rand(rng, 53, 4)
    k = g.SphereGaussianKernel(beta_min=1.0)
    _check_input_grads(k, lambda a, b: osph.sphere_gaussian_kernel(a, b, float(k.beta.detach())), x, y)
    kl = g.SphereLaplaceKernel()
    _check_input_grads(kl, lambda a, b: osph.sphere_laplace_kernel(a, b, float(kl.lengthscale.detach())), x, y)


@pytest.mark.parametrize('d', [2, 3, 5])
def test_spd_kernels_input_gradients(d):
    rng = np.random.default_rng(10 + d)
    X, Y = ospd.spd_sample(rng, 19, d, max_cond=30.0), ospd.spd_sample(rng, 23, d, max_cond=30.0)
    vx = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(X)).numpy()
    vy = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(Y)).numpy()
    unpack = ospd.vector_to_symmetric_matrix_mandel
    kg = g.SpdAffineInvariantGaussianKernel(beta_min=0.3, compute='f64')
    beta = float(kg.beta.detach())
    _check_input_grads(kg, lambda a, b: torch.exp(-beta * ospd.affine_invariant_distance(unpack(a), unpack(b), exact=True) ** 2),
                       vx, vy, rtol=5e-5)
    kl = g.SpdAffineInvariantLaplaceKernel(beta_min=0.3, compute='f64')
    _check_input_grads(kl, lambda a, b: torch.exp(-beta * ospd.affine_invariant_distance(unpack(a), unpack(b), exact=True)),
                       vx, vy, rtol=5e-5)
    kf = g.SpdFrobeniusGaussianKernel()
    ls = float(kf.lengthscale.detach())
    _check_input_grads(kf, lambda a, b: ospd.spd_frobenius_gaussian_kernel(a, b, ls), vx, vy)
    ke = g.SpdLogEuclideanGaussianKernel()
    _check_input_grads(ke, lambda a, b: ospd.spd_log_euclidean_gaussian_kernel(a, b, ls), vx, vy, rtol=5e-5)


def _nested_spd_setup(D=6, d=2, n1=15, n2=11, seed=5):
    rng = np.random.default_rng(seed)
    X, Y = ospd.spd_sample(rng, n1, D, max_cond=30.0), ospd.spd_sample(rng, n2, D, max_cond=30.0)
    vx = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(X)).numpy()
    vy = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(Y)).numpy()
    W = onest.grassmann_rand(rng, D, d)
    return vx, vy, np.asarray(W, dtype=np.float64)


def test_nested_spd_kernels_input_and_projection_gradients():
    D, d = 6, 2
    vx, vy, W = _nested_spd_setup(D, d)
    for cls, oracle in ((g.NestedSpdAffineInvariantGaussianKernel, 'ai'), (g.NestedSpdLogEuclideanGaussianKernel, 'le')):
        k = cls(D, d, beta_min=0.3, compute='f64') if oracle == 'ai' else cls(D, d)
        k.raw_projection_matrix.data = torch.from_numpy(W).to(k.raw_projection_matrix.dtype)
        Wp = k.raw_projection_matrix.detach().double().numpy()        # the fp32-rounded matrix the kernel really uses
        scal = float(k.beta.detach()) if oracle == 'ai' else float(k.lengthscale.detach())
        rng = np.random.default_rng(1)
        a = torch.tensor(vx, requires_grad=True)
        b = torch.tensor(vy, requires_grad=True)
        K = k.forward(a, b)
        R = torch.from_numpy(rng.standard_normal(tuple(K.shape)))
        k.raw_projection_matrix.grad = None
        (K * R.to(K.device)).sum().backward()
        ao = torch.tensor(vx, requires_grad=True)
        bo = torch.tensor(vy, requires_grad=True)
        wo = torch.tensor(Wp, requires_grad=True)
        if oracle == 'ai':
            m1 = onest.projection_from_spd_to_nested_spd(ospd.vector_to_symmetric_matrix_mandel(ao), wo)
            m2 = onest.projection_from_spd_to_nested_spd(ospd.vector_to_symmetric_matrix_mandel(bo), wo)
            Ko = torch.exp(-scal * ospd.affine_invariant_distance(m1, m2, exact=True) ** 2)
        else:
            Ko = onest.nested_spd_log_euclidean_gaussian_kernel(ao, bo, wo, scal)
        (Ko * R).sum().backward()
        _close(K.detach().cpu(), Ko.detach(), 2e-5, 1e-7)
        _close(a.grad, ao.grad, 1e-4, 1e-8)
        _close(b.grad, bo.grad, 1e-4, 1e-8)
        _close(k.raw_projection_matrix.grad.double(), wo.grad, 1e-4, 1e-6)   # gradient stored in the parameter's fp32


def test_nested_sphere_kernel_gradients_and_forward_agreement():
    rng = np.random.default_rng(8)
    D, d = 6, 3
    x, y = osph.rand(rng, 21, D), osph.rand(rng, 17, D)
    k = g.NestedSphereGaussianKernel(D, d, beta_min=1.0)
    axes = [a.detach().double() for a in k.axes]
    dists = [t.double() for t in k.distances_to_axis]
    beta = float(k.beta.detach())
    # the differentiable chain equals the fused projection kernel
    with torch.no_grad():
        p_kernel = ops.nested_sphere_project(torch.from_numpy(x), axes, dists).cpu()
        p_chain = ku._nested_sphere_project_autograd(torch.from_numpy(x), axes, dists).cpu()
    _close(p_chain, p_kernel, 1e-9, 1e-10)
    a = torch.tensor(x, requires_grad=True)
    b = torch.tensor(y, requires_grad=True)
    K = k.forward(a, b)
    R = torch.from_numpy(rng.standard_normal(tuple(K.shape)))
    (K * R.to(K.device)).sum().backward()
    ao = torch.tensor(x, requires_grad=True)
    bo = torch.tensor(y, requires_grad=True)
    axo = [torch.tensor(t.numpy(), requires_grad=True) for t in axes]
    Ko = onsph.nested_sphere_gaussian_kernel(ao, bo, axo, dists, beta)
    (Ko * R).sum().backward()
    _close(K.detach().cpu(), Ko.detach(), 2e-5, 1e-7)
    _close(a.grad, ao.grad, 1e-4, 1e-8)
    _close(b.grad, bo.grad, 1e-4, 1e-8)
    for p, o in zip(k.axes, axo):
        _close(p.grad.double().reshape(-1), o.grad.reshape(-1), 2e-4, 1e-6)


def test_weighted_points_sum_and_logm_backward_building_blocks():
    rng = np.random.default_rng(2)
    gmat = torch.from_numpy(rng.standard_normal((70, 45)))
    b = torch.from_numpy(rng.standard_normal((45, 9)))
    a = torch.from_numpy(rng.standard_normal((70, 5)))
    _close(ops.weighted_points_sum(gmat, b).cpu(), gmat @ b, 1e-12, 1e-12)
    _close(ops.weighted_points_sum(gmat, a, transpose=True).cpu(), gmat.T @ a, 1e-12, 1e-12)
    for d in (1, 2, 4, 8):
        X = torch.from_numpy(ospd.spd_sample(rng, 12, d, max_cond=50.0))
        G = torch.from_numpy(rng.standard_normal((12, d, d)))
        Xo = X.clone().requires_grad_(True)
        (ospd.logm(0.5 * (Xo + Xo.transpose(-1, -2))) * G).sum().backward()
        got = ops.spd_logm_backward(X, G).cpu()
        _close(got, 0.5 * (Xo.grad + Xo.grad.transpose(-1, -2)), 1e-8, 1e-10)


def _make_nested_model(D=6, d=2, n=28, seed=11):
    rng = np.random.default_rng(seed)
    X = ospd.spd_sample(rng, n, D, max_cond=20.0)
    vx = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(X))
    w_true = np.asarray(onest.grassmann_rand(rng, D, d))
    y_lat = np.einsum('da,ndk,kb->nab', w_true, X, w_true)
    y = np.log(np.linalg.det(y_lat)) + 0.3 * np.trace(y_lat, axis1=1, axis2=2)      # depends on W^T X W only
    y = (y - y.mean()) / y.std()
    base = g.NestedSpdAffineInvariantGaussianKernel(D, d, beta_min=0.2, compute='f64')
    model = g.ManifoldGP(vx, torch.from_numpy(y), g.ScaleKernel(base), noise=0.1)
    return model, base


def test_manifold_objective_gradient_matches_finite_differences():
    model, base = _make_nested_model()
    # the reference keeps the projection matrix in fp32; for a finite-difference check hold it in fp64 so that the step
    # h is not swamped by the parameter's own rounding
    base.raw_projection_matrix.data = base.raw_projection_matrix.data.double()
    obj = mgf.ManifoldObjective(model)
    x = obj.current()
    f, grads = obj.cost_grad(x)
    assert abs(f - obj.cost(x)) <= 1e-9 * max(1.0, abs(f))
    man = obj.manifold
    rg = man.egrad2rgrad(x, grads)
    np.random.seed(0)
    for _ in range(3):
        u = man.proj(x, man.rand())
        u = [ui / man.norm(x, u) for ui in u]
        h = 2e-3     # the distances come back rounded to fp32 (spd_utils_torch.py:108-120): the objective carries ~1e-7 noise
        fp = obj.cost(man.retr(x, [h * ui for ui in u]))
        fm = obj.cost(man.retr(x, [-h * ui for ui in u]))
        fd = (fp - fm) / (2 * h)
        an = man.inner(x, rg, u)
        assert abs(fd - an) <= 3e-3 * max(abs(an), 1e-2), (fd, an)


def test_fit_gpytorch_manifold_runs_and_improves():
    np.random.seed(4)
    torch.manual_seed(4)
    model, base = _make_nested_model()
    obj0 = mgf.ManifoldObjective(model)
    f0 = obj0.cost(obj0.current())
    mll = g.ExactMarginalLogLikelihood(None, model)
    out, info = g.fit_gpytorch_manifold(mll, solver=g.ConjugateGradient(maxiter=40), nb_init_candidates=24)
    assert out is mll and set(info) >= {'fopt', 'wall_time', 'opt_log', 'iterations'}
    assert info['fopt'] <= f0 + 1e-9                             # never worse than the starting parameters
    W = base.raw_projection_matrix.detach().double()
    assert float((W.T @ W - torch.eye(W.shape[1], dtype=torch.float64)).abs().max()) < 1e-5   # still on the Grassmannian
    assert model.noise > 0 and float(base.beta.detach()) >= 0.2
    # the fitted model is usable by the acquisition path (projection + Gram + factorisation)
    K = base.forward(model.train_inputs[0], model.train_inputs[0])
    assert torch.isfinite(K).all()


def test_fit_gpytorch_manifold_nested_sphere_axes():
    np.random.seed(6)
    rng = np.random.default_rng(6)
    D, d, n = 5, 3, 30
    x = osph.rand(rng, n, D)
    y = np.sin(3 * x[:, 0]) + x[:, 1] ** 2
    base = g.NestedSphereGaussianKernel(D, d, beta_min=1.0)
    model = g.ManifoldGP(torch.from_numpy(x), torch.from_numpy((y - y.mean()) / y.std()), g.ScaleKernel(base), noise=0.1)
    obj0 = mgf.ManifoldObjective(model)
    f0 = obj0.cost(obj0.current())
    _, info = g.fit_gpytorch_manifold(model, solver=g.ConjugateGradient(maxiter=25), nb_init_candidates=12)
    assert info['fopt'] <= f0 + 1e-9
    for a in base.axes:
        assert abs(float(a.detach().double().norm()) - 1.0) < 1e-5
