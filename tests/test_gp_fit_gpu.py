"""GP marginal likelihood / hyper-parameter fit: ``gabo_gp_mll`` against torch.distributions +
autograd (oracle/gp.py), and the fit against scipy L-BFGS-B run on the oracle objective.  gpytorch / botorch are absent
(PARITY UNPINNED for the optimiser trajectory); tolerances: log-likelihood 1e-9 relative, gradient 1e-7 relative to
its largest component (fp64 Cholesky of matrices with condition numbers up to ~1e8)."""
import math

import numpy as np
import pytest
import torch

import gabotorch_b200 as g
from gabotorch_b200 import gp_fit, ops
from oracle import gp as ogp
from oracle import spd as ospd
from oracle import sphere as osph

pytestmark = pytest.mark.gpu


def _sphere_problem(n, dim, seed):
    rng = np.random.default_rng(seed)
    x = osph.rand(rng, n, dim)
    y = osph.ackley(x) if hasattr(osph, 'ackley') else np.sin(3 * x[:, 0]) + x[:, 1] ** 2
    return x, np.asarray(y, dtype=np.float64).reshape(-1)


@pytest.mark.parametrize('n', [1, 2, 5, 33, 64, 128])
def test_mll_and_gradient_vs_oracle(n):
    x, y = _sphere_problem(n, 4, n)
    d = osph.sphere_distance(torch.from_numpy(x), torch.from_numpy(x)).numpy()
    dm = d * d
    thetas = np.array([[2.0 + math.log(2), 1.0, 2.0, y.mean()], [3.5, 0.3, 1e-2, 0.1], [2.1, 7.0, 1e-4, -0.4],
                       [9.0, 1.7, 0.5, 2.0], [2.0, 0.05, 1e-6, 0.0]])
    ll, grad, alpha, kinv, flags = ops.gp_mll(torch.from_numpy(dm), torch.from_numpy(y), torch.from_numpy(thetas),
                                              want_grad=True, want_factors=True)
    assert not flags.cpu().numpy().any()
    for b, th in enumerate(thetas):
        ref_ll, ref_g = ogp.exact_log_likelihood(dm, y, th)
        assert abs(float(ll[b]) - ref_ll) <= 1e-9 * max(1.0, abs(ref_ll))
        np.testing.assert_allclose(grad[b].cpu().numpy(), ref_g, rtol=0, atol=1e-7 * np.abs(ref_g).max() + 1e-10)
        k = th[1] * np.exp(-th[0] * dm) + th[2] * np.eye(n)
        kin = np.linalg.inv(k)
        np.testing.assert_allclose(kinv[b].cpu().numpy(), kin, rtol=0, atol=1e-8 * np.abs(kin).max())
        np.testing.assert_allclose(alpha[b].cpu().numpy(), kin @ (y - th[3]), rtol=0,
                                   atol=1e-8 * np.abs(kin @ (y - th[3])).max() + 1e-12)


def test_mll_flags_non_positive_definite_and_rejects_bad_sizes():
    x, y = _sphere_problem(12, 3, 0)
    d = osph.sphere_distance(torch.from_numpy(x), torch.from_numpy(x)).numpy()
    thetas = np.array([[6.5, 1.0, 0.1, 0.0], [6.5, 1.0, -3.0, 0.0]])
    ll, grad, _, _, flags = ops.gp_mll(torch.from_numpy(d * d), torch.from_numpy(y), torch.from_numpy(thetas))
    assert flags.cpu().tolist() == [0, 1]
    assert np.isfinite(float(ll[0])) and np.isnan(float(ll[1])) and np.isnan(grad[1].cpu().numpy()).all()
    with pytest.raises(g.GaboError):
        ops.gp_mll(torch.zeros(129, 129, dtype=torch.float64), torch.zeros(129, dtype=torch.float64),
                   torch.tensor([[1.0, 1.0, 1.0, 0.0]], dtype=torch.float64))
    with pytest.raises(ValueError):
        ops.gp_mll(torch.zeros(4, 5, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                   torch.tensor([[1.0, 1.0, 1.0, 0.0]], dtype=torch.float64))


def test_objective_in_raw_parameters_vs_oracle_spd():
    # SPD(3) inputs in Mandel notation, Laplace and Gaussian kernels, priors of gabo_spd.py:165-176
    rng = np.random.default_rng(5)
    xm = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(ospd.spd_sample(rng, 17, 3, max_cond=100.0)))
    y = rng.standard_normal(17)
    dist = ospd.affine_invariant_distance(ospd.vector_to_symmetric_matrix_mandel(xm),
                                          ospd.vector_to_symmetric_matrix_mandel(xm), exact=True).numpy()
    for cls, power in ((g.SpdAffineInvariantGaussianKernel, 2), (g.SpdAffineInvariantLaplaceKernel, 1)):
        dm, p = gp_fit.kernel_distance_matrix(cls(beta_min=0.5), xm)
        assert p == power
        ref_dm = dist ** power
        np.testing.assert_allclose(dm.cpu().numpy(), ref_dm, rtol=0, atol=1e-6 * ref_dm.max() + 1e-12)
        obj = gp_fit.MarginalLogLikelihood(dm, y, 0.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))
        for raw in ([0.0, 0.0, 3.0, 0.2], [1.5, -2.0, -4.0, -1.0]):
            val, grad = obj(np.array(raw))
            ref = ogp.mll_objective(dm.cpu().numpy(), y, raw, 0.5, outputscale_prior=(2.0, 0.15),
                                    noise_prior=(1.1, 0.05))
            assert abs(val - ref) <= 1e-9 * max(1.0, abs(ref))
            num = np.zeros(4)
            for i in range(4):
                e = np.zeros(4)
                e[i] = 1e-5
                num[i] = (ogp.mll_objective(dm.cpu().numpy(), y, np.array(raw) + e, 0.5, outputscale_prior=(2.0, 0.15),
                                            noise_prior=(1.1, 0.05))
                          - ogp.mll_objective(dm.cpu().numpy(), y, np.array(raw) - e, 0.5,
                                              outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))) / 2e-5
            np.testing.assert_allclose(grad, num, rtol=0, atol=1e-6 * max(1.0, np.abs(num).max()))
        vals = obj.batch_values(np.array([[0.0, 0.0, 3.0, 0.2], [1.5, -2.0, -4.0, -1.0]]))
        assert abs(vals[1] - ref) <= 1e-9 * max(1.0, abs(ref))


@pytest.mark.parametrize('optimizer', ['device', 'scipy'])
@pytest.mark.parametrize('raw_samples', [0, 64])
def test_fit_matches_scipy_on_the_oracle_objective(raw_samples, optimizer):
    # the model of gabo_sphere.py:131-147: ScaleKernel(SphereGaussianKernel) with Gamma(2, 0.15) on the outputscale,
    # Gamma(1.1, 0.05) on the noise starting at its mode, constant mean
    from scipy.optimize import minimize
    x, y = _sphere_problem(30, 3, 11)
    beta_min = 6.5
    cov = g.ScaleKernel(g.SphereGaussianKernel(beta_min=beta_min), outputscale_prior=g.GammaPrior(2.0, 0.15))
    noise_prior = g.GammaPrior(1.1, 0.05)
    mode = float((noise_prior.concentration - 1) / noise_prior.rate)
    model = g.ManifoldGP(x, y, cov, noise=mode, mean=0.0, noise_prior=noise_prior)
    mll = g.ExactMarginalLogLikelihood(model.likelihood, model)
    d = osph.sphere_distance(torch.from_numpy(x), torch.from_numpy(x)).numpy()
    dm = d * d

    def f(raw):
        return ogp.mll_objective(dm, y, raw, beta_min, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))

    raw0 = np.array([0.0, 0.0, gp_fit._inv_softplus(mode - 1e-8), 0.0])
    start = f(raw0)
    out = g.fit_gpytorch_model(mll=mll, raw_samples=raw_samples, generator=3, optimizer=optimizer)
    assert model.fit_result['optimizer'] == optimizer
    assert out is mll
    res = model.fit_result
    beta, s, noise, mean = res['theta']
    assert res['objective'] < start and beta >= beta_min and s > 0 and noise > 1e-8
    assert abs(float(cov.base_kernel.beta.detach()) - beta) <= 1e-6 * beta
    assert abs(float(cov.outputscale.detach()) - s) <= 1e-6 * s
    assert model.noise == noise and model.mean == mean
    # the device objective at the fitted point equals the oracle's, and no L-BFGS run on the oracle objective from the
    # same start finds a lower value (same optimiser as botorch's fit_gpytorch_scipy)
    raw_fit = np.array([gp_fit._inv_softplus(beta - beta_min), gp_fit._inv_softplus(s),
                        gp_fit._inv_softplus(noise - 1e-8), mean])
    assert abs(f(raw_fit) - res['objective']) <= 1e-8 * max(1.0, abs(res['objective']))
    ref = minimize(f, raw0, method='L-BFGS-B', options={'maxiter': 500})
    assert res['objective'] <= ref.fun + 1e-6
    if raw_samples == 0:
        assert abs(res['objective'] - ref.fun) <= 1e-5
    # the fitted model drives the acquisition as before
    ei = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
    vals = ei(torch.from_numpy(osph.rand(np.random.default_rng(0), 50, 3))[:, None, :])
    assert vals.shape == (50,) and torch.isfinite(vals).all() and (vals >= 0).all()


def test_device_fit_many_starts_in_one_launch():
    # gabo_gp_fit: one CTA per start; every start ends at a stationary point of the oracle objective, the best is kept
    x, y = _sphere_problem(32, 3, 21)
    d = osph.sphere_distance(torch.from_numpy(x), torch.from_numpy(x)).numpy()
    dm = d * d
    rng = np.random.default_rng(0)
    starts = np.vstack([[0.0, 0.0, 3.0, 0.0], rng.standard_normal((15, 4)) * np.array([2.0, 2.0, 3.0, 0.5])])
    pri = [0.0, 0.0, 2.0, 0.15, 1.1, 0.05]
    raws, fs, info = ops.gp_fit(torch.from_numpy(dm), torch.from_numpy(y), torch.from_numpy(starts), 6.5, 1e-8, pri,
                                [0, 0, 0, 0])
    assert raws.shape == (16, 4) and np.isfinite(fs).all() and (info[:, 0] <= 1).all() and (info[:, 1] >= 1).all()
    for raw, f in zip(raws, fs):
        ref = ogp.mll_objective(dm, y, raw, 6.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))
        assert abs(ref - f) <= 1e-9 * max(1.0, abs(ref))
        num = np.array([(ogp.mll_objective(dm, y, raw + e, 6.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))
                         - ogp.mll_objective(dm, y, raw - e, 6.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05)))
                        / 2e-5 for e in 1e-5 * np.eye(4)])
        assert np.abs(num).max() <= 5e-4                      # stationary (pgtol 1e-5 or flat to ftol)
    f0 = [ogp.mll_objective(dm, y, s0, 6.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05)) for s0 in starts]
    assert (fs <= np.array(f0) + 1e-12).all()
    # a fixed parameter stays where it started
    raws2, _, _ = ops.gp_fit(torch.from_numpy(dm), torch.from_numpy(y), torch.from_numpy(starts[:3]), 6.5, 1e-8, pri,
                             [0, 1, 0, 0])
    np.testing.assert_array_equal(raws2[:, 1], starts[:3, 1])


def test_fit_unscaled_kernel_and_unsupported_kernel():
    x, y = _sphere_problem(15, 5, 2)
    model = g.ManifoldGP(x, y, g.SphereGaussianKernel(beta_min=1.2), noise=0.5)
    g.fit_gpytorch_model(model)
    assert model.fit_result['theta'][1] == 1.0 and model.fit_result['evaluations'] >= 2
    with pytest.raises(NotImplementedError):
        g.fit_gpytorch_model(g.ManifoldGP(x, y, g.SphereLaplaceKernel(), noise=0.5))


@pytest.mark.parametrize('n', [1, 7, 32, 100, 128])
def test_gp_factor_vs_numpy(n):
    # the factors build_device_gp hands to the acquisition kernels: alpha and (s K + noise I)^-1 from a Gram matrix
    x, y = _sphere_problem(n, 5, 40 + n)
    k = ops.sphere_gram(torch.from_numpy(x), torch.from_numpy(x), 1.2 + math.log(2), kind=0)
    kn = 0.7 * k.cpu().numpy() + 1e-2 * np.eye(n)
    kn = np.tril(kn) + np.tril(kn, -1).T
    alpha, kinv = ops.gp_factor(k, torch.from_numpy(y), 0.7, 1e-2, 0.3)
    ref = np.linalg.inv(kn)
    np.testing.assert_allclose(kinv.cpu().numpy(), ref, rtol=0, atol=1e-9 * np.abs(ref).max())
    np.testing.assert_allclose(alpha.cpu().numpy(), ref @ (y - 0.3), rtol=0, atol=1e-9 * np.abs(ref @ (y - 0.3)).max())
    np.testing.assert_array_equal(kinv.cpu().numpy(), kinv.cpu().numpy().T)
    with pytest.raises(ops.NotPositiveDefiniteError):
        ops.gp_factor(k, torch.from_numpy(y), 0.7, -5.0, 0.3)
