"""Oracle (CPU, fp64) for the acquisition value on the hot path.  Test infrastructure only.

The reference evaluates ``botorch.acquisition.ExpectedImprovement(model, best_f, maximize=False)`` over a
``botorch.models.SingleTaskGP`` with ``ScaleKernel(SphereGaussianKernel | SpdAffineInvariantGaussianKernel)``
and a constant mean (call sites ``examples/bo_sphere/benchmark_examples/gabo_sphere.py:131-165``,
``examples/bo_spd/benchmark_examples/gabo_spd.py:165-197``).  botorch / gpytorch are third-party and absent
from ``/root/reference``: PARITY UNPINNED.  Restated from the published definitions:

    k_i   = s * exp(-beta d(x, X_i)^2)                     (ScaleKernel o geodesic Gaussian kernel)
    mu    = m + k^T alpha,   alpha = (K + noise I)^-1 (y - m)
    var   = max(s * k(x,x) - k^T (K + noise I)^-1 k, 1e-9)  (botorch: variance.clamp_min(1e-9))
    u     = (best_f - mu) / sigma                           (maximize=False)
    EI    = sigma * (phi(u) + u * Phi(u))

and the Riemannian gradient in closed form (checked against torch.autograd in the tests):

    grad_x EI = 2 beta sum_i w_i k_i Log_x(X_i),   w_i = -Phi(u) alpha_i - phi(u) (M k)_i / sigma
"""
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import sphere as _sph
from . import spd as _spd

VAR_FLOOR = 1e-9  # botorch analytic EI: sigma = posterior.variance.clamp_min(1e-9).sqrt()


@dataclass
class GPData:
    manifold: str            # 'sphere' | 'spd'
    x_train: np.ndarray      # sphere: (n, D) unit vectors; spd: (n, d, d) SPD matrices
    y: np.ndarray            # (n,)
    mean: float
    outputscale: float
    noise: float
    beta: float
    best_f: float
    alpha: np.ndarray = None  # (n,)   (K + noise I)^-1 (y - mean)
    minv: np.ndarray = None   # (n, n) (K + noise I)^-1
    kxx: float = 1.0          # k(x, x) of the base kernel (1 up to the reference's 1e-15 quirks)


def train_gram(manifold, x_train, beta):
    if manifold == 'sphere':
        return _sph.sphere_gaussian_kernel(x_train, x_train, beta).numpy()
    d = _spd.affine_invariant_distance(torch.as_tensor(x_train), torch.as_tensor(x_train), exact=True)
    return torch.exp(-d * d * beta).numpy()


def make_gp(manifold, x_train, y, beta, outputscale=1.0, noise=1e-2, mean=None, best_f=None):
    x_train = np.asarray(x_train, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    mean = float(np.mean(y)) if mean is None else float(mean)
    best_f = float(np.min(y)) if best_f is None else float(best_f)
    K = outputscale * train_gram(manifold, x_train, beta) + noise * np.eye(len(y))
    K = 0.5 * (K + K.T)
    minv = np.linalg.inv(K)
    minv = 0.5 * (minv + minv.T)
    alpha = np.linalg.solve(K, y - mean)
    return GPData(manifold, x_train, y, mean, float(outputscale), float(noise), float(beta), best_f,
                  alpha=alpha, minv=minv)


def _phi(u):
    return math.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def _Phi(u):
    return 0.5 * math.erfc(-u / math.sqrt(2.0))


def _dist_and_logs(gp, x, want_logs):
    """Distances d(x, X_i) and (optionally) Log_x(X_i) for all training points."""
    if gp.manifold == 'sphere':
        xt = gp.x_train
        c = np.clip(xt @ x, -1. + _sph.CLAMP_EPS, 1. - _sph.CLAMP_EPS)
        d = np.arccos(c)
        logs = None
        if want_logs:
            xb = np.broadcast_to(x, xt.shape)
            logs = _sph.log(xb, xt)
        return d, logs
    xt = gp.x_train
    c = np.linalg.cholesky(x)
    ci = np.linalg.inv(c)
    w = ci @ xt @ ci.T
    w = 0.5 * (w + np.swapaxes(w, -1, -2))
    lam, v = np.linalg.eigh(w)
    ll = np.log(lam)
    d = np.sqrt(np.sum(ll * ll, axis=-1) + _spd.DIST_EPS)
    logs = None
    if want_logs:
        lw = (v * ll[:, None, :]) @ np.swapaxes(v, -1, -2)
        logs = c @ lw @ c.T
    return d, logs


def ei_and_grad(gp, x, want_grad=True):
    """EI(x) and its Riemannian gradient at a single point x (sphere: (D,), spd: (d,d))."""
    d, logs = _dist_and_logs(gp, x, want_grad)
    k = gp.outputscale * np.exp(-gp.beta * d * d)
    mk = gp.minv @ k
    mu = gp.mean + float(k @ gp.alpha)
    var_raw = gp.outputscale * gp.kxx - float(k @ mk)
    clamped = var_raw < VAR_FLOOR
    sigma = math.sqrt(max(var_raw, VAR_FLOOR))
    u = (gp.best_f - mu) / sigma
    pdf, cdf = _phi(u), _Phi(u)
    ei = sigma * (pdf + u * cdf)
    if not want_grad:
        return ei, None
    w = -cdf * gp.alpha
    if not clamped:
        w = w - pdf * mk / sigma
    coef = 2.0 * gp.beta * w * k
    grad = np.tensordot(coef, logs, axes=(0, 0))
    return ei, grad


def ei_batch(gp, xs):
    return np.array([ei_and_grad(gp, x, want_grad=False)[0] for x in xs])


def ei_torch(gp, x):
    """Same EI written with differentiable torch ops, for the autograd cross-check of the closed-form
    gradient (mirrors what the reference does through ``pymanopt_addons/tools/autodiff/_pytorch.py:83-101``)."""
    xt = torch.as_tensor(gp.x_train)
    if gp.manifold == 'sphere':
        c = (xt @ x).clamp(-1. + _sph.CLAMP_EPS, 1. - _sph.CLAMP_EPS)
        d = torch.acos(c)
    else:
        ch = torch.linalg.cholesky(x)
        ci = torch.inverse(ch)
        w = ci @ xt @ ci.T
        lam = torch.linalg.eigvalsh(0.5 * (w + w.transpose(-1, -2)))
        d = torch.sqrt(torch.sum(torch.log(lam) ** 2, dim=-1) + _spd.DIST_EPS)
    k = gp.outputscale * torch.exp(-gp.beta * d * d)
    minv = torch.as_tensor(gp.minv)
    mu = gp.mean + k @ torch.as_tensor(gp.alpha)
    var = (gp.outputscale * gp.kxx - k @ (minv @ k)).clamp_min(VAR_FLOOR)
    sigma = var.sqrt()
    u = (gp.best_f - mu) / sigma
    normal = torch.distributions.Normal(torch.zeros((), dtype=x.dtype), torch.ones((), dtype=x.dtype))
    return sigma * (torch.exp(normal.log_prob(u)) + u * normal.cdf(u))


# ----------------------------------------------------------------------------------------------------------------
# Marginal likelihood (gpytorch ExactMarginalLogLikelihood; botorch fit_gpytorch_model).  PARITY UNPINNED (third-party);
# restated from the published definitions:  ll = log N(y | m 1, s exp(-beta Dm) + noise I);  gpytorch divides the sum
# of ll and the log-priors by n (ExactMarginalLogLikelihood.forward), priors are evaluated on the TRANSFORMED values.
# ----------------------------------------------------------------------------------------------------------------

def exact_log_likelihood(dmat, y, theta):
    """ll and its gradient w.r.t. theta = (beta, s, noise, m) by torch.autograd on torch.distributions."""
    dmat = torch.as_tensor(dmat, dtype=torch.float64)
    y = torch.as_tensor(y, dtype=torch.float64)
    th = torch.as_tensor(theta, dtype=torch.float64).clone().requires_grad_(True)
    n = y.shape[0]
    k = th[1] * torch.exp(-th[0] * dmat) + th[2] * torch.eye(n, dtype=torch.float64)
    k = 0.5 * (k + k.T)
    dist = torch.distributions.MultivariateNormal(th[3] * torch.ones(n, dtype=torch.float64), covariance_matrix=k)
    ll = dist.log_prob(y)
    grad, = torch.autograd.grad(ll, th)
    return float(ll.detach()), grad.numpy()


def gamma_log_prob(x, concentration, rate):
    """torch.distributions.Gamma(concentration, rate).log_prob(x) (gpytorch GammaPrior)."""
    return (concentration * math.log(rate) + (concentration - 1.0) * math.log(x) - rate * x
            - math.lgamma(concentration))


def softplus(r):
    return math.log1p(math.exp(-abs(r))) + max(r, 0.0)


def mll_objective(dmat, y, raw, beta_min, noise_min=1e-8, outputscale_prior=None, noise_prior=None, beta_prior=None):
    """-(ll + sum of log-priors) / n at raw = (raw_beta, raw_outputscale, raw_noise, mean): what fit_gpytorch_model
    hands to scipy (constraints GreaterThan(beta_min) / Positive / GreaterThan(noise_min) as softplus transforms)."""
    beta = beta_min + softplus(raw[0])
    s = softplus(raw[1])
    noise = noise_min + softplus(raw[2])
    ll, _ = exact_log_likelihood(dmat, y, (beta, s, noise, raw[3]))
    for val, prior in ((s, outputscale_prior), (noise, noise_prior), (beta, beta_prior)):
        if prior is not None:
            ll += gamma_log_prob(val, *prior)
    return -ll / len(y)
