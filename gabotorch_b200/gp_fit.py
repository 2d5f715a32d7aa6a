"""GP hyper-parameter fit on the device.

The reference fits its GP at the top of every BO iteration with ``botorch.fit_gpytorch_model(mll=mll_fct)``
(examples/bo_sphere/benchmark_examples/gabo_sphere.py:162; model set-up :131-147): scipy L-BFGS-B over the raw
(unconstrained) parameters of ``SingleTaskGP`` = constant mean + ``ScaleKernel(geodesic kernel)`` +
``GaussianLikelihood``, objective ``-ExactMarginalLogLikelihood`` = ``-(log N(y | m, K) + sum of log-priors) / n``.
Every objective evaluation there rebuilds the Gram matrix, factorises it and back-propagates through both.

Here the geodesic distance matrix is computed ONCE per fit by the fused Gram kernels; each objective evaluation is then
one launch of ``gabo_gp_mll`` (Cholesky, solves, log-determinant, closed-form gradient, one CTA) and a 40-byte
read-back.  The optimiser itself is scipy's L-BFGS-B, as in botorch; constraints are gpytorch's softplus transforms
(``GreaterThan(lb)``: ``lb + softplus(raw)``; ``Positive``: ``softplus(raw)``), priors are evaluated on the
transformed values.  ``raw_samples > 0`` screens that many random hyper-parameter sets in ONE batched launch first and
starts from the best (the batch dimension of the kernel), which botorch does not do.

gpytorch / botorch are third-party and absent from the image: PARITY UNPINNED for the optimiser trajectory; the
objective and its gradient are tested against torch.distributions + autograd (oracle/gp.py).
"""
import math

import numpy as np
import torch

from . import _lib, ops
from .kernel_utils import (SphereGaussianKernel, SpdAffineInvariantGaussianKernel, SpdAffineInvariantLaplaceKernel)
from .manifold_optimization import ManifoldGP

NOISE_MIN = 1e-8     # gpytorch.constraints.GreaterThan(1e-8), gabo_sphere.py:140-142


def _softplus(r):
    return math.log1p(math.exp(-abs(r))) + max(r, 0.0)


def _sigmoid(r):
    return 1.0 / (1.0 + math.exp(-r)) if r >= 0 else math.exp(r) / (1.0 + math.exp(r))


def _inv_softplus(x):
    x = max(float(x), 1e-300)
    return x + math.log(-math.expm1(-x)) if x < 30 else x


def _prior_of(module, name):
    entry = getattr(module, '_priors', {}).get(name)
    if entry is None:
        return None
    prior = entry[0]
    return float(prior.concentration), float(prior.rate)


def kernel_distance_matrix(base, x):
    """(Dm, power) with the kernel = exp(-beta Dm): squared geodesic distances for the Gaussian kernels, distances for
    the Laplace kernel.  One launch of the fused Gram kernel in distance mode; fp64 eigen-solve for SPD inputs."""
    x = ops.to_dev64(x)
    if isinstance(base, SphereGaussianKernel):
        d = ops.sphere_gram(x, x, kind=_lib.KIND_DIST)
        return d * d, 2
    if isinstance(base, SpdAffineInvariantLaplaceKernel):
        return ops.spd_ai_gram(x, x, kind=_lib.KIND_DIST, compute=_lib.GABO_F64), 1
    if isinstance(base, SpdAffineInvariantGaussianKernel):
        d = ops.spd_ai_gram(x, x, kind=_lib.KIND_DIST, compute=_lib.GABO_F64)
        return d * d, 2
    raise NotImplementedError('GP fit supports SphereGaussianKernel and SpdAffineInvariant{Gaussian,Laplace}Kernel, '
                              'got %s' % type(base).__name__)


class MarginalLogLikelihood:
    """-(ll + log-priors) / n and its gradient in the raw parameters (raw_beta, raw_outputscale, raw_noise, mean)."""

    def __init__(self, dmat, y, beta_min, noise_min=NOISE_MIN, outputscale_prior=None, noise_prior=None,
                 beta_prior=None):
        self.dmat = ops.to_dev64(dmat)
        self.y = ops.to_dev64(y).reshape(-1)
        self.n = int(self.y.shape[0])
        self.beta_min, self.noise_min = float(beta_min), float(noise_min)
        self.priors = (beta_prior, outputscale_prior, noise_prior)
        self.evaluations = 0

    def transform(self, raw):
        return (self.beta_min + _softplus(raw[0]), _softplus(raw[1]), self.noise_min + _softplus(raw[2]),
                float(raw[3]))

    def inverse_transform(self, theta):
        return np.array([_inv_softplus(theta[0] - self.beta_min), _inv_softplus(theta[1]),
                         _inv_softplus(theta[2] - self.noise_min), float(theta[3])])

    def _prior_terms(self, theta):
        lp, dlp = 0.0, [0.0, 0.0, 0.0]
        for i, prior in enumerate(self.priors):
            if prior is None:
                continue
            c, r = prior
            v = theta[i]
            lp += c * math.log(r) + (c - 1.0) * math.log(v) - r * v - math.lgamma(c)
            dlp[i] = (c - 1.0) / v - r
        return lp, dlp

    def batch_values(self, raws):
        """Objective for many raw parameter sets in one launch (no gradient): numpy (B,), inf where K is not PD."""
        thetas = np.array([self.transform(r) for r in raws])
        ll, _, _, _, flags = ops.gp_mll(self.dmat, self.y, torch.from_numpy(thetas), want_grad=False)
        ll = ll.cpu().numpy()
        bad = flags.cpu().numpy() != 0
        out = np.array([-(l + self._prior_terms(t)[0]) / self.n for l, t in zip(ll, thetas)])
        out[bad | ~np.isfinite(out)] = np.inf
        self.evaluations += len(thetas)
        return out

    def __call__(self, raw):
        """(objective, gradient) at one raw parameter vector: what scipy's L-BFGS-B consumes (jac=True)."""
        theta = self.transform(raw)
        ll, grad, _, _, flags = ops.gp_mll(self.dmat, self.y, torch.tensor([theta], dtype=torch.float64))
        packed = torch.cat([ll, grad.reshape(-1), flags.double()]).cpu().numpy()   # one 48-byte read-back
        self.evaluations += 1
        if packed[5] != 0 or not np.isfinite(packed[0]):
            return 1e10, np.zeros(4)
        lp, dlp = self._prior_terms(theta)
        g = packed[1:5].copy()
        for i in range(3):
            g[i] = (g[i] + dlp[i]) * _sigmoid(raw[i])
        return -(packed[0] + lp) / self.n, -g / self.n


def _model_parts(model):
    cov = model.covar_module
    base = getattr(cov, 'base_kernel', None)
    scaled = base is not None
    if not scaled:
        base = cov
    return cov, base, scaled


def fit_gpytorch_model(mll, options=None, raw_samples=0, generator=None, optimizer='device', num_restarts=1, **kwargs):
    """Drop-in for ``botorch.fit_gpytorch_model(mll=...)`` on a ``ManifoldGP`` (or an object with ``.model``): fits
    (beta, outputscale, noise, constant mean) in place and returns its argument.

    ``optimizer='device'`` (default): the whole quasi-Newton fit is ONE launch of ``gabo_gp_fit`` (one CTA per start,
    ``num_restarts`` starts side by side, best objective kept) and one read-back.  ``optimizer='scipy'``: scipy's
    L-BFGS-B on the host as in botorch, one ``gabo_gp_mll`` launch + read-back per evaluation (``options`` are passed to
    it; botorch's ``fit_gpytorch_scipy`` default: maxiter 15000).  A CALLABLE ``optimizer`` is called as
    ``optimizer(mll, **kwargs)`` like botorch does -- ``fit_gpytorch_model(mll, optimizer=fit_gpytorch_manifold,
    solver=..., nb_init_candidates=...)`` is the call of hd_gabo_spd.py:205-206."""
    if callable(optimizer):
        if options is not None:
            kwargs['options'] = options
        optimizer(mll, **kwargs)
        return mll
    model = getattr(mll, 'model', mll)
    if not isinstance(model, ManifoldGP):
        raise NotImplementedError('fit_gpytorch_model expects a gabotorch_b200.ManifoldGP (or an mll holding one)')
    cov, base, scaled = _model_parts(model)
    x = model.train_inputs[0]
    dmat, _ = kernel_distance_matrix(base, x.reshape(-1, x.shape[-1]))
    objective = MarginalLogLikelihood(
        dmat, model.train_targets, float(base.beta_min), getattr(model, 'noise_min', NOISE_MIN),
        outputscale_prior=_prior_of(cov, 'outputscale_prior') if scaled else None,
        noise_prior=getattr(model, 'noise_prior', None), beta_prior=_prior_of(base, 'beta_prior'))
    theta0 = (float(base.beta.detach().reshape(-1)[0]), float(cov.outputscale.detach()) if scaled else 1.0,
              max(model.noise, objective.noise_min * (1 + 1e-6) + 1e-300), model.mean)
    raw0 = objective.inverse_transform(theta0)
    rng = np.random.default_rng(None if generator is None else generator)
    starts = raw0[None]
    if raw_samples and raw_samples > 0:
        # batched screening of random raw parameter sets around the start (one launch), keep the best
        cand = raw0[None] + rng.standard_normal((int(raw_samples), 4)) * np.array([2.0, 2.0, 3.0, 0.0])
        cand[:, 3] = raw0[3]
        cand = np.vstack([raw0[None], cand])
        order = np.argsort(objective.batch_values(cand), kind='stable')
        starts = cand[order[:max(1, int(num_restarts))]]
    elif num_restarts > 1:
        extra = raw0[None] + rng.standard_normal((int(num_restarts) - 1, 4)) * np.array([2.0, 2.0, 3.0, 0.0])
        starts = np.vstack([raw0[None], extra])
    opts = {'maxiter': 15000}
    opts.update(options or {})
    if optimizer == 'device':
        pri = []
        for p in objective.priors:
            pri += [p[0], p[1]] if p is not None else [0.0, 0.0]
        if not scaled:
            starts = starts.copy()
            starts[:, 1] = raw0[1]
        raws, fs, info = ops.gp_fit(objective.dmat, objective.y, torch.from_numpy(np.ascontiguousarray(starts)),
                                    objective.beta_min, objective.noise_min, pri, [0, 0 if scaled else 1, 0, 0],
                                    maxiter=int(opts['maxiter']), pgtol=float(opts.get('gtol', 1e-5)),
                                    ftol=float(opts.get('ftol', 2.220446049250313e-09)))
        fs = np.where(np.isfinite(fs), fs, np.inf)
        best = int(np.argmin(fs))
        if not np.isfinite(fs[best]):
            raise _lib.GaboError('fit_gpytorch_model: the covariance is not positive definite at any start')
        xbest, fbest = raws[best], float(fs[best])
        nit, nev, success = int(info[best, 1]), int(info[:, 2].sum()), bool(info[best, 0] in (0, 1))
    elif optimizer == 'scipy':
        from scipy.optimize import minimize
        lower = np.array([-np.inf, -np.inf if scaled else raw0[1], -np.inf, -np.inf])
        upper = np.array([np.inf, np.inf if scaled else raw0[1], np.inf, np.inf])
        res = minimize(objective, starts[0], jac=True, method='L-BFGS-B', bounds=list(zip(lower, upper)), options=opts)
        xbest, fbest, nit, nev, success = res.x, float(res.fun), int(res.nit), objective.evaluations, bool(res.success)
    else:
        raise ValueError("optimizer must be 'device' or 'scipy'")
    beta, s, noise, mean = objective.transform(xbest)
    base.beta = beta
    if scaled:
        cov.outputscale = s
    model.noise, model.mean = noise, mean
    model.fit_result = {'objective': fbest, 'iterations': nit, 'evaluations': nev, 'converged': success,
                        'theta': (beta, s, noise, mean), 'optimizer': optimizer}
    return mll


class ExactMarginalLogLikelihood:
    """Holder with gpytorch's constructor shape ``(likelihood, model)`` so that the reference's two lines
    (``mll_fct = ExactMarginalLogLikelihood(model.likelihood, model)``; ``fit_gpytorch_model(mll=mll_fct)``,
    gabo_sphere.py:147,162) keep working."""

    def __init__(self, likelihood, model):
        self.likelihood = likelihood
        self.model = model
