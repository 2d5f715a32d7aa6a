"""GaBO on the sphere with gabotorch_b200: the loop of the reference's
``examples/bo_sphere/benchmark_examples/gabo_sphere.py`` (:86-197) written against the drop-in modules -- same model
(constant mean + ScaleKernel(SphereGaussianKernel(beta_min)) with Gamma(2, 0.15) on the outputscale, Gamma(1.1, 0.05) on
the noise starting at its mode), same steps per iteration (fit the GP, Expected Improvement, multi-start trust regions
with 5 restarts out of 100 raw samples), same objective (Ackley on the tangent space at (1, 0, ...),
``BoManifolds/BO_test_functions/test_functions_sphere.py:34-65``).  Needs a B200 (no CPU fallback).

    python examples/gabo_sphere.py [--dim 3] [--iters 25] [--restarts 5] [--raw-samples 100] [--seed 1234]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gabotorch_b200 as g  # noqa: E402

BETA_MIN = {3: 6.5, 4: 2.0, 5: 1.2, 11: 0.6, 21: 0.35, 51: 0.21, 101: 0.21}      # gabo_sphere.py:115-128


def ackley_sphere(manifold, x):
    """Scalar objective evaluated once per BO iteration on the host (test_functions_sphere.py:34-65)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    dim = x.shape[-1]
    base = np.zeros((1, dim))
    base[0, 0] = 1.0
    proj = np.asarray(manifold.log(base, x)).reshape(-1)[1:]
    r = dim - 1
    return float(-20.0 * np.exp(-0.2 * np.sqrt(np.sum(proj ** 2) / r)) - np.exp(np.sum(np.cos(2 * np.pi * proj) / r))
                 + 20.0 + np.exp(1.0))


def run(dim=3, n_iters=25, num_restarts=5, raw_samples=100, nb_data_init=5, seed=1234, solver=None, verbose=True):
    """Returns (x_data (n, dim), y_data (n,), best_f per iteration)."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    manifold = g.Sphere(dim)
    x_data = torch.tensor(np.array([manifold.rand() for _ in range(nb_data_init)]))
    y_data = torch.tensor([ackley_sphere(manifold, x) for x in x_data.numpy()], dtype=torch.float64)
    beta_min = BETA_MIN.get(dim, 0.21)
    covar = g.ScaleKernel(g.SphereGaussianKernel(beta_min=beta_min), outputscale_prior=g.GammaPrior(2.0, 0.15))
    noise_prior = g.GammaPrior(1.1, 0.05)
    noise = float((noise_prior.concentration - 1) / noise_prior.rate)            # prior mode, gabo_sphere.py:137-142
    mean = 0.0
    solver = solver or g.TrustRegions()
    bounds = torch.stack([-torch.ones(dim, dtype=torch.float64), torch.ones(dim, dtype=torch.float64)])
    best_f = [float(y_data.min())]
    for it in range(n_iters):
        model = g.ManifoldGP(x_data, y_data, covar, noise=noise, mean=mean, noise_prior=noise_prior)
        g.fit_gpytorch_model(g.ExactMarginalLogLikelihood(model.likelihood, model))
        noise, mean = model.noise, model.mean                                     # warm start of the next fit
        acq = g.ExpectedImprovement(model=model, best_f=best_f[-1], maximize=False)
        new_x = g.joint_optimize_manifold(acq, manifold, solver, q=1, num_restarts=num_restarts,
                                          raw_samples=raw_samples, bounds=bounds, approx_hessian=True,
                                          options={'seed': seed + it})
        new_x = new_x.reshape(1, dim).to('cpu', torch.float64)
        new_y = ackley_sphere(manifold, new_x.numpy())
        x_data = torch.cat((x_data, new_x))
        y_data = torch.cat((y_data, torch.tensor([new_y], dtype=torch.float64)))
        best_f.append(min(best_f[-1], new_y))
        if verbose:
            print('iteration %2d  f(x) = %.5f  best = %.5f  beta = %.3f  noise = %.2e'
                  % (it + 1, new_y, best_f[-1], float(covar.base_kernel.beta.detach()), noise))
    return x_data, y_data, best_f


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--dim', type=int, default=3)
    ap.add_argument('--iters', type=int, default=25)
    ap.add_argument('--restarts', type=int, default=5)
    ap.add_argument('--raw-samples', type=int, default=100)
    ap.add_argument('--seed', type=int, default=1234)
    a = ap.parse_args()
    run(a.dim, a.iters, a.restarts, a.raw_samples, seed=a.seed)
