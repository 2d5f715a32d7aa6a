"""GaBO on the SPD manifold with gabotorch_b200: the loop of the reference's
``examples/bo_spd/benchmark_examples/gabo_spd.py`` (:92-210) written against the drop-in modules -- GP inputs in Mandel
notation, solver iterates as matrices (pre / post processing), ``ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100)``
with the finite-difference Hessian and one max-eigenvalue inequality constraint, eigenvalue domain [0.001, 5] for the
raw samples (``spd_sample`` law), objective = Ackley on the tangent space at 2 I
(``BoManifolds/BO_test_functions/test_functions_spd.py:34-69``).  Needs a B200 (no CPU fallback).

    python examples/gabo_spd.py [--dim 2] [--iters 25] [--restarts 5] [--raw-samples 100] [--seed 1234]
"""
import argparse
import functools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gabotorch_b200 as g  # noqa: E402
from gabotorch_b200 import riemannian_utils as ru  # noqa: E402

BETA_MIN = {2: 0.6, 3: 0.5, 5: 0.25, 7: 0.22}                                     # gabo_spd.py:151-162 (d <= 8 here)


def ackley_spd(manifold, x_mandel):
    """Scalar objective evaluated once per BO iteration on the host (test_functions_spd.py:34-69)."""
    x = ru.vector_to_symmetric_matrix_mandel_torch(torch.as_tensor(x_mandel, dtype=torch.float64).reshape(1, -1))[0]
    d = x.shape[-1]
    proj = torch.as_tensor(manifold.log(2.0 * np.eye(d), x.cpu().numpy()), dtype=torch.float64)
    v = ru.symmetric_matrix_to_vector_mandel_torch(proj.reshape(1, d, d))[0].cpu().numpy().copy()
    v[d:] /= 2.0 ** 0.5
    dv = v.shape[0]
    return float(-20.0 * np.exp(-0.2 * np.sqrt(np.sum(v ** 2) / dv)) - np.exp(np.sum(np.cos(2 * np.pi * v) / dv))
                 + 20.0 + np.exp(1.0))


def run(dim=2, n_iters=25, num_restarts=5, raw_samples=100, nb_data_init=5, seed=1234, verbose=True):
    """Returns (x_data (n, dim(dim+1)/2) Mandel vectors, y_data (n,), best_f per iteration)."""
    torch.manual_seed(seed)
    manifold = g.PositiveDefinite(dim)
    min_eig, max_eig = 0.001, 5.0                                                 # gabo_spd.py:121-124
    manifold.min_eig, manifold.max_eig = min_eig, max_eig
    constraints = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=max_eig)]
    gen = torch.Generator(device=g.ops.device())
    gen.manual_seed(seed)
    x_data = ru.symmetric_matrix_to_vector_mandel_torch(manifold.rand_batch(nb_data_init, generator=gen)).cpu()
    y_data = torch.tensor([ackley_spd(manifold, x) for x in x_data], dtype=torch.float64)
    covar = g.ScaleKernel(g.SpdAffineInvariantGaussianKernel(beta_min=BETA_MIN.get(dim, 0.2)),
                          outputscale_prior=g.GammaPrior(2.0, 0.15))
    noise_prior = g.GammaPrior(1.1, 0.05)
    noise = float((noise_prior.concentration - 1) / noise_prior.rate)
    mean = 0.0
    solver = g.ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100)
    best_f = [float(y_data.min())]
    for it in range(n_iters):
        model = g.ManifoldGP(x_data, y_data, covar, noise=noise, mean=mean, noise_prior=noise_prior)
        g.fit_gpytorch_model(g.ExactMarginalLogLikelihood(model.likelihood, model))
        noise, mean = model.noise, model.mean
        acq = g.ExpectedImprovement(model=model, best_f=best_f[-1], maximize=False)
        new_x = g.joint_optimize_manifold(acq, manifold, solver, q=1, num_restarts=num_restarts,
                                          raw_samples=raw_samples, bounds=None,
                                          pre_processing_manifold=ru.vector_to_symmetric_matrix_mandel_torch,
                                          post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch,
                                          approx_hessian=True, inequality_constraints=constraints,
                                          options={'seed': seed + it})
        new_x = new_x.reshape(1, -1).to('cpu', torch.float64)
        new_y = ackley_spd(manifold, new_x[0])
        x_data = torch.cat((x_data, new_x))
        y_data = torch.cat((y_data, torch.tensor([new_y], dtype=torch.float64)))
        best_f.append(min(best_f[-1], new_y))
        if verbose:
            print('iteration %2d  f(x) = %.5f  best = %.5f  beta = %.3f  noise = %.2e'
                  % (it + 1, new_y, best_f[-1], float(covar.base_kernel.beta.detach()), noise))
    return x_data, y_data, best_f


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--dim', type=int, default=2)
    ap.add_argument('--iters', type=int, default=25)
    ap.add_argument('--restarts', type=int, default=5)
    ap.add_argument('--raw-samples', type=int, default=100)
    ap.add_argument('--seed', type=int, default=1234)
    a = ap.parse_args()
    run(a.dim, a.iters, a.restarts, a.raw_samples, seed=a.seed)
