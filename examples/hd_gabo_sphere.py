"""HD-GaBO on the sphere with gabotorch_b200: the loop of the reference's
``examples/hd_bo_sphere/benchmark_examples/hd_gabo_sphere.py`` (:95-225) written against the drop-in modules.

Every BO iteration: (1) ``fit_gpytorch_manifold`` fits the GP on S^D with the nested-sphere kernel, including the axes of
its projection chain (one ``Sphere`` parameter manifold per level); (2) the data are projected to the latent S^d
(``projection_from_sphere_to_subsphere``) and a latent GP with the plain sphere kernel takes over the fitted
hyper-parameters; (3) ``optimize_reconstruction_parameters_nested_sphere`` fits the distances-to-axis of the map back to
the ambient sphere; (4) EI is maximised on the LATENT sphere with multi-start trust regions (one launch); (5) the candidate
is lifted with ``projection_from_subsphere_to_sphere`` and evaluated.  Needs a B200 (no CPU fallback).

    python examples/hd_gabo_sphere.py [--dim 5] [--latent-dim 3] [--iters 25] [--seed 1234]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gabotorch_b200 as g  # noqa: E402
from gabotorch_b200 import nested_mappings as nm  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gabo_sphere import ackley_sphere  # noqa: E402

BETA_MIN = {3: 6.5, 4: 2.0, 5: 1.2}


def run(dim=5, latent_dim=3, n_iters=25, num_restarts=5, raw_samples=100, nb_data_init=5, seed=1234, verbose=True):
    """Returns (x_data (n, dim), y_data (n,), best_f per iteration)."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    manifold, latent_manifold = g.Sphere(dim), g.Sphere(latent_dim)
    # nested test function (nested_test_functions_sphere.py:14-41): Ackley of the projection onto a fixed subsphere
    test_axes = [torch.from_numpy(g.Sphere(k).rand()).reshape(1, -1) for k in range(dim, latent_dim, -1)]
    test_dists = [torch.full((1, 1), np.pi / 2, dtype=torch.float64) for _ in test_axes]

    def objective(x):
        sub = nm.projection_from_sphere_to_subsphere(torch.as_tensor(x, dtype=torch.float64).reshape(1, -1), test_axes,
                                                     test_dists)[-1]
        return ackley_sphere(latent_manifold, sub.cpu().numpy())

    x_data = torch.tensor(np.array([manifold.rand() for _ in range(nb_data_init)]))
    y_data = torch.tensor([objective(x) for x in x_data], dtype=torch.float64)
    beta_min = BETA_MIN.get(latent_dim, 1.0)
    k_fct = g.ScaleKernel(g.NestedSphereGaussianKernel(dim=dim, latent_dim=latent_dim, beta_min=beta_min),
                          outputscale_prior=g.GammaPrior(2.0, 0.15))
    latent_k_fct = g.ScaleKernel(g.SphereGaussianKernel(beta_min=beta_min), outputscale_prior=g.GammaPrior(2.0, 0.15))
    noise_prior = g.GammaPrior(1.1, 0.05)
    noise = float((noise_prior.concentration - 1) / noise_prior.rate)
    mean = 0.0
    reconstruction_solver, solver = g.TrustRegions(), g.TrustRegions()            # hd_gabo_sphere.py:161-162
    bounds = torch.stack([-torch.ones(dim, dtype=torch.float64), torch.ones(dim, dtype=torch.float64)])
    best_f = [float(y_data.min())]
    for it in range(n_iters):
        model = g.ManifoldGP(x_data, y_data, k_fct, noise=noise, mean=mean, noise_prior=noise_prior)
        g.fit_gpytorch_model(g.ExactMarginalLogLikelihood(model.likelihood, model), optimizer=g.fit_gpytorch_manifold)
        noise, mean = model.noise, model.mean
        axes = [a.detach().clone().double().cpu() for a in k_fct.base_kernel.axes]
        dists = [r.detach().clone().double().cpu() for r in k_fct.base_kernel.distances_to_axis]
        x_sub = nm.projection_from_sphere_to_subsphere(x_data, axes, dists)[-1].to('cpu', torch.float64)
        latent_k_fct.base_kernel.beta = float(k_fct.base_kernel.beta.detach())        # :181-183
        latent_k_fct.outputscale = k_fct.outputscale.detach()
        latent_model = g.ManifoldGP(x_sub, y_data, latent_k_fct, noise=noise, mean=mean)
        dists = g.optimize_reconstruction_parameters_nested_sphere(x_data, x_sub, axes, reconstruction_solver)
        acq = g.ExpectedImprovement(model=latent_model, best_f=best_f[-1], maximize=False)
        new_sub = g.joint_optimize_manifold(acq, latent_manifold, solver, q=1, num_restarts=num_restarts,
                                            raw_samples=raw_samples, bounds=bounds, approx_hessian=True,
                                            options={'seed': seed + it})
        new_sub = new_sub.reshape(1, latent_dim).to('cpu', torch.float64)
        new_x = nm.projection_from_subsphere_to_sphere(new_sub, axes, dists)[-1].to('cpu', torch.float64).reshape(1, dim)
        new_y = objective(new_x)
        x_data = torch.cat((x_data, new_x))
        y_data = torch.cat((y_data, torch.tensor([new_y], dtype=torch.float64)))
        best_f.append(min(best_f[-1], new_y))
        if verbose:
            print('iteration %2d  f(x) = %.5f  best = %.5f  |x| = %.12f  distances to axis %s'
                  % (it + 1, new_y, best_f[-1], float(new_x.norm()), [round(float(r), 3) for r in dists]))
    return x_data, y_data, best_f


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--dim', type=int, default=5)
    ap.add_argument('--latent-dim', type=int, default=3)
    ap.add_argument('--iters', type=int, default=25)
    ap.add_argument('--restarts', type=int, default=5)
    ap.add_argument('--raw-samples', type=int, default=100)
    ap.add_argument('--seed', type=int, default=1234)
    a = ap.parse_args()
    run(a.dim, a.latent_dim, a.iters, a.restarts, a.raw_samples, seed=a.seed)
