"""HD-GaBO on the SPD manifold with gabotorch_b200: the loop of the reference's
``examples/hd_bo_spd/benchmark_examples/hd_gabo_spd.py`` (:92-300) written against the drop-in modules.

Every BO iteration: (1) ``fit_gpytorch_manifold`` fits the GP on S^D_++ with the nested log-Euclidean kernel, including its
Grassmann projection matrix; (2) the data are projected to the latent S^d_++ and a latent GP with the plain log-Euclidean
kernel takes over the fitted hyper-parameters; (3) ``optimize_reconstruction_parameters_nested_spd`` fits the map back to
the ambient manifold (ALM around CG, log-Euclidean cost); (4) EI is maximised on the LATENT manifold with
``StrictConstrainedTrustRegions`` under eigenvalue constraints expressed on the AMBIENT matrix
(``max/min_eigenvalue_nested_spd_constraint``), raw samples drawn by projecting ambient samples; (5) the candidate is
reconstructed and evaluated.  Needs a B200 (no CPU fallback).

    python examples/hd_gabo_spd.py [--dim 5] [--latent-dim 2] [--iters 30] [--seed 1234]
"""
import argparse
import functools
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gabotorch_b200 as g  # noqa: E402
from gabotorch_b200 import riemannian_utils as ru  # noqa: E402


def nested_objective(x_mandel, w_test):
    """Scalar test function on S^D_++ that only depends on the projection W_test^T X W_test (the structure of
    ``projected_function_spd``, nested_test_functions_spd.py:17-46): Ackley of the latent matrix' log at 2 I."""
    x = ru.vector_to_symmetric_matrix_mandel(np.asarray(x_mandel, dtype=np.float64))
    y = w_test.T @ x @ w_test
    lam, q = np.linalg.eigh(0.5 * (y + y.T) / 2.0)
    v = ru.symmetric_matrix_to_vector_mandel((q * np.log(lam)) @ q.T)
    v[y.shape[0]:] /= 2.0 ** 0.5
    dv = v.shape[0]
    return float(-20.0 * np.exp(-0.2 * np.sqrt(np.sum(v ** 2) / dv)) - np.exp(np.sum(np.cos(2 * np.pi * v) / dv))
                 + 20.0 + np.exp(1.0))


def run(dim=5, latent_dim=2, n_iters=30, num_restarts=5, raw_samples=100, nb_data_init=5, seed=1234, verbose=True):
    """Returns (x_data (n, dim(dim+1)/2) Mandel vectors, y_data (n,), best_f per iteration)."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    spd_manifold, latent_manifold = g.PositiveDefinite(dim), g.PositiveDefinite(latent_dim)
    min_eig, max_eig = 1e-4, 5.0                                                   # hd_gabo_spd.py:137-143
    for m in (spd_manifold, latent_manifold):
        m.min_eig, m.max_eig = min_eig, max_eig
    spd_manifold.rand = types.MethodType(ru.spd_sample, spd_manifold)              # :120
    w_test, _ = np.linalg.qr(np.random.randn(dim, latent_dim))                     # Grassmann(dim, latent_dim).rand()
    x_init = np.array([spd_manifold.rand() for _ in range(nb_data_init)])
    x_data = torch.from_numpy(x_init)
    x_data_vec = torch.from_numpy(np.array([ru.symmetric_matrix_to_vector_mandel(m) for m in x_init]))
    y_data = torch.tensor([nested_objective(v.numpy(), w_test) for v in x_data_vec], dtype=torch.float64)

    k_fct = g.ScaleKernel(g.NestedSpdLogEuclideanGaussianKernel(dim=dim, latent_dim=latent_dim),
                          outputscale_prior=g.GammaPrior(2.0, 0.15))
    latent_k_fct = g.ScaleKernel(g.SpdLogEuclideanGaussianKernel(), outputscale_prior=g.GammaPrior(2.0, 0.15))
    noise_prior = g.GammaPrior(1.1, 0.05)
    noise = float((noise_prior.concentration - 1) / noise_prior.rate)
    mean = 0.0
    projection_solver = g.ConjugateGradient(maxiter=50)                            # :191-193
    reconstruction_solver = g.ConjugateGradient(maxiter=100)
    acquisition_solver = g.StrictConstrainedTrustRegions(mingradnorm=2e-4, maxiter=100, minstepsize=1e-4)
    best_f = [float(y_data.min())]
    for it in range(n_iters):
        model = g.ManifoldGP(x_data_vec, y_data, k_fct, noise=noise, mean=mean, noise_prior=noise_prior)
        g.fit_gpytorch_model(g.ExactMarginalLogLikelihood(model.likelihood, model), optimizer=g.fit_gpytorch_manifold,
                             solver=projection_solver, nb_init_candidates=20)
        noise, mean = model.noise, model.mean
        w = k_fct.base_kernel.projection_matrix.detach().clone().double().cpu()
        x_proj = g.projection_from_spd_to_nested_spd(x_data, w).double()
        x_proj_vec = ru.symmetric_matrix_to_vector_mandel_torch(x_proj).cpu()
        latent_k_fct.base_kernel.lengthscale = k_fct.base_kernel.lengthscale.detach()      # :215-217
        latent_k_fct.outputscale = k_fct.outputscale.detach()
        latent_model = g.ManifoldGP(x_proj_vec, y_data, latent_k_fct, noise=noise, mean=mean)
        v, c, k = g.optimize_reconstruction_parameters_nested_spd(
            x_data, x_proj, w, reconstruction_solver, cost_function=g.min_log_euclidean_distance_reconstruction_cost)
        latent_manifold.rand = types.MethodType(
            functools.partial(g.random_nested_spd_with_spd_eigenvalue_constraints, random_spd_fct=spd_manifold.rand,
                              projection_matrix=w), latent_manifold)
        pars = dict(projection_matrix=w, projection_complement_matrix=v, bottom_spd_matrix=c, contraction_matrix=k)
        constraints = [functools.partial(g.max_eigenvalue_nested_spd_constraint, maximum_eigenvalue=max_eig, **pars),
                       functools.partial(g.min_eigenvalue_nested_spd_constraint, minimum_eigenvalue=min_eig, **pars)]
        acq = g.ExpectedImprovement(model=latent_model, best_f=best_f[-1], maximize=False)
        new_proj = g.joint_optimize_manifold(acq, latent_manifold, acquisition_solver, q=1, num_restarts=num_restarts,
                                             raw_samples=raw_samples, bounds=None,
                                             pre_processing_manifold=ru.vector_to_symmetric_matrix_mandel_torch,
                                             post_processing_manifold=ru.symmetric_matrix_to_vector_mandel_torch,
                                             approx_hessian=True, inequality_constraints=constraints,
                                             options={'seed': seed + it})
        new_proj = ru.vector_to_symmetric_matrix_mandel_torch(new_proj.reshape(1, -1).to('cpu', torch.float64))
        new_x = g.projection_from_nested_spd_to_spd(new_proj, w, v, c, k).to('cpu', torch.float64)
        new_x_vec = ru.symmetric_matrix_to_vector_mandel_torch(new_x).cpu()
        new_y = nested_objective(new_x_vec[0].numpy(), w_test)
        x_data = torch.cat((x_data, new_x))
        x_data_vec = torch.cat((x_data_vec, new_x_vec))
        y_data = torch.cat((y_data, torch.tensor([new_y], dtype=torch.float64)))
        best_f.append(min(best_f[-1], new_y))
        if verbose:
            print('iteration %2d  f(x) = %.5f  best = %.5f  reconstruction cost %.3e  ambient eigenvalues [%.3g, %.3g]'
                  % (it + 1, new_y, best_f[-1], g.optimize_reconstruction_parameters_nested_spd.last_log['cost'],
                     float(torch.linalg.eigvalsh(new_x[0]).min()), float(torch.linalg.eigvalsh(new_x[0]).max())))
    return x_data_vec, y_data, best_f


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--dim', type=int, default=5)
    ap.add_argument('--latent-dim', type=int, default=2)
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--restarts', type=int, default=5)
    ap.add_argument('--raw-samples', type=int, default=100)
    ap.add_argument('--seed', type=int, default=1234)
    a = ap.parse_args()
    run(a.dim, a.latent_dim, a.iters, a.restarts, a.raw_samples, seed=a.seed)
