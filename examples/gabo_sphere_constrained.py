"""Constrained GaBO on S^2 with gabotorch_b200: the three loops of the reference's
``examples/bo_sphere/constrained_benchmark_examples/`` written against the drop-in modules.

  --constraint bound       gabo_sphere_bound_constraints.py: five inequality callables x >= 0, |y| <= 0.6, |z| <= 0.6,
                           ``ConstrainedTrustRegions(maxiter=200)``
  --constraint inequality  gabo_sphere_inequality_constraints.py: geodesic ball of radius pi/4 around (1, 0, 0)
  --constraint equality    gabo_sphere_equality_constraints.py: the great circle y = 0
  --solver ctr | alm       ``ConstrainedTrustRegions(maxiter=200)`` or
                           ``AugmentedLagrangeMethod(maxiter=200, inner_solver=TrustRegions(maxiter=200), gammas_fact=0.05)``

The constraints are plain torch callables of ONE point, as in the reference (differentiated with torch.autograd per restart
on the device); the feasible raw samples come from a sampler re-bound on the manifold object (``sphere_manifold.rand = ...``).
Needs a B200 (no CPU fallback).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import gabotorch_b200 as g  # noqa: E402
from gabo_sphere import ackley_sphere  # noqa: E402


def make_problem(kind, dim=3):
    """(inequality constraints, equality constraints, feasible sampler, feasibility check) of the three examples."""
    if kind == 'bound':                                   # gabo_sphere_bound_constraints.py:93-133
        xl, yl, yu, zl, zu = 0.0, -0.6, 0.6, -0.6, 0.6
        ineq = [lambda x: x[0] - xl, lambda x: x[1] - yl, lambda x: yu - x[1], lambda x: x[2] - zl, lambda x: zu - x[2]]

        def sample():
            while True:
                s = np.array([np.random.uniform(xl, 1.0), np.random.uniform(yl, yu), np.random.uniform(zl, zu)])
                s = s / np.linalg.norm(s)
                if s[0] > xl and yl < s[1] < yu and zl < s[2] < zu:
                    return s
        return ineq, None, sample, lambda x: min(float(c(torch.as_tensor(x))) for c in ineq)
    if kind == 'inequality':                              # gabo_sphere_inequality_constraints.py:109-150
        centre = np.zeros(dim)
        centre[0] = 1.0
        angle = np.pi / 4.0

        def domain_constraint(x):
            c = torch.as_tensor(centre, dtype=x.dtype, device=x.device)
            one = torch.ones(1, dtype=x.dtype, device=x.device)
            in_prod = torch.max(torch.min(torch.mm(x[None], c[:, None]), one), -one)
            return angle - torch.acos(in_prod)[0, 0]

        def sample():
            while True:                                   # uniform on the sphere, kept when inside the ball
                s = np.random.randn(dim)
                s /= np.linalg.norm(s)
                if np.arccos(np.clip(s @ centre, -1, 1)) < angle:
                    return s
        return [domain_constraint], None, sample, lambda x: float(domain_constraint(torch.as_tensor(x)))
    if kind == 'equality':                                # gabo_sphere_equality_constraints.py:101-118
        def y_great_circle(x):
            return x[1] - 0.0

        def sample():
            s = np.random.randn(dim)
            s[1] = 0.0
            return s / np.linalg.norm(s)
        return None, [y_great_circle], sample, lambda x: -abs(float(x[1]))
    raise ValueError('constraint must be bound, inequality or equality')


def run(constraint='inequality', solver_name='ctr', n_iters=25, num_restarts=5, raw_samples=100, nb_data_init=5, seed=1234,
        verbose=True):
    """Returns (x_data (n, 3), y_data (n,), best_f per iteration, worst constraint value over the proposed candidates)."""
    dim = 3
    np.random.seed(seed)
    torch.manual_seed(seed)
    manifold = g.Sphere(dim)
    ineq, eq, sample, feasibility = make_problem(constraint, dim)
    manifold.rand = sample                                # the reference's way of installing the feasible sampler
    x_data = torch.tensor(np.array([manifold.rand() for _ in range(nb_data_init)]))
    y_data = torch.tensor([ackley_sphere(g.Sphere(dim), x) for x in x_data.numpy()], dtype=torch.float64)
    covar = g.ScaleKernel(g.SphereGaussianKernel(beta_min=6.5), outputscale_prior=g.GammaPrior(2.0, 0.15))
    noise_prior = g.GammaPrior(1.1, 0.05)
    noise = float((noise_prior.concentration - 1) / noise_prior.rate)
    mean = 0.0
    if solver_name == 'ctr':
        solver = g.ConstrainedTrustRegions(maxiter=200)
    else:
        solver = g.AugmentedLagrangeMethod(maxiter=200, inner_solver=g.TrustRegions(maxiter=200), gammas_fact=0.05)
    bounds = torch.stack([-torch.ones(dim, dtype=torch.float64), torch.ones(dim, dtype=torch.float64)])
    best_f, worst = [float(y_data.min())], float('inf')
    for it in range(n_iters):
        model = g.ManifoldGP(x_data, y_data, covar, noise=noise, mean=mean, noise_prior=noise_prior)
        g.fit_gpytorch_model(g.ExactMarginalLogLikelihood(model.likelihood, model))
        noise, mean = model.noise, model.mean
        acq = g.ExpectedImprovement(model=model, best_f=best_f[-1], maximize=False)
        new_x = g.joint_optimize_manifold(acq, manifold, solver, q=1, num_restarts=num_restarts, raw_samples=raw_samples,
                                          bounds=bounds, inequality_constraints=ineq, equality_constraints=eq,
                                          options={'seed': seed + it})
        new_x = new_x.reshape(1, dim).to('cpu', torch.float64)
        new_y = ackley_sphere(g.Sphere(dim), new_x.numpy())
        worst = min(worst, feasibility(new_x[0]))
        x_data = torch.cat((x_data, new_x))
        y_data = torch.cat((y_data, torch.tensor([new_y], dtype=torch.float64)))
        best_f.append(min(best_f[-1], new_y))
        if verbose:
            print('iteration %2d  f(x) = %.5f  best = %.5f  x = %s  constraint margin %.2e'
                  % (it + 1, new_y, best_f[-1], np.round(new_x[0].numpy(), 4), feasibility(new_x[0])))
    return x_data, y_data, best_f, worst


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--constraint', default='inequality', choices=['bound', 'inequality', 'equality'])
    ap.add_argument('--solver', default='ctr', choices=['ctr', 'alm'])
    ap.add_argument('--iters', type=int, default=25)
    ap.add_argument('--seed', type=int, default=1234)
    a = ap.parse_args()
    run(a.constraint, a.solver, a.iters, seed=a.seed)
