"""Reconstruction-parameter fit of the nested-sphere mapping.

Mirrors ``BoManifolds/nested_mappings/nested_spheres_optimization.py`` of the reference:
``min_error_reconstruction_cost`` (:20-38) and ``optimize_reconstruction_parameters_nested_sphere`` (:41-98) -- the step
of hd_gabo_sphere.py (:196-199) between the latent acquisition optimiser and ``projection_from_subsphere_to_sphere``: the
distances-to-axis r_k in (0, pi) of every level are chosen so that the data reconstructed from their latent projections are
as close as possible (sum of squared geodesic distances) to the data.  The reference treats the r_k through a sigmoid
(gpytorch ``Interval(0, pi)``) as a product of ``Euclidean(1)`` manifolds, screens ``nb_init_candidates`` random draws and
hands the best to a pymanopt solver (``TrustRegions()`` in the example).

Here the cost and its gradient are device work: the inverse chain (rotation of the north pole to each axis applied in O(k),
nested_spheres_utils.py:149-213) and the row-wise geodesic distance run as fp64 tensor code on the B200 for ALL candidates at
once (one pass, one read-back), and for the solver's evaluations with torch.autograd over the same code; the solver itself is
the host loop of ``manifold_gp_fit`` (pymanopt ``TrustRegions`` / ``ConjugateGradient`` semantics, finite-difference
Hessian), a handful of scalars.  The SPD counterpart (``nested_spd_optimization.py:95-186``: augmented Lagrangian over
Grassmann x SPD(D - d) x Sphere x R with affine-invariant distances on SPD(20)) is NOT provided: its cost lives on matrices
beyond the d <= 8 register kernels of this package.
"""
import math

import numpy as np
import torch

from . import ops
from .kernel_utils import _dev64_keep_grad
from .manifold_gp_fit import EuclideanParam, ProductParam, solve_on_manifold


def _rotate(v, p_from, p_to):
    """R v for the rotation R that moves the unit vector ``p_from`` to ``p_to`` along their geodesic
    (rotation_from_sphere_points_torch, sphere_utils_torch.py:58-93), applied without forming R; v: (..., N, k)."""
    ct = (p_from * p_to).sum().clamp(-1.0 + 1e-15, 1.0 - 1e-15)
    c = p_from - p_to * ct
    c = c / c.norm()
    st = torch.sin(torch.acos(ct))
    vc, vy = (v * c).sum(-1, keepdim=True), (v * p_to).sum(-1, keepdim=True)
    return v + st * (vc * p_to - vy * c) + (ct - 1.0) * (vy * p_to + vc * c)


def _reconstruct(x_sub, axes, dists):
    """projection_from_subsphere_to_sphere (nested_spheres_utils.py:184-213), last level only; ``dists``: (..., r) so that a
    leading candidate dimension broadcasts: x_sub (N, d) -> (..., N, D)."""
    x = x_sub
    nb = len(axes)
    for s in range(nb):
        axis = axes[nb - s - 1]
        r = dists[..., nb - s - 1][..., None, None]
        north = torch.zeros_like(axis)
        north[-1] = 1.0
        lifted = torch.cat([torch.sin(r) * x, torch.cos(r) * torch.ones_like(x[..., :1])], dim=-1)
        x = _rotate(lifted, north, axis)
    return x


def _cost_from_distances(x_data, x_sub, axes, dists):
    xr = _reconstruct(x_sub, axes, dists)
    inner = (x_data * xr).sum(-1).clamp(-1.0 + 1e-15, 1.0 - 1e-15)      # sphere_distance_torch(diag=True)
    d = torch.acos(inner)
    return (d * d).sum(-1)


def min_error_reconstruction_cost(x_data, x_subsphere, sphere_axes, sphere_distances):
    """Sum of squared geodesic distances between the data and their reconstruction from the subsphere
    (nested_spheres_optimization.py:20-38).  Differentiable device code; returns a 0-d fp64 tensor."""
    xd = _dev64_keep_grad(x_data)
    xs = _dev64_keep_grad(x_subsphere).to(xd.device)
    axes = [_dev64_keep_grad(a).reshape(-1).to(xd.device) for a in sphere_axes]
    dists = torch.stack([_dev64_keep_grad(r).reshape(()).to(xd.device) for r in sphere_distances])
    return _cost_from_distances(xd, xs, axes, dists)


def optimize_reconstruction_parameters_nested_sphere(x_data, x_subsphere, sphere_axes, solver, nb_init_candidates=100):
    """Distances-to-axis of ``projection_from_subsphere_to_sphere`` that minimise the reconstruction error
    (nested_spheres_optimization.py:41-98).  Returns a list of (1,) float32 tensors like the reference
    (``radius_constraint.transform(torch.Tensor(distance))``)."""
    xd = ops.to_dev64(x_data)
    xs = ops.to_dev64(x_subsphere)
    axes = [ops.to_dev64(a).reshape(-1) for a in sphere_axes]
    nlev = len(axes)
    if xd.shape[-1] - xs.shape[-1] != nlev:
        raise ValueError('need one axis per level: data on S^%d, subsphere data on S^%d, %d axes'
                         % (xd.shape[-1] - 1, xs.shape[-1] - 1, nlev))
    manifold = ProductParam([EuclideanParam(1) for _ in range(nlev)])

    def transform(p):                                   # gpytorch Interval(0, pi): sigmoid(raw) * pi
        return torch.sigmoid(p) * math.pi

    # candidate screening: every candidate in one pass over the device code, one read-back
    cands = [manifold.rand() for _ in range(int(nb_init_candidates))]
    raw = torch.from_numpy(np.array([[float(c[0]) for c in cand] for cand in cands])).to(xd.device)
    with torch.no_grad():
        vals = _cost_from_distances(xd, xs, axes, transform(raw)).cpu().numpy()
    x0 = cands[int(np.argmin(vals))]

    def cost(x):
        with torch.no_grad():
            p = torch.tensor([float(v[0]) for v in x], dtype=torch.float64, device=xd.device)
            return float(_cost_from_distances(xd, xs, axes, transform(p)))

    def cost_grad(x):
        p = torch.tensor([float(v[0]) for v in x], dtype=torch.float64, device=xd.device, requires_grad=True)
        with torch.enable_grad():
            f = _cost_from_distances(xd, xs, axes, transform(p))
            f.backward()
        g = p.grad.cpu().numpy()
        return float(f.detach()), [np.array([gi]) for gi in g]

    opt, log = solve_on_manifold(manifold, cost, cost_grad, x0, solver)
    out = [transform(torch.tensor(np.asarray(v), dtype=torch.float32).reshape(1)) for v in opt]
    optimize_reconstruction_parameters_nested_sphere.last_log = dict(log, start_cost=float(vals.min()))
    return out
