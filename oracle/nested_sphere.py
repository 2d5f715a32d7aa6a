"""Oracle (CPU, fp64) for the nested-sphere projection of HD-GaBO on spheres.  Test infrastructure only.

Follows ``BoManifolds/Riemannian_utils/sphere_utils_torch.py:58-93`` (``rotation_from_sphere_points_torch``),
``BoManifolds/nested_mappings/nested_spheres_utils.py:13-147`` (projection to a nested sphere, identification with
the next subsphere, chain over several levels), ``:149-213`` (the inverse chain) and ``kernel_utils/kernels_nested_sphere.py:129-152`` (the kernel).
Pinned on the reference's own code through ``tests/golden`` (``nsph_*`` arrays).
"""
import math

import torch

from . import sphere as _sph

CLAMP_EPS = 1e-15
DIV_EPS = 1e-6  # nested_spheres_utils.py:58, :107, :112


def rotation_from_sphere_points(x, y):
    """Rotation moving x to y along the geodesic (sphere_utils_torch.py:73-93); x, y: (1, k)."""
    x = x.reshape(1, -1)
    y = y.reshape(1, -1)
    k = x.shape[1]
    inner = torch.mm(x, y.T).clamp(-1. + CLAMP_EPS, 1. - CLAMP_EPS)
    c_vec = x - y * inner
    c_vec = c_vec / torch.norm(c_vec)
    return (torch.eye(k, dtype=inner.dtype)
            + torch.sin(torch.acos(inner)) * (torch.mm(y.T, c_vec) - torch.mm(c_vec.T, y))
            + (inner - 1.) * (torch.mm(y.T, y) + torch.mm(c_vec.T, c_vec)))


def projection_to_nested_sphere(x, axis, dist_to_axis):
    """nested_spheres_utils.py:13-67: closest points at distance r from the axis, in the coordinates of S^{k-1}."""
    x = torch.as_tensor(x, dtype=torch.float64)
    axis = torch.as_tensor(axis, dtype=torch.float64).reshape(1, -1)
    r = torch.as_tensor(dist_to_axis, dtype=torch.float64).reshape(1, 1)
    k = x.shape[-1]
    north = torch.zeros_like(axis)
    north[:, -1] = 1.
    rot = rotation_from_sphere_points(axis, north)
    x_rot = torch.mm(rot, x.T).T
    d_axis = _sph.sphere_distance(x_rot, north).repeat((1, k))
    x_ns_rot = torch.sin(r) * x_rot + torch.sin(d_axis - r) * north
    x_ns_rot = x_ns_rot / (torch.sin(d_axis) + DIV_EPS)
    return torch.mm(rot.T, x_ns_rot.T).T


def projection_from_subsphere_to_next_sphere(x_sub, axis, dist_to_axis):
    """nested_spheres_utils.py:149-180: (N, k-1) on S^{k-2} -> (N, k) on the nested sphere of S^{k-1}."""
    x_sub = torch.as_tensor(x_sub, dtype=torch.float64)
    axis = torch.as_tensor(axis, dtype=torch.float64).reshape(1, -1)
    r = torch.as_tensor(dist_to_axis, dtype=torch.float64).reshape(1, 1)
    north = torch.zeros_like(axis)
    north[:, -1] = 1.
    rot = rotation_from_sphere_points(north, axis)
    cos_vec = torch.cos(r) * torch.ones(x_sub.shape[0], 1, dtype=torch.float64)
    return torch.mm(rot, torch.cat((torch.sin(r) * x_sub, cos_vec), 1).T).T


def projection_from_subsphere_to_sphere(x_sub, axes, dists):
    """nested_spheres_utils.py:182-213: axes in the order of the projection, consumed last to first."""
    out = [torch.as_tensor(x_sub, dtype=torch.float64)]
    for a, r in zip(reversed(list(axes)), reversed(list(dists))):
        out.append(projection_from_subsphere_to_next_sphere(out[-1], a, r))
    return out


def projection_to_next_subsphere(x, axis, dist_to_axis):
    """nested_spheres_utils.py:13-67 followed by :70-118: points (N, k) on S^{k-1} -> (N, k-1) on S^{k-2}."""
    x = torch.as_tensor(x, dtype=torch.float64)
    axis = torch.as_tensor(axis, dtype=torch.float64).reshape(1, -1)
    r = torch.as_tensor(dist_to_axis, dtype=torch.float64).reshape(1, 1)
    k = x.shape[-1]
    north = torch.zeros_like(axis)
    north[:, -1] = 1.
    rot = rotation_from_sphere_points(axis, north)
    x_rot = torch.mm(rot, x.T).T
    d_axis = _sph.sphere_distance(x_rot, north).repeat((1, k))
    x_ns_rot = torch.sin(r) * x_rot + torch.sin(d_axis - r) * north
    x_ns_rot = x_ns_rot / (torch.sin(d_axis) + DIV_EPS)
    x_ns = torch.mm(rot.T, x_ns_rot.T).T                                   # back on the nested sphere of S^{k-1}
    x_sub = torch.mm(rot[:-1, :], x_ns.T).T / (torch.sin(r) + DIV_EPS)     # identified with S^{k-2}
    norm = torch.norm(x_sub, dim=-1, keepdim=True)
    return x_sub / (norm + DIV_EPS)


def projection_from_sphere_to_subsphere(x, axes, dists):
    """nested_spheres_utils.py:120-147: list of the projections on every level; the kernel uses the last one."""
    out = [torch.as_tensor(x, dtype=torch.float64)]
    for a, r in zip(axes, dists):
        out.append(projection_to_next_subsphere(out[-1], a, r))
    return out


def nested_sphere_gaussian_kernel(x1, x2, axes, dists, beta, diag=False):
    """kernels_nested_sphere.py:143-152."""
    p1 = projection_from_sphere_to_subsphere(x1, axes, dists)[-1]
    p2 = projection_from_sphere_to_subsphere(x2, axes, dists)[-1]
    d = _sph.sphere_distance(p1, p2, diag=diag)
    return torch.exp(-torch.mul(d, d).mul(torch.as_tensor(beta, dtype=torch.float64)))


HALF_PI = math.pi / 2
