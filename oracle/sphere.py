"""Oracle (CPU, fp64) for the sphere part of the hot path.  Test infrastructure only.

Follows ``BoManifolds/Riemannian_utils/sphere_utils_torch.py:12-55`` (distance),
``BoManifolds/kernel_utils/kernels_sphere.py:71-94`` (Gaussian kernel) and ``:112-134``
(Laplace kernel) of the reference, and pymanopt 0.2.x ``manifolds/sphere.py`` for the
manifold operations (third-party: PARITY UNPINNED, cross-checked against the
reference's own numpy formulas ``Riemannian_utils/sphere_utils.py:14-123``).
"""
import numpy as np
import torch

CLAMP_EPS = 1e-15  # sphere_utils_torch.py:53


def sphere_distance(x1, x2, diag=False):
    """acos(clamp(<x1_i, x2_j>)) -- sphere_utils_torch.py:29-55.

    The reference materialises the (N1,N2,D) broadcasts and uses a batch of 1xD @ Dx1
    ``bmm``; the arithmetic per pair is a plain D-term dot product, which is what the
    einsum below computes (same summation order for D <= 8; pinned on the golden vectors).
    """
    x1 = torch.as_tensor(x1)
    x2 = torch.as_tensor(x2)
    if diag is False:
        inner = torch.einsum('...id,...jd->...ij', x1, x2)
    else:
        inner = (x1 * x2).sum(-1, keepdim=True)  # (N,1), sphere_utils_torch.py:45-49
    inner = inner.clamp(-1. + CLAMP_EPS, 1. - CLAMP_EPS)
    return torch.acos(inner)


def sphere_distance_loop(x1, x2):
    """The reference's own op sequence (cat-broadcast + bmm), sphere_utils_torch.py:31-43.
    Used as the timed CPU baseline ("port") because it has the reference's memory traffic."""
    x1 = torch.as_tensor(x1)
    x2 = torch.as_tensor(x2)
    a = x1.unsqueeze(-2)
    b = x2.unsqueeze(-3)
    a = torch.cat(b.shape[-2] * [a], dim=-2)
    b = torch.cat(a.shape[-3] * [b], dim=-3)
    a = a.unsqueeze(-2)
    b = b.unsqueeze(-1)
    inner = torch.bmm(a.reshape(-1, 1, a.shape[-1]), b.reshape(-1, b.shape[-2], 1)).view(a.shape[:-2])
    inner = inner.clamp(-1. + CLAMP_EPS, 1. - CLAMP_EPS)
    return torch.acos(inner)


def sphere_gaussian_kernel(x1, x2, beta, diag=False, loop=False):
    """exp(-beta d^2) -- kernels_sphere.py:89-94 (beta already constrained, cast to double)."""
    d = sphere_distance_loop(x1, x2) if (loop and not diag) else sphere_distance(x1, x2, diag=diag)
    d2 = torch.mul(d, d)
    return torch.exp(-d2.mul(torch.as_tensor(beta, dtype=torch.float64)))


def sphere_laplace_kernel(x1, x2, lengthscale, diag=False):
    """exp(-d / l^2) -- kernels_sphere.py:129-134."""
    d = sphere_distance(x1, x2, diag=diag)
    ls = torch.as_tensor(lengthscale, dtype=torch.float64)
    return torch.exp(-d.div(ls * ls))


def beta_from_raw(raw_beta, beta_min):
    """gpytorch GreaterThan(beta_min).transform = beta_min + softplus(raw) (kernels_sphere.py:56-60)."""
    raw = torch.as_tensor(raw_beta, dtype=torch.float64)
    return beta_min + torch.nn.functional.softplus(raw)


# ----------------------------------------------------------------------------------------------
# Manifold operations: pymanopt 0.2.x Sphere (third-party, restated).  numpy, fp64, batched on
# the leading axis: x, u of shape (..., D).
# ----------------------------------------------------------------------------------------------

def inner(x, u, v):
    return np.sum(u * v, axis=-1)


def norm(x, u):
    return np.sqrt(np.sum(u * u, axis=-1))


def proj(x, h):
    """h - <x,h> x  (pymanopt Sphere.proj; also egrad2rgrad)."""
    return h - np.sum(x * h, axis=-1, keepdims=True) * x


egrad2rgrad = proj


def retr(x, u):
    y = x + u
    return y / np.linalg.norm(y, axis=-1, keepdims=True)


def dist(x, y):
    c = np.clip(np.sum(x * y, axis=-1), -1., 1.)
    return np.arccos(c)


def exp(x, u):
    """x cos|u| + u sin|u|/|u|, retraction when |u| <= 1e-3 (pymanopt Sphere.exp).
    Same closed form as the reference's numpy ``expmap`` (sphere_utils.py:33-36)."""
    nu = np.linalg.norm(u, axis=-1, keepdims=True)
    safe = np.where(nu > 1e-3, nu, 1.0)
    big = x * np.cos(nu) + u * np.sin(nu) / safe
    return np.where(nu > 1e-3, big, retr(x, u))


def log(x, y):
    """proj(x, y-x) rescaled to length dist(x,y) when dist > 1e-6 (pymanopt Sphere.log).
    Algebraically the reference's ``logmap`` (sphere_utils.py:60-63): (y - x cos t) t / sin t."""
    p = proj(x, y - x)
    d = dist(x, y)[..., None]
    npn = np.linalg.norm(p, axis=-1, keepdims=True)
    scale = np.where(d > 1e-6, d / np.where(npn > 0, npn, 1.0), 1.0)
    return p * scale


def transp(x, y, u):
    """Projection transport used by the solvers (pymanopt Sphere.transp)."""
    return proj(y, u)


def parallel_transport(x1, x2, v):
    """True great-circle parallel transport of v from T_x1 to T_x2, i.e. the operator of the
    reference's ``parallel_transport_operator`` (sphere_utils.py:93-123) applied to v."""
    xdir = log(x1, x2)
    n = np.linalg.norm(xdir, axis=-1, keepdims=True)
    small = n < 1e-16
    e = xdir / np.where(small, 1.0, n)
    ev = np.sum(e * v, axis=-1, keepdims=True)
    out = -x1 * np.sin(n) * ev + e * np.cos(n) * ev + v - e * ev
    return np.where(small, v, out)


def rand(rng, n, dim):
    """normalize(randn(D)) -- pymanopt Sphere.rand, used at gabo_sphere.py:109."""
    y = rng.standard_normal((n, dim))
    return y / np.linalg.norm(y, axis=-1, keepdims=True)


def ackley(x):
    """Ackley function on the sphere, base point (1,0,..,0): test_functions_sphere.py:34-65.
    x: (N, D) unit vectors -> (N,)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    D = x.shape[-1]
    base = np.zeros((1, D))
    base[0, 0] = 1.
    xp = log(np.broadcast_to(base, x.shape), x)[:, 1:]
    r = D - 1
    a, b, c = 20., 0.2, 2. * np.pi
    t1 = -a * np.exp(-b * np.sqrt(np.sum(xp ** 2, axis=-1) / r))
    t2 = -np.exp(np.sum(np.cos(c * xp) / r, axis=-1))
    return t1 + t2 + a + np.exp(1.)
