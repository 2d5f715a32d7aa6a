"""pymanopt-shaped manifold objects whose operations run on the B200 (M1 / M2 of SURVEY.md section 8).

The reference hands ``pymanopt.manifolds.Sphere`` / ``PositiveDefinite`` objects to its solvers and test functions
(examples/bo_sphere/benchmark_examples/gabo_sphere.py:87, examples/bo_spd/benchmark_examples/gabo_spd.py:98-102) and
calls ``rand / inner / norm / dist / proj / egrad2rgrad / retr / exp / log / transp / zerovec`` on them one point at a
time.  These classes keep that duck-typed surface (same method names and argument order, numpy in -> numpy out for a
single point) and add batching: every method also accepts ``(n, ...)`` stacks of points (numpy or torch), which is one
CUDA launch (``gabo_sphere_op`` / ``gabo_spd_op``, include/gabo_b200.h) instead of n Python calls.
Attributes read by the reference's examples and test functions are kept: ``_shape``, ``_n``, ``dim``, ``typicaldist``,
and the user-set ``min_eig`` / ``max_eig`` of the SPD manifold (Riemannian_utils/spd_utils.py:299,
BO_test_functions/test_functions_spd.py:35).
"""
import numpy as np
import torch

from . import _lib, ops


def _wrap(fn):
    """numpy in -> numpy out, torch in -> torch out (on the input's device)."""
    def call(self, *args):
        first = args[0]
        is_np = not torch.is_tensor(first)
        out = fn(self, *[torch.as_tensor(a) for a in args])
        if is_np:
            return out.cpu().numpy()
        return out if first.is_cuda else out.to(first.device)
    call.__name__ = fn.__name__
    call.__doc__ = fn.__doc__
    return call


class Sphere:
    """Unit sphere of R^n (pymanopt ``Sphere(n)``): points are unit vectors of length n."""

    def __init__(self, *shape):
        if len(shape) != 1:
            raise NotImplementedError('only Sphere(n) (vectors) is supported')
        self._shape = tuple(shape)
        self._n = shape[0]
        self.name = 'Sphere manifold of %d-vectors' % shape[0]

    def __str__(self):
        return self.name

    @property
    def dim(self):
        return self._n - 1

    @property
    def typicaldist(self):
        return np.pi

    def rand(self):
        y = np.random.randn(*self._shape)
        return y / np.linalg.norm(y)

    def rand_batch(self, n, generator=None, device=None):
        """n i.i.d. uniform points, (n, dim) fp64 on the compute device."""
        dev = device or ops.device()
        y = torch.randn(n, self._n, dtype=torch.float64, device=dev, generator=generator)
        return y / y.norm(dim=-1, keepdim=True)

    def randvec(self, x):
        h = np.random.randn(*np.shape(x))
        p = self.proj(np.asarray(x), h)
        return p / np.linalg.norm(p, axis=-1, keepdims=True)

    def zerovec(self, x):
        return np.zeros(np.shape(x)) if not torch.is_tensor(x) else torch.zeros_like(x)

    @_wrap
    def inner(self, x, u, v):
        return (ops.to_dev64(u) * ops.to_dev64(v)).sum(-1)

    @_wrap
    def norm(self, x, u):
        return ops.to_dev64(u).norm(dim=-1)

    @_wrap
    def dist(self, x, y):
        return ops.sphere_dist(x, y)

    @_wrap
    def proj(self, x, h):
        return ops.sphere_op(_lib.OP_PROJ, x, h)

    egrad2rgrad = proj

    @_wrap
    def ehess2rhess(self, x, egrad, ehess, u):
        dev = ops.to_dev64(x)
        eg, eh, uu = ops.to_dev64(egrad), ops.to_dev64(ehess), ops.to_dev64(u)
        return ops.sphere_op(_lib.OP_PROJ, dev, eh) - (dev * eg).sum(-1, keepdim=True) * uu

    @_wrap
    def retr(self, x, u):
        return ops.sphere_op(_lib.OP_RETR, x, u)

    @_wrap
    def exp(self, x, u):
        return ops.sphere_op(_lib.OP_EXP, x, u)

    @_wrap
    def log(self, x, y):
        return ops.sphere_op(_lib.OP_LOG, x, y)

    @_wrap
    def transp(self, x, y, u):
        return ops.sphere_op(_lib.OP_TRANSP, y, u)

    @_wrap
    def parallel_transport(self, x, y, u):
        """Great-circle parallel transport (the reference's parallel_transport_operator, sphere_utils.py:93-123)."""
        return ops.sphere_op(_lib.OP_PTRANSP, x, y, u)

    def pairmean(self, x, y):
        m = np.asarray(x) + np.asarray(y)
        return m / np.linalg.norm(m, axis=-1, keepdims=True)


class PositiveDefinite:
    """SPD(n) with the affine-invariant metric (pymanopt ``PositiveDefinite(n)``), n <= 8."""

    def __init__(self, n, k=1):
        if k != 1:
            raise NotImplementedError('only PositiveDefinite(n, k=1) is supported')
        if n > _lib.MAX_SPD_DIM:
            raise NotImplementedError('SPD(%d): the in-register eigen-solvers support n <= %d' % (n, _lib.MAX_SPD_DIM))
        self._n = n
        self._k = k
        self._shape = (n, n)
        self.name = 'Manifold of positive definite (%d x %d) matrices' % (n, n)

    def __str__(self):
        return self.name

    @property
    def dim(self):
        return self._n * (self._n + 1) // 2

    @property
    def typicaldist(self):
        return np.sqrt(self.dim)

    def rand(self):
        # pymanopt PositiveDefinite.rand: eigenvalues 1 + U[0,1), orthogonal factor from QR of a Gaussian matrix.
        # (The reference's examples replace this method by spd_sample, gabo_spd.py:100-102.)
        d = np.ones(self._n) + np.random.rand(self._n)
        u, _ = np.linalg.qr(np.random.randn(self._n, self._n))
        return u @ np.diag(d) @ u.T

    def rand_batch(self, n, generator=None, device=None, min_eig=None, max_eig=None):
        """n i.i.d. SPD matrices with the law of the reference's spd_sample (spd_utils.py:290-306) when min_eig /
        max_eig are given (or set on the manifold), else pymanopt's rand law.  (n, d, d) fp64 on the compute device."""
        dev = device or ops.device()
        lo = min_eig if min_eig is not None else getattr(self, 'min_eig', None)
        hi = max_eig if max_eig is not None else getattr(self, 'max_eig', None)
        if lo is None or hi is None:
            lo, hi = 1.0, 2.0
        lam = lo + (hi - lo) * torch.rand(n, self._n, dtype=torch.float64, device=dev, generator=generator)
        q, _ = torch.linalg.qr(torch.randn(n, self._n, self._n, dtype=torch.float64, device=dev, generator=generator))
        m = (q * lam.unsqueeze(-2)) @ q.transpose(-1, -2)
        return 0.5 * (m + m.transpose(-1, -2))

    def randvec(self, x):
        u = np.random.randn(*np.shape(x))
        u = 0.5 * (u + np.swapaxes(u, -1, -2))
        return u / np.asarray(self.norm(np.asarray(x), u))[..., None, None]

    def zerovec(self, x):
        return np.zeros(np.shape(x)) if not torch.is_tensor(x) else torch.zeros_like(x)

    @_wrap
    def inner(self, x, u, v):
        return ops.spd_scalar(2, x, u, v)

    @_wrap
    def norm(self, x, u):
        return ops.spd_scalar(1, x, u)

    @_wrap
    def dist(self, x, y):
        return ops.spd_scalar(0, x, y)

    @_wrap
    def proj(self, x, g):
        return ops.spd_op(_lib.OP_PROJ, x, g)

    @_wrap
    def egrad2rgrad(self, x, g):
        return ops.spd_op(_lib.OP_EGRAD2RGRAD, x, g)

    @_wrap
    def exp(self, x, u):
        return ops.spd_op(_lib.OP_EXP, x, u)

    retr = exp

    @_wrap
    def log(self, x, y):
        return ops.spd_op(_lib.OP_LOG, x, y)

    @_wrap
    def transp(self, x, y, u):
        return ops.spd_op(_lib.OP_TRANSP, y, u)

    @_wrap
    def parallel_transport(self, x, y, u):
        """E U E^T with E = (Y X^-1)^(1/2) (the reference's parallel_transport_operator, spd_utils.py:200-213)."""
        return ops.spd_op(_lib.OP_PTRANSP, x, y, u)
