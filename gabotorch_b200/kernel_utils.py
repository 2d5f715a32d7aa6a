"""Geodesic kernels with the reference's class names and signatures (BoManifolds/kernel_utils), computed on the B200.

Mirrors ``BoManifolds/kernel_utils/kernels_sphere.py`` and ``kernels_spd.py`` of the reference: same constructor
arguments, same ``beta`` / ``lengthscale`` parameterisation (``beta = beta_min + softplus(raw_beta)``,
kernels_sphere.py:56-60), same ``forward`` signatures -- including the reference's inconsistency that the sphere
kernels take ``diag=`` while the SPD kernels take ``diagonal_distance=`` (kernels_sphere.py:71 vs kernels_spd.py:72).
``forward`` returns float64 like the reference (``self.beta.double()``, kernels_sphere.py:93) on the device of ``x1``.

The arithmetic is one fused CUDA launch per Gram matrix (``gabo_sphere_gram`` / ``gabo_spd_factor`` +
``gabo_spd_ai_gram``, include/gabo_b200.h); there is no CPU implementation behind these classes.  Gradients: the Gram
path is differentiable with respect to the kernel parameter (``raw_beta`` / ``raw_lengthscale``), which is what GP
hyper-parameter fitting needs.  The sphere kernels (``_SphereDistance``) and the SPD
affine-invariant Gaussian kernel (``_SpdAiDistance2``) also back-propagate to their inputs; the batched acquisition
optimiser does not go through autograd (closed-form Riemannian gradient in its kernels).
"""
import math

import torch

from . import _lib, ops
from ._compat import GreaterThan, Kernel


def _needs_param_grad(param):
    return torch.is_grad_enabled() and param.requires_grad


def _reject_input_grad(*xs):
    if torch.is_grad_enabled() and any(torch.is_tensor(x) and x.requires_grad for x in xs):
        raise NotImplementedError(
            'gabotorch_b200 kernels do not back-propagate to their inputs through autograd; the acquisition '
            'optimiser (manifold_optimization.gen_candidates_manifold) evaluates the Riemannian gradient in closed form')


def _host_result(x1, x2):
    """True when the caller works with host tensors and the result is wanted on the host (two distinct operands)."""
    return (torch.is_tensor(x1) and torch.is_tensor(x2) and not x1.is_cuda and not x2.is_cuda and x1 is not x2
            and x1.dim() == 2 and x2.dim() == 2)


def _finish(out, like):
    """Result on the caller's device (the reference returns CPU float64 for CPU inputs).  Host results land in pinned
    memory (torch's caching host allocator recycles the block), so the read-back runs at PCIe speed, not pageable speed."""
    if like.is_cuda or not out.is_cuda:
        return out
    if out.requires_grad:
        return out.to(like.device)
    host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return host


class _BetaKernel(Kernel):
    """Shared by the kernels parameterised with ``beta >= beta_min`` (kernels_sphere.py:30-69, kernels_spd.py:33-70)."""

    def __init__(self, beta_min, beta_prior=None, **kwargs):
        super().__init__(has_lengthscale=False, **kwargs)
        self.beta_min = beta_min
        self.register_parameter(name='raw_beta', parameter=torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, 1)))
        if beta_prior is not None:
            self.register_prior('beta_prior', beta_prior, lambda: self.beta, lambda v: self._set_beta(v))
        self.register_constraint('raw_beta', GreaterThan(self.beta_min))

    @property
    def beta(self):
        return self.raw_beta_constraint.transform(self.raw_beta)

    @beta.setter
    def beta(self, value):
        self._set_beta(value)

    def _set_beta(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_beta)
        self.initialize(raw_beta=self.raw_beta_constraint.inverse_transform(value))

    def _beta_scalar(self):
        b = self.beta
        if b.numel() != 1:
            raise NotImplementedError('batched beta (batch_shape != []) is not supported by the fused Gram kernels')
        return b


class _SphereDistance(torch.autograd.Function):
    """d_ij = acos(clamp(<x1_i, x2_j>)) from the fused kernel, differentiable with respect to the inputs the way the
    reference's op sequence is (sphere_utils_torch.py:29-55 under autograd): dd/dc = -1 / sqrt(1 - c^2) = -1 / sin d,
    zero where the clamp is active.  Forward is one launch of ``gabo_sphere_gram(KIND_DIST)``; the backward is one launch
    of ``gabo_weighted_points_sum`` per operand (weights -g_ij / sin d_ij formed inside the reduction)."""

    @staticmethod
    def forward(ctx, x1, x2):
        a, b = ops.to_dev64(x1), ops.to_dev64(x2)
        d = ops.sphere_gram(a, b, kind=_lib.KIND_DIST)
        ctx.save_for_backward(a, b, d)
        ctx.devs = (x1.device, x2.device, x1.dtype, x2.dtype)
        return d

    @staticmethod
    def backward(ctx, g):
        a, b, d = ctx.saved_tensors
        g = ops.to_dev64(g).contiguous()
        ga = ops.weighted_points_sum(g, b, transpose=False, dist=d) if ctx.needs_input_grad[0] else None
        gb = ops.weighted_points_sum(g, a, transpose=True, dist=d) if ctx.needs_input_grad[1] else None
        d1, d2, t1, t2 = ctx.devs
        return (None if ga is None else ga.to(device=d1, dtype=t1)), (None if gb is None else gb.to(device=d2, dtype=t2))


def _sphere_distance_with_grad(x1, x2, diag):
    if diag:
        raise NotImplementedError('input gradients of the diag=True branch are not provided')
    if x1.dim() != 2 or x2.dim() != 2:
        raise NotImplementedError('input gradients are provided for (N, D) inputs')
    return _SphereDistance.apply(x1, x2)


def _wants_input_grad(*xs):
    return torch.is_grad_enabled() and any(torch.is_tensor(x) and x.requires_grad for x in xs)


class _SpdAiDistance2(torch.autograd.Function):
    """d_AI(X1_i, X2_j)^2 for Mandel-vectorised inputs, differentiable with respect to both inputs.  Forward: the fused
    Gram kernel; backward: ``gabo_spd_ai_gram_backward`` (weighted sums of whitened log maps, one warp per point) and a
    Mandel pack.  The reference reaches the same gradient through torch.autograd over cholesky / inverse / bmm /
    symeig(eigenvectors=True) (spd_utils_torch.py:87-120)."""

    @staticmethod
    def forward(ctx, x1, x2, compute):
        a, b = ops.to_dev64(x1), ops.to_dev64(x2)
        d = ops.mandel_dim(a.shape[-1])
        flags = torch.zeros(1, dtype=torch.int32, device=a.device)
        f1, f2 = ops.spd_factor_pair(a, b, d, True, flags)
        dist = ops.spd_ai_gram_from_factors(f1, f2, d, kind=_lib.KIND_DIST, compute=compute)
        ops.check_spd_flags(flags)
        ctx.save_for_backward(f1, f2)
        ctx.meta = (d, compute, x1.device, x2.device, x1.dtype, x2.dtype)
        return dist * dist

    @staticmethod
    def backward(ctx, g):
        f1, f2 = ctx.saved_tensors
        d, compute, dev1, dev2, t1, t2 = ctx.meta
        g = ops.to_dev64(g)
        g1 = g2 = None
        if ctx.needs_input_grad[0]:
            g1 = ops.mandel_pack(ops.spd_ai_gram_backward(f1, f2, d, g, False, compute)).to(device=dev1, dtype=t1)
        if ctx.needs_input_grad[1]:
            g2 = ops.mandel_pack(ops.spd_ai_gram_backward(f2, f1, d, g, True, compute)).to(device=dev2, dtype=t2)
        return g1, g2, None


class _MandelUnpack(torch.autograd.Function):
    """Mandel vectors (n, dv) -> symmetric matrices (n, d, d); backward = Mandel pack of the symmetrised gradient."""

    @staticmethod
    def forward(ctx, v):
        ctx.meta = (v.device, v.dtype)
        return ops.mandel_unpack(ops.to_dev64(v))

    @staticmethod
    def backward(ctx, g):
        g = ops.to_dev64(g)
        dev, dt = ctx.meta
        return ops.mandel_pack(0.5 * (g + g.transpose(-1, -2))).to(device=dev, dtype=dt)


class _SpdLogm(torch.autograd.Function):
    """logm of SPD matrices (``gabo_spd_logm``) with the adjoint of its Frechet derivative (``gabo_spd_logm_backward``):
    what autograd gives the reference through ``logm_torch`` (symeig with eigenvectors, spd_utils_torch.py:13-30)."""

    @staticmethod
    def forward(ctx, m):
        m = ops.to_dev64(m)
        ctx.save_for_backward(m)
        return ops.spd_logm(m)

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return ops.spd_logm_backward(m, ops.to_dev64(g).contiguous())


class _FrobeniusDistance2(torch.autograd.Function):
    """||M1_i - M2_j + 1e-15||_F^2 (spd_utils_torch.py:156) for (n1, d, d), (n2, d, d); the backward
    2 sum_j g_ij (M1_i - M2_j + 1e-15) is a row sum and one ``gabo_weighted_points_sum`` per operand."""

    @staticmethod
    def forward(ctx, m1, m2):
        m1, m2 = ops.to_dev64(m1), ops.to_dev64(m2)
        ctx.save_for_backward(m1, m2)
        d = ops.frobenius_gram(m1, m2, kind=_lib.KIND_DIST)
        return d * d

    @staticmethod
    def backward(ctx, g):
        m1, m2 = ctx.saved_tensors
        g = ops.to_dev64(g).contiguous()
        n1, n2, dd = m1.shape[0], m2.shape[0], m1.shape[-1] * m1.shape[-2]
        f1, f2 = m1.reshape(n1, dd), m2.reshape(n2, dd)
        g1 = g2 = None
        if ctx.needs_input_grad[0]:
            g1 = 2.0 * (g.sum(1, keepdim=True) * (f1 + 1e-15) - ops.weighted_points_sum(g, f2, transpose=False))
            g1 = g1.reshape(m1.shape)
        if ctx.needs_input_grad[1]:
            g2 = 2.0 * (g.sum(0).unsqueeze(1) * (f2 - 1e-15) - ops.weighted_points_sum(g, f1, transpose=True))
            g2 = g2.reshape(m2.shape)
        return g1, g2


class _NestedSpdProject(torch.autograd.Function):
    """Mandel(W^T X W) (``gabo_nested_spd_project_f64``) with gradients for the inputs AND the projection matrix
    (``gabo_nested_spd_project_backward``): dX_n = W G_n W^T, dW = 2 sum_n X_n W G_n (nested_spd_utils.py:13-48 under
    autograd).  The gradient with respect to W is the EUCLIDEAN one; ``fit_gpytorch_manifold`` projects it onto the
    Grassmann tangent space as pymanopt's ``egrad2rgrad`` does."""

    @staticmethod
    def forward(ctx, x, w):
        xd, wd = ops.to_dev64(x), ops.to_dev64(w)
        ctx.save_for_backward(xd, wd)
        ctx.meta = (x.device, x.dtype, w.device, w.dtype, tuple(x.shape))
        return ops.nested_spd_project_f64(xd, wd)

    @staticmethod
    def backward(ctx, gy):
        xd, wd = ctx.saved_tensors
        xdev, xdt, wdev, wdt, xshape = ctx.meta
        gmat = ops.mandel_unpack(ops.to_dev64(gy).reshape(-1, gy.shape[-1]))
        want_x, want_w = ctx.needs_input_grad
        xm = ops.mandel_unpack(xd.reshape(-1, xd.shape[-1])) if want_w else None
        gx, gw = ops.nested_spd_project_backward(xm, wd, gmat, want_x=want_x, want_w=want_w)
        if gx is not None:
            gx = ops.mandel_pack(gx).reshape(xshape).to(device=xdev, dtype=xdt)
        if gw is not None:
            gw = gw.to(device=wdev, dtype=wdt)
        return gx, gw


def _dev64_keep_grad(x):
    """fp64 on the compute device WITHOUT leaving the autograd graph (``ops.to_dev64`` detaches)."""
    x = torch.as_tensor(x)
    return x.to(device=x.device if x.is_cuda else ops.device(), dtype=torch.float64)


def _rotate_to_north(x, axis):
    """R x for the rotation R that moves the unit vector ``axis`` to the north pole e_last along their geodesic
    (rotation_from_sphere_points_torch, sphere_utils_torch.py:58-93), applied in O(k) without forming R:
    R = I + sin(t) (y c^T - c y^T) + (cos(t) - 1) (y y^T + c c^T), y = e_last, cos t = <axis, y>, c = unit(axis - y cos t).
    Differentiable torch code on the device: used only when the axes are being fitted (requires_grad)."""
    ct = axis[-1].clamp(-1.0 + 1e-15, 1.0 - 1e-15)
    y = torch.zeros_like(axis)
    y[-1] = 1.0
    c = axis - y * ct
    c = c / c.norm()
    st = torch.sin(torch.acos(ct))
    xc, xy = x @ c, x[:, -1]
    return x + st * (xc[:, None] * y - xy[:, None] * c) + (ct - 1.0) * (xy[:, None] * y + xc[:, None] * c)


def _nested_sphere_project_autograd(x, axes, dists):
    """The nested-sphere projection chain S^{D-1} -> S^{d-1} (nested_spheres_utils.py:13-147) in differentiable device
    code, one level per axis: rotate the axis to the north pole, project onto the small circle at ``dist`` from it,
    drop the last coordinate and renormalise (with the reference's 1e-6 guards).  The forward values equal
    ``gabo_nested_sphere_project`` (tested); this form exists for the gradients with respect to the axes."""
    x = _dev64_keep_grad(x)
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    for axis, r in zip(axes, dists):
        a = _dev64_keep_grad(axis).reshape(-1).to(x.device)
        r = _dev64_keep_grad(r).reshape(()).to(x.device)
        xr = _rotate_to_north(x, a)
        dist = torch.acos(xr[:, -1].clamp(-1.0 + 1e-15, 1.0 - 1e-15))[:, None]
        north = torch.zeros_like(xr)
        north[:, -1] = 1.0
        # on the small circle, still in the rotated frame (the reference rotates back and forth: R^T then R[:-1] = identity)
        xn = (torch.sin(r) * xr + torch.sin(dist - r) * north) / (torch.sin(dist) + 1e-6)
        xs = xn[:, :-1] / (torch.sin(r) + 1e-6)
        x = xs / (xs.norm(dim=-1, keepdim=True) + 1e-6)
    return x.reshape(lead + (x.shape[-1],))


def _param_gram(dist_fn, param, power):
    """exp(-param * d^power) with autograd to ``param``: distances from the fused kernel, the exp in torch on-device."""
    d = dist_fn()
    p = param.double().to(d.device).reshape(())
    return torch.exp(-(d * d if power == 2 else d).mul(p))


class SphereGaussianKernel(_BetaKernel):
    """exp(-beta d(x1,x2)^2) on the sphere (kernels_sphere.py:15-94)."""

    def forward(self, x1, x2, diag=False, **params):
        beta = self._beta_scalar()
        if _wants_input_grad(x1, x2):      # autograd callers (acquisition via torch.autograd as in the reference)
            out = _param_gram(lambda: _sphere_distance_with_grad(x1, x2, diag), beta, 2)
        elif _needs_param_grad(self.raw_beta):
            out = _param_gram(lambda: ops.sphere_gram(x1, x2, kind=_lib.KIND_DIST, diag=diag), beta, 2)
        else:
            out = ops.sphere_gram(x1, x2, float(beta.detach()), _lib.KIND_GAUSS, diag=diag)
        return _finish(out, x1)


class SphereLaplaceKernel(Kernel):
    """exp(-d(x1,x2) / lengthscale^2) on the sphere (kernels_sphere.py:97-134)."""
    has_lengthscale = True

    def __init__(self, **kwargs):
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)

    def forward(self, x1, x2, diag=False, **params):
        ls = self.lengthscale.reshape(()).double()
        inv = 1.0 / (ls * ls)
        if _wants_input_grad(x1, x2):
            out = _param_gram(lambda: _sphere_distance_with_grad(x1, x2, diag), inv, 1)
        elif _needs_param_grad(self.raw_lengthscale):
            out = _param_gram(lambda: ops.sphere_gram(x1, x2, kind=_lib.KIND_DIST, diag=diag), inv, 1)
        else:
            out = ops.sphere_gram(x1, x2, float(inv.detach()), _lib.KIND_LAPLACE, diag=diag)
        return _finish(out, x1)


def _spd_diag_ones(x2):
    # diagonal_distance=True: the reference returns zero distances of shape (..., N, 1) (spd_utils_torch.py:72-75)
    return torch.ones(tuple(x2.shape[:-1]) + (1,), dtype=torch.float64, device=x2.device)


class SpdAffineInvariantGaussianKernel(_BetaKernel):
    """exp(-beta d_AI(X1,X2)^2) on SPD(d), inputs in Mandel notation (kernels_spd.py:17-100).

    ``compute`` selects the arithmetic of the per-pair Jacobi eigen-solve: ``'f32'`` (default) or ``'f64'``.
    """

    def __init__(self, beta_min, beta_prior=None, compute='f32', **kwargs):
        super().__init__(beta_min, beta_prior=beta_prior, **kwargs)
        self.compute = compute
        self._host_ws = ops.HostGramWorkspace()

    def _compute(self):
        return _lib.GABO_F64 if self.compute == 'f64' else _lib.GABO_F32

    def forward(self, x1, x2, diagonal_distance=False, **params):
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        beta = self._beta_scalar()
        if _wants_input_grad(x1, x2):      # autograd callers: exp(-beta d^2) on the differentiable squared distance
            if x1.dim() != 2 or x2.dim() != 2:
                raise NotImplementedError('input gradients are provided for (N, dv) Mandel inputs')
            d2 = _SpdAiDistance2.apply(x1, x2, self._compute())
            out = torch.exp(-d2 * beta.double().to(d2.device).reshape(()))
        elif _needs_param_grad(self.raw_beta):
            out = _param_gram(lambda: ops.spd_ai_gram(x1, x2, kind=_lib.KIND_DIST, compute=self._compute()), beta, 2)
        else:
            # host inputs: the per-pair kernel stores straight into a pinned host tensor (the PCIe transfer of the
            # Gram matrix overlaps its computation); the mirrored x1-is-x2 form stays on the device path
            if _host_result(x1, x2) and x1.dtype == torch.float64 and x2.dtype == torch.float64:
                return ops.spd_ai_gram_host(self._host_ws, x1, x2, float(beta.detach()), _lib.KIND_GAUSS,
                                            compute=self._compute())
            out = ops.spd_ai_gram(x1, x2, float(beta.detach()), _lib.KIND_GAUSS, compute=self._compute())
        return _finish(out, x1)


class SpdAffineInvariantLaplaceKernel(SpdAffineInvariantGaussianKernel):
    """exp(-beta d_AI(X1,X2)) (kernels_spd.py:103-187)."""

    def forward(self, x1, x2, diagonal_distance=False, **params):
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        beta = self._beta_scalar()
        if _wants_input_grad(x1, x2):      # d = sqrt(d^2): the reference's 1e-15 under the root keeps d'(0) finite
            if x1.dim() != 2 or x2.dim() != 2:
                raise NotImplementedError('input gradients are provided for (N, dv) Mandel inputs')
            d2 = _SpdAiDistance2.apply(x1, x2, self._compute())
            out = torch.exp(-torch.sqrt(d2) * beta.double().to(d2.device).reshape(()))
        elif _needs_param_grad(self.raw_beta):
            out = _param_gram(lambda: ops.spd_ai_gram(x1, x2, kind=_lib.KIND_DIST, compute=self._compute()), beta, 1)
        else:
            out = ops.spd_ai_gram(x1, x2, float(beta.detach()), _lib.KIND_LAPLACE, compute=self._compute())
        return _finish(out, x1)


class SpdFrobeniusGaussianKernel(Kernel):
    """exp(-||X1 - X2||_F^2 / lengthscale^2) on Mandel-vectorised symmetric matrices (kernels_spd.py:190-241)."""
    has_lengthscale = True

    def __init__(self, **kwargs):
        super().__init__(has_lengthscale=True, ard_num_dims=None, **kwargs)

    def _matrices(self, x):
        return ops.mandel_unpack(x)

    def _matrices_autograd(self, x):
        return _MandelUnpack.apply(x)

    def forward(self, x1, x2, diagonal_distance=False, **params):
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        ls = self.lengthscale.reshape(()).double()
        inv = 1.0 / (ls * ls)
        if _wants_input_grad(x1, x2):
            if x1.dim() != 2 or x2.dim() != 2:
                raise NotImplementedError('input gradients are provided for (N, dv) Mandel inputs')
            d2 = _FrobeniusDistance2.apply(self._matrices_autograd(x1), self._matrices_autograd(x2))
            return _finish(torch.exp(-d2 * inv.to(d2.device)), x1)
        m1 = self._matrices(x1)
        m2 = m1 if x2 is x1 else self._matrices(x2)
        if _needs_param_grad(self.raw_lengthscale):
            out = _param_gram(lambda: ops.frobenius_gram(m1, m2, kind=_lib.KIND_DIST), inv, 2)
        else:
            out = ops.frobenius_gram(m1, m2, float(inv.detach()), _lib.KIND_GAUSS)
        return _finish(out, x1)


class SpdLogEuclideanGaussianKernel(SpdFrobeniusGaussianKernel):
    """exp(-||logm X1 - logm X2||_F^2 / lengthscale^2) (kernels_spd.py:244-313)."""

    def _matrices(self, x):
        return ops.spd_logm(ops.mandel_unpack(x))

    def _matrices_autograd(self, x):
        return _SpdLogm.apply(_MandelUnpack.apply(x))


def _grassmann_rand(D, d):
    """pymanopt ``Grassmann(D, d).rand()``: the Q factor of a Gaussian matrix (kernels_nested_spd.py:75-78)."""
    q, _ = torch.linalg.qr(torch.randn(D, d, dtype=torch.float64))
    return q


class _SphereStub:
    """What the reference stores as ``raw_axis_S<k>_manifold`` (pymanopt ``Sphere(k)``)."""

    def __init__(self, n):
        self._n = n

    def rand(self):
        v = torch.randn(1, self._n, dtype=torch.float64)
        return (v / v.norm()).numpy()


class _GrassmannStub:
    """What the reference stores as ``raw_projection_matrix_manifold`` (pymanopt ``Grassmann``): ``rand`` and the sizes."""

    def __init__(self, n, p):
        self._n, self._p = n, p

    def rand(self):
        return _grassmann_rand(self._n, self._p).numpy()


class _NestedSpdMixin:
    """Projection parameter shared by the nested SPD kernels (kernels_nested_spd.py:74-100, :176-191):
    ``raw_projection_matrix`` lives on the Grassmann manifold G(dim, latent_dim) (``raw_projection_matrix_manifold``) and
    is fitted there by ``manifold_gp_fit.fit_gpytorch_manifold``; the projection back-propagates to it and to the inputs
    (``_NestedSpdProject``)."""

    def _init_projection(self, dim, latent_dim):
        self.dim = dim
        self.latent_dim = latent_dim
        self.raw_projection_matrix_manifold = _GrassmannStub(dim, latent_dim)
        w = _grassmann_rand(dim, latent_dim).to(torch.float32).repeat(*self.batch_shape, 1, 1)
        self.register_parameter(name='raw_projection_matrix', parameter=torch.nn.Parameter(w))

    @property
    def projection_matrix(self):
        return self.raw_projection_matrix

    @projection_matrix.setter
    def projection_matrix(self, value):
        self._set_projection_matrix(value)

    def _set_projection_matrix(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value)
        self.initialize(raw_projection_matrix=value.to(self.raw_projection_matrix))

    def _wants_autograd(self, *xs):
        return _wants_input_grad(*xs) or _needs_param_grad(self.raw_projection_matrix)

    #: rows from which ``projection='auto'`` streams the inputs through the tensor-core projection
    TENSOR_CORE_ROWS = 16384

    def _project(self, x, autograd=False):
        """Mandel vectors of SPD(dim) -> Mandel vectors of SPD(latent_dim): G3 -> P1 -> G3 of SURVEY section 8.

        Two device paths.  ``gabo_nested_spd_project_f64`` (fp64 in / out) for training-set-sized inputs: the projected
        matrices feed an affine-invariant distance whose error grows with their condition number.  The tensor-core GEMM
        ``gabo_nested_spd_project`` (3xTF32, fp32 in / out, ~1e-6 relative, HBM-bound) for the raw-sample screening of
        hd_gabo_spd.py:239-256 -- hundreds of thousands to millions of Mandel vectors whose kernel values are only ranked.
        ``self.projection``: 'auto' (tensor cores from ``TENSOR_CORE_ROWS`` rows when ``compute == 'f32'``), 'f64', 'tf32'."""
        w = self.raw_projection_matrix
        if w.dim() != 2:
            raise NotImplementedError('batched projection matrices (batch_shape != []) are not supported')
        if autograd:
            if x.dim() != 2:
                raise NotImplementedError('gradients are provided for (N, dv) Mandel inputs')
            return _NestedSpdProject.apply(x, w)
        mode = getattr(self, 'projection', 'auto')
        rows = int(x.numel() // x.shape[-1]) if x.shape[-1] else 0
        tensor = mode == 'tf32' or (mode == 'auto' and getattr(self, 'compute', 'f64') == 'f32'
                                    and rows >= self.TENSOR_CORE_ROWS)
        if tensor and self.latent_dim <= _lib.MAX_SPD_DIM:
            wd = w.detach().double()
            key = (wd.data_ptr(), w._version, str(wd.device))
            if getattr(self, '_pack_key', None) != key:              # operator packed once per value of W
                self._pack = ops.nested_projection_matrix(wd)
                self._pack_key = key
            flat = torch.as_tensor(x).reshape(-1, x.shape[-1])
            y = ops.nested_spd_project(flat, self.dim, self.latent_dim, self._pack)
            return y.double().reshape(tuple(x.shape[:-1]) + (y.shape[-1],))
        return ops.nested_spd_project_f64(x, w.detach().double())


class NestedSpdAffineInvariantGaussianKernel(_NestedSpdMixin, _BetaKernel):
    """exp(-beta d_AI(W^T X1 W, W^T X2 W)^2) for Mandel-vectorised SPD(dim) inputs (kernels_nested_spd.py:19-136)."""

    def __init__(self, dim, latent_dim, beta_min, beta_prior=None, compute='f32', projection='auto', **kwargs):
        _BetaKernel.__init__(self, beta_min, beta_prior=beta_prior, **kwargs)
        self.compute = compute
        self.projection = projection
        self._init_projection(dim, latent_dim)

    def forward(self, x1, x2, diagonal_distance=False, **params):
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        beta = self._beta_scalar()
        comp = _lib.GABO_F64 if self.compute == 'f64' else _lib.GABO_F32
        if self._wants_autograd(x1, x2):   # gradients for the inputs and / or the Grassmann parameter
            y1 = self._project(x1, autograd=True)
            y2 = y1 if x2 is x1 else self._project(x2, autograd=True)
            d2 = _SpdAiDistance2.apply(y1, y2, comp)
            return _finish(torch.exp(-d2 * beta.double().to(d2.device).reshape(())), x1)
        y1 = self._project(x1)
        y2 = y1 if x2 is x1 else self._project(x2)
        if _needs_param_grad(self.raw_beta):
            out = _param_gram(lambda: ops.spd_ai_gram(y1, y2, kind=_lib.KIND_DIST, compute=comp), beta, 2)
        else:
            out = ops.spd_ai_gram(y1, y2, float(beta.detach()), _lib.KIND_GAUSS, compute=comp)
        return _finish(out, x1)


class NestedSpdLogEuclideanGaussianKernel(_NestedSpdMixin, Kernel):
    """exp(-||logm(W^T X1 W) - logm(W^T X2 W)||_F^2 / lengthscale^2) (kernels_nested_spd.py:139-246)."""
    has_lengthscale = True

    def __init__(self, dim, latent_dim, **kwargs):
        Kernel.__init__(self, has_lengthscale=True, ard_num_dims=None, **kwargs)
        self._init_projection(dim, latent_dim)

    def forward(self, x1, x2, diagonal_distance=False, **params):
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        ls = self.lengthscale.reshape(()).double()
        inv = 1.0 / (ls * ls)
        if self._wants_autograd(x1, x2):
            a1 = _SpdLogm.apply(_MandelUnpack.apply(self._project(x1, autograd=True)))
            a2 = a1 if x2 is x1 else _SpdLogm.apply(_MandelUnpack.apply(self._project(x2, autograd=True)))
            d2 = _FrobeniusDistance2.apply(a1, a2)
            return _finish(torch.exp(-d2 * inv.to(d2.device)), x1)
        m1 = ops.spd_logm(ops.mandel_unpack(self._project(x1)))
        m2 = m1 if x2 is x1 else ops.spd_logm(ops.mandel_unpack(self._project(x2)))
        if _needs_param_grad(self.raw_lengthscale):
            out = _param_gram(lambda: ops.frobenius_gram(m1, m2, kind=_lib.KIND_DIST), inv, 2)
        else:
            out = ops.frobenius_gram(m1, m2, float(inv.detach()), _lib.KIND_GAUSS)
        return _finish(out, x1)


class NestedSphereGaussianKernel(_BetaKernel):
    """exp(-beta d(p(x1), p(x2))^2) with p the nested-sphere projection S^{dim-1} -> S^{latent_dim-1}
    (kernels_nested_sphere.py:19-152).  One axis parameter ``raw_axis_S<k>`` per level k = dim .. latent_dim+1 (each on
    its own sphere manifold ``raw_axis_S<k>_manifold``, fitted by ``manifold_gp_fit.fit_gpytorch_manifold``) and the
    distances to the axes fixed at pi/2, as in the reference."""

    def __init__(self, dim, latent_dim, beta_min, beta_prior=None, **kwargs):
        super().__init__(beta_min, beta_prior=beta_prior, **kwargs)
        self.dim = dim
        self.latent_dim = latent_dim
        for k in range(dim, latent_dim, -1):
            axis = torch.randn(1, k)
            axis = (axis / torch.norm(axis)).repeat(*self.batch_shape, 1, 1)
            self.register_parameter(name='raw_axis_S%d' % k, parameter=torch.nn.Parameter(axis))
            setattr(self, 'raw_axis_S%d_manifold' % k, _SphereStub(k))
        self.distances_to_axis = [math.pi / 2 * torch.ones(1, 1) for _ in range(dim, latent_dim, -1)]

    @property
    def axes(self):
        return [self._parameters['raw_axis_S%d' % k] for k in range(self.dim, self.latent_dim, -1)]

    @axes.setter
    def axes(self, values_list):
        self._set_axes(values_list)

    def _set_axes(self, values_list):
        for k in range(self.dim, self.latent_dim, -1):
            name = 'raw_axis_S%d' % k
            value = values_list[self.dim - k]
            if not torch.is_tensor(value):
                value = torch.as_tensor(value)
            self.initialize(**{name: value.to(self._parameters[name]).reshape(self._parameters[name].shape)})

    def _project(self, x):
        return ops.nested_sphere_project(x, [a.detach().double() for a in self.axes], self.distances_to_axis)

    def forward(self, x1, x2, diag=False, **params):
        beta = self._beta_scalar()
        if _wants_input_grad(x1, x2) or (torch.is_grad_enabled() and any(a.requires_grad for a in self.axes)):
            # gradients for the inputs and / or the axes: the projection chain in differentiable device code
            q1 = _nested_sphere_project_autograd(x1, self.axes, self.distances_to_axis)
            q2 = q1 if x2 is x1 else _nested_sphere_project_autograd(x2, self.axes, self.distances_to_axis)
            out = _param_gram(lambda: _sphere_distance_with_grad(q1, q2, diag), beta, 2)
            return _finish(out, x1)
        p1 = self._project(x1)
        p2 = p1 if x2 is x1 else self._project(x2)
        if _needs_param_grad(self.raw_beta):
            out = _param_gram(lambda: ops.sphere_gram(p1, p2, kind=_lib.KIND_DIST, diag=diag), beta, 2)
        else:
            out = ops.sphere_gram(p1, p2, float(beta.detach()), _lib.KIND_GAUSS, diag=diag)
        return _finish(out, x1)
