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
    zero where the clamp is active.  Forward is one launch of ``gabo_sphere_gram(KIND_DIST)``; the backward is two
    device matrix products on the (N1, N2) weight matrix  -g_ij / sin d_ij."""

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
        lo = math.acos(1.0 - 1e-15)                       # clamp active: d == acos(1 - 1e-15) or pi - that
        live = (d > lo * (1 + 1e-9)) & (d < math.pi - lo * (1 + 1e-9))
        w = torch.where(live, -g.to(d) / torch.sin(d), torch.zeros_like(d))
        ga = torch.matmul(w, b) if ctx.needs_input_grad[0] else None
        gb = torch.matmul(w.transpose(-1, -2), a) if ctx.needs_input_grad[1] else None
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
        _reject_input_grad(x1, x2)
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        beta = self._beta_scalar()
        if _needs_param_grad(self.raw_beta):
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

    def forward(self, x1, x2, diagonal_distance=False, **params):
        _reject_input_grad(x1, x2)
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        ls = self.lengthscale.reshape(()).double()
        inv = 1.0 / (ls * ls)
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


def _grassmann_rand(D, d):
    """pymanopt ``Grassmann(D, d).rand()``: the Q factor of a Gaussian matrix (kernels_nested_spd.py:75-78)."""
    q, _ = torch.linalg.qr(torch.randn(D, d, dtype=torch.float64))
    return q


class _GrassmannStub:
    """What the reference stores as ``raw_projection_matrix_manifold`` (pymanopt ``Grassmann``): ``rand`` and the sizes."""

    def __init__(self, n, p):
        self._n, self._p = n, p

    def rand(self):
        return _grassmann_rand(self._n, self._p).numpy()


class _NestedSpdMixin:
    """Projection parameter shared by the nested SPD kernels (kernels_nested_spd.py:74-100, :176-191).

    Deviation: ``raw_projection_matrix`` is created with ``requires_grad=False`` -- in the reference it is fitted on the
    Grassmann manifold by ``fit_gpytorch_manifold`` (GP fitting is outside this package's scope); asking for
    its gradient raises instead of silently returning none."""

    def _init_projection(self, dim, latent_dim):
        self.dim = dim
        self.latent_dim = latent_dim
        self.raw_projection_matrix_manifold = _GrassmannStub(dim, latent_dim)
        w = _grassmann_rand(dim, latent_dim).to(torch.float32).repeat(*self.batch_shape, 1, 1)
        self.register_parameter(name='raw_projection_matrix', parameter=torch.nn.Parameter(w, requires_grad=False))

    @property
    def projection_matrix(self):
        return self.raw_projection_matrix

    @projection_matrix.setter
    def projection_matrix(self, value):
        self._set_projection_matrix(value)

    def _set_projection_matrix(self, value):
        self.initialize(raw_projection_matrix=value)

    def _project(self, x):
        """Mandel vectors of SPD(dim) -> Mandel vectors of SPD(latent_dim): G3 -> P1 -> G3 of SURVEY section 8."""
        w = self.raw_projection_matrix
        if torch.is_grad_enabled() and w.requires_grad:
            raise NotImplementedError('gradients with respect to the projection matrix are not provided '
                                      '(fit_gpytorch_manifold is outside the scope of gabotorch_b200)')
        if w.dim() != 2:
            raise NotImplementedError('batched projection matrices (batch_shape != []) are not supported')
        return ops.nested_spd_project_f64(x, w.detach().double())


class NestedSpdAffineInvariantGaussianKernel(_NestedSpdMixin, _BetaKernel):
    """exp(-beta d_AI(W^T X1 W, W^T X2 W)^2) for Mandel-vectorised SPD(dim) inputs (kernels_nested_spd.py:19-136)."""

    def __init__(self, dim, latent_dim, beta_min, beta_prior=None, compute='f32', **kwargs):
        _BetaKernel.__init__(self, beta_min, beta_prior=beta_prior, **kwargs)
        self.compute = compute
        self._init_projection(dim, latent_dim)

    def forward(self, x1, x2, diagonal_distance=False, **params):
        _reject_input_grad(x1, x2)
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        beta = self._beta_scalar()
        comp = _lib.GABO_F64 if self.compute == 'f64' else _lib.GABO_F32
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
        _reject_input_grad(x1, x2)
        if diagonal_distance is True:
            return _spd_diag_ones(x2)
        ls = self.lengthscale.reshape(()).double()
        inv = 1.0 / (ls * ls)
        m1 = ops.spd_logm(ops.mandel_unpack(self._project(x1)))
        m2 = m1 if x2 is x1 else ops.spd_logm(ops.mandel_unpack(self._project(x2)))
        if _needs_param_grad(self.raw_lengthscale):
            out = _param_gram(lambda: ops.frobenius_gram(m1, m2, kind=_lib.KIND_DIST), inv, 2)
        else:
            out = ops.frobenius_gram(m1, m2, float(inv.detach()), _lib.KIND_GAUSS)
        return _finish(out, x1)


class NestedSphereGaussianKernel(_BetaKernel):
    """exp(-beta d(p(x1), p(x2))^2) with p the nested-sphere projection S^{dim-1} -> S^{latent_dim-1}
    (kernels_nested_sphere.py:19-152).  One axis parameter ``raw_axis_S<k>`` per level k = dim .. latent_dim+1 and the
    distances to the axes fixed at pi/2, as in the reference; the axes are created with ``requires_grad=False`` (they
    are fitted on sphere manifolds by ``fit_gpytorch_manifold`` in the reference, which is outside this package)."""

    def __init__(self, dim, latent_dim, beta_min, beta_prior=None, **kwargs):
        super().__init__(beta_min, beta_prior=beta_prior, **kwargs)
        self.dim = dim
        self.latent_dim = latent_dim
        for k in range(dim, latent_dim, -1):
            axis = torch.randn(1, k)
            axis = (axis / torch.norm(axis)).repeat(*self.batch_shape, 1, 1)
            self.register_parameter(name='raw_axis_S%d' % k, parameter=torch.nn.Parameter(axis, requires_grad=False))
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
        if torch.is_grad_enabled() and any(a.requires_grad for a in self.axes):
            raise NotImplementedError('gradients with respect to the nested-sphere axes are not provided '
                                      '(fit_gpytorch_manifold is outside the scope of gabotorch_b200)')
        return ops.nested_sphere_project(x, [a.detach().double() for a in self.axes], self.distances_to_axis)

    def forward(self, x1, x2, diag=False, **params):
        _reject_input_grad(x1, x2)
        beta = self._beta_scalar()
        p1 = self._project(x1)
        p2 = p1 if x2 is x1 else self._project(x2)
        if _needs_param_grad(self.raw_beta):
            out = _param_gram(lambda: ops.sphere_gram(p1, p2, kind=_lib.KIND_DIST, diag=diag), beta, 2)
        else:
            out = ops.sphere_gram(p1, p2, float(beta.detach()), _lib.KIND_GAUSS, diag=diag)
        return _finish(out, x1)
