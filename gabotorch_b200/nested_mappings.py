"""Mirror of ``BoManifolds/nested_mappings``: ``nested_spd_utils.py`` (projection SPD(D) -> SPD(d), :13-48, and its
approximate inverse, :51-118) and ``nested_spheres_utils.py`` (projection chain S^D -> S^d and back, :13-213).

``projection_from_spd_to_nested_spd(x_spd, projection_matrix)`` keeps the reference's signature (matrices in, matrices
out).  The work is one tensor-core contraction in Mandel coordinates (``gabo_nested_spd_project``); callers that already
hold Mandel vectors (the GP inputs of the reference are Mandel vectors, kernels_nested_spd.py:122-127) should use
``projection_mandel`` and skip the pack / unpack either side.
"""
import torch

from . import ops


class NestedSpdProjection:
    """Pre-packed projection operator for a fixed W (D x d): build once, apply to many batches."""

    def __init__(self, projection_matrix):
        w = ops.to_dev64(projection_matrix)
        if w.dim() != 2 or w.shape[1] > w.shape[0]:
            raise ValueError('projection_matrix must be (D, d) with d <= D')
        self.D, self.d = int(w.shape[0]), int(w.shape[1])
        self.pack = ops.nested_projection_matrix(w)

    def mandel(self, x_mandel):
        """(..., D(D+1)/2) -> (..., d(d+1)/2), float32 on the device."""
        x = torch.as_tensor(x_mandel)
        flat = x.reshape(-1, x.shape[-1])
        y = ops.nested_spd_project(flat, self.D, self.d, self.pack)
        return y.reshape(tuple(x.shape[:-1]) + (y.shape[-1],))


def projection_mandel(x_mandel, projection_matrix):
    return NestedSpdProjection(projection_matrix).mandel(x_mandel)


def projection_from_spd_to_nested_spd(x_spd, projection_matrix, tensor_cores=None):
    """Y = W^T X W for (..., D, D) matrices -> (..., d, d), dtype / device of ``x_spd`` (nested_spd_utils.py:13-48).

    float64 inputs (the reference's dtype) take the fp64 contraction ``gabo_nested_spd_project_f64`` -- the reference
    computes an exact fp64 ``bmm`` (:44), and the nested kernel classes use the same entry, so the function and the
    kernels agree to rounding.  float32 inputs (or ``tensor_cores=True``) take the 3xTF32 tensor-core contraction
    ``gabo_nested_spd_project`` (about 1e-6 relative): that is the path for screening millions of raw samples."""
    x = torch.as_tensor(x_spd)
    single = x.dim() == 2
    if single:
        x = x[None]
    if tensor_cores is None:
        tensor_cores = x.dtype == torch.float32
    xm = ops.mandel_pack(x)                      # fp64 Mandel vectors on the device
    if tensor_cores:
        ym = NestedSpdProjection(projection_matrix).mandel(xm.to(torch.float32)).to(torch.float64)
    else:
        ym = ops.nested_spd_project_f64(xm, projection_matrix)
    y = ops.mandel_unpack(ym)
    if single:
        y = y[0]
    y = y.to(x.dtype)
    return y if x.is_cuda else y.to(x.device)


def _like(out, ref):
    """Result in the dtype / device of the caller's tensor (the reference is dtype-preserving)."""
    ref = torch.as_tensor(ref)
    out = out.to(ref.dtype)
    return out if ref.is_cuda else out.to(ref.device)


class NestedSpdReconstruction:
    """Point-independent part of ``projection_from_nested_spd_to_spd`` for fixed (W, V, C, K): build once, apply to
    many batches."""

    def __init__(self, projection_matrix, projection_complement_matrix, bottom_spd_matrix, contraction_matrix):
        self.D = int(projection_matrix.shape[0])
        self.d = int(projection_matrix.shape[1])
        self.pack = ops.nested_spd_reconstruct_pack(projection_matrix, projection_complement_matrix,
                                                    bottom_spd_matrix, contraction_matrix)

    def __call__(self, x_spd_low_dimension):
        return ops.nested_spd_reconstruct(x_spd_low_dimension, self.D, self.pack)


def projection_from_nested_spd_to_spd(x_spd_low_dimension, projection_matrix, projection_complement_matrix,
                                      bottom_spd_matrix, contraction_matrix):
    """X = R [Y B; B^T C] R^T with R = [W V], B = Y^(1/2) K C^(1/2) (nested_spd_utils.py:51-118):
    (d, d) or (N, d, d) -> (D, D) or (N, D, D)."""
    y = torch.as_tensor(x_spd_low_dimension)
    rec = NestedSpdReconstruction(projection_matrix, projection_complement_matrix, bottom_spd_matrix,
                                  contraction_matrix)
    single = y.dim() == 2
    x = rec(y[None] if single else y)
    return _like(x[0] if single else x, y)


def _as_list(v):
    return list(v) if isinstance(v, (list, tuple)) else [v]


def projection_from_sphere_to_nested_sphere(x, sphere_axis, sphere_distance_to_axis):
    """nested_spheres_utils.py:13-67."""
    return _like(ops.nested_sphere_to_nested(x, sphere_axis, sphere_distance_to_axis), x)


def projection_from_sphere_to_next_subsphere(x, sphere_axis, sphere_distance_to_axis):
    """nested_spheres_utils.py:68-116: (..., k) on S^{k-1} -> (..., k-1) on S^{k-2}."""
    return _like(ops.nested_sphere_project(x, [sphere_axis], [sphere_distance_to_axis]), x)


def projection_from_sphere_to_subsphere(x, sphere_axes, sphere_distances_to_axes):
    """nested_spheres_utils.py:117-147: list [x, S^{D-2} points, ..., S^{d-1} points]."""
    levels = ops.nested_sphere_chain(x, _as_list(sphere_axes), _as_list(sphere_distances_to_axes))
    return [x] + [_like(v, x) for v in levels]


def projection_from_subsphere_to_next_sphere(x_subsphere, sphere_axis, sphere_distance_to_axis):
    """nested_spheres_utils.py:149-180: (N, k-1) -> (N, k)."""
    return _like(ops.nested_sphere_reconstruct(x_subsphere, [sphere_axis], [sphere_distance_to_axis])[-1], x_subsphere)


def projection_from_subsphere_to_sphere(x_subsphere, sphere_axes, sphere_distances_to_axes):
    """nested_spheres_utils.py:182-213: list [x_subsphere, next sphere, ..., S^{D-1} points]; the axes are given in
    the order of the projection and consumed last to first, as in the reference."""
    levels = ops.nested_sphere_reconstruct(x_subsphere, _as_list(sphere_axes), _as_list(sphere_distances_to_axes))
    return [x_subsphere] + [_like(v, x_subsphere) for v in levels]


# ----------------------------------------------------------------------------------------------------------------
# nested_spd_constraints_utils.py of the reference: eigenvalue constraints of the AMBIENT matrix for a latent optimiser
# ----------------------------------------------------------------------------------------------------------------

class _ExtremeEigenvalue(torch.autograd.Function):
    """Largest (``sign`` > 0) or smallest eigenvalue of a batch of symmetric matrices from ONE ``gabo_sym_eig`` launch
    (D <= 32); backward ``g u u^T`` with the eigenvector u (what autograd gives through ``torch.symeig``)."""

    @staticmethod
    def forward(ctx, mat, sign):
        lam, vec, _ = ops.sym_eig(mat)
        i = lam.argmax(-1) if sign > 0 else lam.argmin(-1)
        u = vec.gather(-1, i[..., None, None].expand(vec.shape[:-1] + (1,))).squeeze(-1)
        ctx.save_for_backward(u)
        return lam.gather(-1, i.unsqueeze(-1)).squeeze(-1)

    @staticmethod
    def backward(ctx, g):
        u, = ctx.saved_tensors
        return g[..., None, None] * u.unsqueeze(-1) * u.unsqueeze(-2), None


def _nested_to_ambient_autograd(x_nested_spd, projection_matrix, projection_complement_matrix, bottom_spd_matrix,
                                contraction_matrix):
    """projection_from_nested_spd_to_spd (nested_spd_utils.py:51-118) as differentiable device code, (..., d, d) ->
    (..., D, D): the square roots are ``_SpectralFn`` (eigenpairs from ``gabo_sym_eig``, Daleckii-Krein backward)."""
    from .kernel_utils import _dev64_keep_grad
    from .nested_optimization import _SpectralFn, _reconstruct_spd
    y = _dev64_keep_grad(x_nested_spd)
    w, v, c, k = (_dev64_keep_grad(t).to(y.device) for t in (projection_matrix, projection_complement_matrix,
                                                             bottom_spd_matrix, contraction_matrix))
    single = y.dim() == 2
    yb = y[None] if single else y.reshape((-1,) + tuple(y.shape[-2:]))
    x = _reconstruct_spd(yb, _SpectralFn.apply(yb, 1), w, v, c, k)
    return x[0] if single else x.reshape(tuple(y.shape[:-2]) + tuple(x.shape[-2:]))


def max_eigenvalue_nested_spd_constraint(x_nested_spd, maximum_eigenvalue, projection_matrix,
                                         projection_complement_matrix, bottom_spd_matrix, contraction_matrix):
    """``maximum_eigenvalue - lambda_max`` of the ambient matrix reconstructed from the nested SPD matrix
    (nested_spd_constraints_utils.py:13-44; satisfied when >= 0).  Differentiable; accepts one matrix or a batch."""
    x = _nested_to_ambient_autograd(x_nested_spd, projection_matrix, projection_complement_matrix, bottom_spd_matrix,
                                    contraction_matrix)
    return maximum_eigenvalue - _ExtremeEigenvalue.apply(x, 1)


def min_eigenvalue_nested_spd_constraint(x_nested_spd, minimum_eigenvalue, projection_matrix,
                                         projection_complement_matrix, bottom_spd_matrix, contraction_matrix):
    """``lambda_min - minimum_eigenvalue`` of the reconstructed ambient matrix (nested_spd_constraints_utils.py:47-78)."""
    x = _nested_to_ambient_autograd(x_nested_spd, projection_matrix, projection_complement_matrix, bottom_spd_matrix,
                                    contraction_matrix)
    return _ExtremeEigenvalue.apply(x, -1) - minimum_eigenvalue


max_eigenvalue_nested_spd_constraint.supports_batch = True
min_eigenvalue_nested_spd_constraint.supports_batch = True


def random_nested_spd_with_spd_eigenvalue_constraints(self, random_spd_fct, projection_matrix):
    """A nested SPD sample: a sample of the ambient manifold that respects the constraints there, projected
    (nested_spd_constraints_utils.py:81-97).  Bound as the ``rand`` method of the latent manifold in hd_gabo_spd.py:236-239;
    returns numpy like the reference (pymanopt's format)."""
    x_spd = torch.as_tensor(random_spd_fct(), dtype=torch.as_tensor(projection_matrix).dtype)
    return projection_from_spd_to_nested_spd(x_spd, projection_matrix).cpu().numpy()
