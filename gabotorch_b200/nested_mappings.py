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
