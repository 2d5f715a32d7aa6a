"""Mirror of ``BoManifolds/nested_mappings/nested_spd_utils.py:13-48``: the HD-GaBO projection SPD(D) -> SPD(d).

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


def projection_from_spd_to_nested_spd(x_spd, projection_matrix):
    """Y = W^T X W for (..., D, D) matrices -> (..., d, d), dtype / device of ``x_spd`` (nested_spd_utils.py:13-48)."""
    x = torch.as_tensor(x_spd)
    proj = NestedSpdProjection(projection_matrix)
    single = x.dim() == 2
    if single:
        x = x[None]
    xm = ops.mandel_pack(x)                      # fp64 Mandel vectors on the device
    ym = proj.mandel(xm.to(torch.float32))       # tensor-core contraction, fp32
    y = ops.mandel_unpack(ym.to(torch.float64))
    if single:
        y = y[0]
    y = y.to(x.dtype)
    return y if x.is_cuda else y.to(x.device)
