"""Function-level mirror of ``BoManifolds/Riemannian_utils/{sphere,spd}_utils_torch.py``: same names, same argument
meaning, results on the caller's device in float64 -- computed by the fused CUDA kernels (no CPU implementation).

    sphere_distance_torch                      sphere_utils_torch.py:12-55
    affine_invariant_distance_torch            spd_utils_torch.py:53-120
    frobenius_distance_torch                   spd_utils_torch.py:124-156
    logm_torch                                 spd_utils_torch.py:13-30   (also accepts a batch)
    vector_to_symmetric_matrix_mandel_torch    spd_utils_torch.py:159-194
    symmetric_matrix_to_vector_mandel_torch    spd_utils_torch.py:197-226
    max_/min_eigenvalue_constraint_torch       spd_constraints_utils_torch.py:17-50 (plain torch, as in the reference;
                                               recognised and batched in closed form by the constrained solver)
"""
import torch

from . import _lib, ops


def _back(out, like):
    like = torch.as_tensor(like)
    return out if like.is_cuda else out.to(like.device)


def sphere_distance_torch(x1, x2, diag=False):
    return _back(ops.sphere_gram(x1, x2, kind=_lib.KIND_DIST, diag=diag), x1)


def affine_invariant_distance_torch(x1, x2, diagonal_distance=False, compute='f32'):
    if diagonal_distance is True:  # spd_utils_torch.py:72-75
        x2 = torch.as_tensor(x2)
        return torch.zeros(tuple(x2.shape[:-2]) + (1,), dtype=torch.as_tensor(x1).dtype, device=x2.device)
    c = _lib.GABO_F64 if compute == 'f64' else _lib.GABO_F32
    return _back(ops.spd_ai_gram(x1, x2, kind=_lib.KIND_DIST, is_mandel=False, compute=c), x1)


def frobenius_distance_torch(x1, x2, diagonal_distance=False):
    if diagonal_distance is True:  # spd_utils_torch.py:142-145
        x2 = torch.as_tensor(x2)
        return torch.zeros(tuple(x2.shape[:-2]) + (1,), dtype=torch.as_tensor(x1).dtype, device=x2.device)
    return _back(ops.frobenius_gram(x1, x2, kind=_lib.KIND_DIST), x1)


def logm_torch(x):
    return _back(ops.spd_logm(x), x)


def vector_to_symmetric_matrix_mandel_torch(vectors):
    return _back(ops.mandel_unpack(vectors), vectors)


def symmetric_matrix_to_vector_mandel_torch(matrices):
    return _back(ops.mandel_pack(matrices), matrices)


def max_eigenvalue_constraint_torch(x, maximum_eigenvalue):
    """``maximum_eigenvalue - lambda_max(x)``: positive when satisfied (spd_constraints_utils_torch.py:17-32).  One matrix
    or a batch; differentiable (the extreme eigenpair comes from ``gabo_sym_eig``).  The solvers recognise
    ``functools.partial`` objects of this function and evaluate them in closed form (in the kernel where there is one)."""
    from .kernel_utils import _dev64_keep_grad
    from .nested_mappings import _ExtremeEigenvalue
    return maximum_eigenvalue - _ExtremeEigenvalue.apply(_dev64_keep_grad(x), 1)


def min_eigenvalue_constraint_torch(x, minimum_eigenvalue):
    """``lambda_min(x) - minimum_eigenvalue``: positive when satisfied (spd_constraints_utils_torch.py:35-50)."""
    from .kernel_utils import _dev64_keep_grad
    from .nested_mappings import _ExtremeEigenvalue
    return _ExtremeEigenvalue.apply(_dev64_keep_grad(x), -1) - minimum_eigenvalue
