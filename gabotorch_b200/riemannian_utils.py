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


# ----------------------------------------------------------------------------------------------------------------
# host helpers the reference's examples import from Riemannian_utils/spd_utils.py (numpy, one matrix at a time: data
# generation and the `rand` method bound to the manifold object, gabo_spd.py:100-124).  Not device work in the reference
# either; the batched device equivalents are PositiveDefinite.rand_batch and the *_mandel_torch functions above.
# ----------------------------------------------------------------------------------------------------------------

def symmetric_matrix_to_vector_mandel(M):
    """Mandel vector of ONE symmetric matrix: diagonal, then the k-th super-diagonals times sqrt 2 (spd_utils.py:57-76)."""
    import numpy as np
    M = np.asarray(M)
    return np.concatenate([M.diagonal()] + [2.0 ** 0.5 * M.diagonal(i) for i in range(1, M.shape[0])])


def vector_to_symmetric_matrix_mandel(v):
    """Inverse of ``symmetric_matrix_to_vector_mandel`` (spd_utils.py:79-101)."""
    import numpy as np
    v = np.asarray(v)
    n = int((-1.0 + (1.0 + 8.0 * v.shape[0]) ** 0.5) / 2.0)
    M = np.diag(v[:n]).astype(v.dtype, copy=True)
    start = n
    for i in range(1, n):
        off = v[start:start + n - i] / 2.0 ** 0.5
        M += np.diag(off, i) + np.diag(off, -i)
        start += n - i
    return M


def spd_sample(self):
    """A random SPD matrix with eigenvalues uniform in [self.min_eig, self.max_eig] and a Haar-like orthogonal factor
    (spd_utils.py:290-306); bound as the ``rand`` method of a ``PositiveDefinite`` object by the examples."""
    import numpy as np
    d = self.min_eig + (self.max_eig - self.min_eig) * np.random.rand(self._n)
    u, _ = np.linalg.qr(np.random.randn(self._n, self._n))
    return u @ np.diag(d) @ u.T
