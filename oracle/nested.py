"""Oracle (CPU, fp64) for the nested SPD projection of HD-GaBO.  Test infrastructure only.

Follows ``BoManifolds/nested_mappings/nested_spd_utils.py:13-48`` (``Y = W^T X W`` through two ``bmm``)
and ``kernel_utils/kernels_nested_spd.py:104-136`` (nested affine-invariant Gaussian kernel); the approximate inverse
``projection_from_nested_spd_to_spd`` follows ``nested_spd_utils.py:51-118`` with ``sqrtm_torch``
(``Riemannian_utils/spd_utils_torch.py:33-50``).
"""
import numpy as np
import torch

from . import spd as _spd


def projection_from_spd_to_nested_spd(x_spd, projection_matrix):
    """nested_spd_utils.py:31-48."""
    x_spd = torch.as_tensor(x_spd)
    w = torch.as_tensor(projection_matrix).to(x_spd.dtype)
    return torch.matmul(torch.matmul(w.transpose(-2, -1), x_spd), w)


def mandel_projection_matrix(w):
    """The (dv_low x dv_high) matrix P with  mandel(W^T X W) = P @ mandel(X)  (our derivation).

    For X = sum_{p<=q} x_pq E_pq (E_pq the symmetric unit matrices) and Mandel weights m_ii = 1,
    m_{i!=j} = sqrt(2):  P[(a,b),(p,q)] = m_ab (W_pa W_qb + [p!=q] W_qa W_pb) / m_pq.
    """
    w = np.asarray(w, dtype=np.float64)
    D, d = w.shape
    rh, ch = _spd.mandel_index(D)
    rl, cl = _spd.mandel_index(d)
    P = np.zeros((len(rl), len(rh)))
    for o, (a, b) in enumerate(zip(rl, cl)):
        mab = 1.0 if a == b else _spd.SQRT2
        for i, (p, q) in enumerate(zip(rh, ch)):
            if p == q:
                P[o, i] = mab * w[p, a] * w[p, b]
            else:
                P[o, i] = mab * (w[p, a] * w[q, b] + w[q, a] * w[p, b]) / _spd.SQRT2
    return P


def projection_mandel(x_mandel, w):
    """Mandel-to-Mandel form of the projection: the composition G3 -> P1 -> G3 of SURVEY section 8."""
    x = _spd.vector_to_symmetric_matrix_mandel(torch.as_tensor(x_mandel, dtype=torch.float64))
    y = projection_from_spd_to_nested_spd(x, torch.as_tensor(w, dtype=torch.float64))
    return _spd.symmetric_matrix_to_vector_mandel(y)


def nested_spd_affine_invariant_gaussian_kernel(x1, x2, w, beta):
    """kernels_nested_spd.py:122-136: Mandel unpack, project both inputs, AI distance, exp(-beta d^2)."""
    m1 = projection_from_spd_to_nested_spd(_spd.vector_to_symmetric_matrix_mandel(x1), w)
    m2 = projection_from_spd_to_nested_spd(_spd.vector_to_symmetric_matrix_mandel(x2), w)
    d = _spd.affine_invariant_distance(m1, m2)
    return torch.exp(-torch.mul(d, d).mul(torch.as_tensor(beta, dtype=torch.float64)))


def nested_spd_log_euclidean_gaussian_kernel(x1, x2, w, lengthscale):
    """kernels_nested_spd.py:209-246: Mandel unpack, project, logm per matrix, Frobenius distance, exp(-d^2 / l^2)."""
    m1 = _spd.logm(projection_from_spd_to_nested_spd(_spd.vector_to_symmetric_matrix_mandel(x1), w))
    m2 = _spd.logm(projection_from_spd_to_nested_spd(_spd.vector_to_symmetric_matrix_mandel(x2), w))
    d = _spd.frobenius_distance(m1, m2)
    ls = torch.as_tensor(lengthscale, dtype=torch.float64)
    return torch.exp(-torch.mul(d, d).div(ls * ls))


def grassmann_rand(rng, D, d):
    """pymanopt Grassmann.rand: Q of QR(randn(D,d))."""
    q, _ = np.linalg.qr(rng.standard_normal((D, d)))
    return q


def projection_from_nested_spd_to_spd(y, w, v, c, k):
    """nested_spd_utils.py:77-118: X = R [Y B; B^T C] R^T, R = [W V], B = sqrtm(Y) K sqrtm(C)."""
    y = torch.as_tensor(y, dtype=torch.float64)
    w, v, c, k = (torch.as_tensor(t, dtype=torch.float64) for t in (w, v, c, k))
    single = y.dim() == 2
    if single:
        y = y[None]
    rot = torch.cat((w, v), dim=1)
    sqrt_c = _spd.sqrtm(c)
    out = []
    for n in range(y.shape[0]):
        side = torch.mm(torch.mm(_spd.sqrtm(y[n]), k), sqrt_c)
        xr = torch.cat((torch.cat((y[n], side), dim=1), torch.cat((side.T, c), dim=1)), dim=0)
        out.append(torch.mm(rot, torch.mm(xr, rot.T)))
    out = torch.stack(out)
    return out[0] if single else out


def min_affine_invariant_distance_reconstruction_cost(x_data, y, w, v, c, k):
    """nested_spd_optimization.py:22-55: sum over the data of d_AI(X_n, Xrec_n)^2, one pair at a time, accumulated in a
    float32 tensor like the reference (``cost = torch.zeros(n_data)``, :49)."""
    x_data = torch.as_tensor(x_data, dtype=torch.float64)
    xr = projection_from_nested_spd_to_spd(y, w, v, c, k)
    cost = torch.zeros(x_data.shape[0])
    for n in range(x_data.shape[0]):
        cost[n] = _spd.affine_invariant_distance(x_data[n].unsqueeze(0), xr[n].unsqueeze(0))
    return torch.sum(cost * cost)


def min_log_euclidean_distance_reconstruction_cost(x_data, y, w, v, c, k):
    """nested_spd_optimization.py:58-92: the same with ||logm X_n - logm Xrec_n + 1e-15||_F (spd_utils_torch.py:156)."""
    x_data = torch.as_tensor(x_data, dtype=torch.float64)
    xr = projection_from_nested_spd_to_spd(y, w, v, c, k)
    cost = torch.zeros(x_data.shape[0])
    for n in range(x_data.shape[0]):
        cost[n] = _spd.frobenius_distance(_spd.logm(x_data[n]).unsqueeze(0), _spd.logm(xr[n]).unsqueeze(0))
    return torch.sum(cost * cost)
