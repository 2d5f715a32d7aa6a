"""Oracle (CPU, fp64) for the SPD part of the hot path.  Test infrastructure only.

Follows ``BoManifolds/Riemannian_utils/spd_utils_torch.py`` of the reference: Mandel conversions
(``:159-226``), affine-invariant distance (``:53-120``), Frobenius distance (``:124-156``), ``logm_torch``
(``:13-30``); the kernel wrappers of ``kernel_utils/kernels_spd.py`` (``:72-100`` Gaussian, ``:160-187``
Laplace, ``:217-241`` Frobenius, ``:268-313`` log-Euclidean); and pymanopt 0.2.x
``manifolds/psd.py`` for the manifold operations (third-party: PARITY UNPINNED, cross-checked against the
reference's numpy formulas ``Riemannian_utils/spd_utils.py:104-213``).
"""
import numpy as np
import torch

SQRT2 = 2.0 ** 0.5
DIST_EPS = 1e-15  # spd_utils_torch.py:120
FROB_EPS = 1e-15  # spd_utils_torch.py:156


# ----------------------------------------------------------------------------------------------
# Mandel notation: v = [diag | sqrt2*1st super-diagonal | sqrt2*2nd super-diagonal | ...]
# ----------------------------------------------------------------------------------------------

def mandel_dim(d_vec):
    """spd_utils_torch.py:175."""
    return int((-1.0 + (1.0 + 8.0 * d_vec) ** 0.5) / 2.0)


def mandel_index(d):
    """(row, col) of every Mandel entry, diagonal by diagonal (spd_utils_torch.py:181-187)."""
    rows, cols = [], []
    for k in range(d):
        for i in range(d - k):
            rows.append(i)
            cols.append(i + k)
    return np.array(rows), np.array(cols)


def vector_to_symmetric_matrix_mandel(vectors):
    """spd_utils_torch.py:159-194, vectorised (no Python loop over matrices)."""
    vectors = torch.as_tensor(vectors)
    d_vec = vectors.shape[-1]
    d = mandel_dim(d_vec)
    r, c = mandel_index(d)
    scale = torch.ones(d_vec, dtype=vectors.dtype)
    scale[d:] = 1.0 / SQRT2
    vals = vectors * scale
    out = torch.zeros(vectors.shape[:-1] + (d, d), dtype=vectors.dtype)
    out[..., r, c] = vals
    out[..., c, r] = vals
    return out


def symmetric_matrix_to_vector_mandel(matrices):
    """spd_utils_torch.py:197-226: off-diagonals are 0.5*(sqrt2*upper + sqrt2*lower)."""
    matrices = torch.as_tensor(matrices)
    d = matrices.shape[-1]
    r, c = mandel_index(d)
    up = matrices[..., r, c]
    lo = matrices[..., c, r]
    v = 0.5 * (SQRT2 * up + SQRT2 * lo)
    v[..., :d] = up[..., :d]
    return v


# ----------------------------------------------------------------------------------------------
# Distances
# ----------------------------------------------------------------------------------------------

def whitened(x1, x2):
    """W_ij = L_i^-1 X2_j L_i^-T with L_i = chol(X1_i) -- spd_utils_torch.py:87-103."""
    x1 = torch.as_tensor(x1)
    x2 = torch.as_tensor(x2)
    chol = torch.linalg.cholesky(x1)
    cinv = torch.inverse(chol)
    a = cinv.unsqueeze(-3)        # (..., N1, 1, d, d)
    b = x2.unsqueeze(-4)          # (..., 1, N2, d, d)
    return torch.matmul(torch.matmul(a, b), a.transpose(-2, -1))


def affine_invariant_distance(x1, x2, diagonal_distance=False, exact=False):
    """spd_utils_torch.py:53-120.

    ``exact=False`` is the faithful restatement: eigenvalues of the fp64 whitened matrix are rounded to
    float32 (the reference stores them in a default-dtype ``torch.zeros`` buffer, ``:108``), log, square,
    sum and sqrt(+1e-15) then run in float32 and only the result is widened (``:117-120``).
    ``exact=True`` keeps fp64 throughout (used to quantify the reference's own noise floor).
    """
    x1 = torch.as_tensor(x1)
    x2 = torch.as_tensor(x2)
    if diagonal_distance is True:
        return torch.zeros(tuple(x2.shape[:-2]) + (1,), dtype=x1.dtype)  # :72-75
    w = whitened(x1, x2)
    eig = torch.linalg.eigh(w, UPLO='U').eigenvalues
    if not exact:
        eig = eig.to(torch.float32)
    logeig = torch.log(eig)
    return torch.sqrt(torch.sum(logeig * logeig, dim=-1) + DIST_EPS).double()


def affine_invariant_distance_loop(x1, x2):
    """The reference's own structure: materialised (N1,N2,d,d) operands, two ``bmm`` and one
    single-matrix symmetric eigen-solve per pair inside a Python loop (spd_utils_torch.py:92-110).
    Used only as the timed CPU baseline ("port"); results equal ``affine_invariant_distance``."""
    x1 = torch.as_tensor(x1)
    x2 = torch.as_tensor(x2)
    dim = x1.shape[-1]
    a = x1.unsqueeze(-3)
    b = x2.unsqueeze(-4)
    chol = torch.linalg.cholesky(a)
    cinv = torch.inverse(chol)
    cinv = torch.cat(b.shape[-3] * [cinv], dim=-3)
    b = torch.cat(a.shape[-4] * [b], dim=-4)
    w = torch.bmm(torch.bmm(cinv.reshape(-1, dim, dim), b.reshape(-1, dim, dim)),
                  cinv.reshape(-1, dim, dim).transpose(-2, -1))
    eig_values = torch.zeros(w.shape[0], dim)  # float32, as in the reference
    for i in range(w.shape[0]):
        eig_values[i] = torch.linalg.eigh(w[i], UPLO='U').eigenvalues
    eigv = eig_values.view(tuple(b.shape[:-2]) + (dim,))
    logeigv = torch.log(eigv)
    return torch.sqrt(torch.sum(logeigv * logeigv, dim=-1) + DIST_EPS).double()


def frobenius_distance(x1, x2, diagonal_distance=False):
    """|| X1_i - X2_j + 1e-15 ||_F -- spd_utils_torch.py:124-156 (the eps is added to every entry)."""
    x1 = torch.as_tensor(x1)
    x2 = torch.as_tensor(x2)
    if diagonal_distance is True:
        return torch.zeros(tuple(x2.shape[:-2]) + (1,), dtype=x1.dtype)
    diff = x1.unsqueeze(-3) - x2.unsqueeze(-4) + FROB_EPS
    return torch.sqrt(torch.sum(diff * diff, dim=(-2, -1))).double()


def logm(x):
    """V diag(log lambda) V^-1 -- spd_utils_torch.py:13-30, batched."""
    x = torch.as_tensor(x)
    lam, vec = torch.linalg.eigh(x, UPLO='U')
    return torch.matmul(vec * torch.log(lam).unsqueeze(-2), torch.inverse(vec))


def sqrtm(x):
    """spd_utils_torch.py:33-50, batched."""
    x = torch.as_tensor(x)
    lam, vec = torch.linalg.eigh(x, UPLO='U')
    return torch.matmul(vec * torch.sqrt(lam).unsqueeze(-2), torch.inverse(vec))


# ----------------------------------------------------------------------------------------------
# Kernels on Mandel vectors (kernels_spd.py)
# ----------------------------------------------------------------------------------------------

def _beta64(beta):
    return torch.as_tensor(beta, dtype=torch.float64)


def spd_affine_invariant_gaussian_kernel(x1, x2, beta, diagonal_distance=False, loop=False, exact=False):
    """kernels_spd.py:90-100."""
    m1 = vector_to_symmetric_matrix_mandel(x1)
    m2 = vector_to_symmetric_matrix_mandel(x2)
    if loop and not diagonal_distance:
        d = affine_invariant_distance_loop(m1, m2)
    else:
        d = affine_invariant_distance(m1, m2, diagonal_distance=diagonal_distance, exact=exact)
    return torch.exp(-torch.mul(d, d).mul(_beta64(beta)))


def spd_affine_invariant_laplace_kernel(x1, x2, beta, diagonal_distance=False):
    """kernels_spd.py:178-187."""
    m1 = vector_to_symmetric_matrix_mandel(x1)
    m2 = vector_to_symmetric_matrix_mandel(x2)
    d = affine_invariant_distance(m1, m2, diagonal_distance=diagonal_distance)
    return torch.exp(-d.mul(_beta64(beta)))


def spd_frobenius_gaussian_kernel(x1, x2, lengthscale, diagonal_distance=False):
    """kernels_spd.py:230-241."""
    m1 = vector_to_symmetric_matrix_mandel(x1)
    m2 = vector_to_symmetric_matrix_mandel(x2)
    d = frobenius_distance(m1, m2, diagonal_distance=diagonal_distance)
    ls = _beta64(lengthscale)
    return torch.exp(-torch.mul(d, d).div(ls * ls))


def spd_log_euclidean_gaussian_kernel(x1, x2, lengthscale, diagonal_distance=False):
    """kernels_spd.py:283-313."""
    m1 = logm(vector_to_symmetric_matrix_mandel(x1))
    m2 = logm(vector_to_symmetric_matrix_mandel(x2))
    d = frobenius_distance(m1, m2, diagonal_distance=diagonal_distance)
    ls = _beta64(lengthscale)
    return torch.exp(-torch.mul(d, d).div(ls * ls))


# ----------------------------------------------------------------------------------------------
# Manifold operations: pymanopt 0.2.x PositiveDefinite (third-party, restated). numpy fp64, batched
# on leading axes: x (..., d, d) SPD, u (..., d, d) symmetric.
# ----------------------------------------------------------------------------------------------

def _sym(a):
    return 0.5 * (a + np.swapaxes(a, -1, -2))


def _chol_inv(x):
    c = np.linalg.cholesky(x)
    return c, np.linalg.inv(c)


def _eigfun(w, f):
    lam, v = np.linalg.eigh(w)
    return (v * f(lam)[..., None, :]) @ np.swapaxes(v, -1, -2)


def inner(x, u, v):
    """tr(X^-1 U X^-1 V)."""
    a = np.linalg.solve(x, u)
    b = np.linalg.solve(x, v)
    return np.sum(a * np.swapaxes(b, -1, -2), axis=(-2, -1))


def norm(x, u):
    """|| L^-1 U L^-T ||_F."""
    _, ci = _chol_inv(x)
    w = ci @ u @ np.swapaxes(ci, -1, -2)
    return np.sqrt(np.sum(w * w, axis=(-2, -1)))


def dist(x, y):
    """|| logm(L^-1 Y L^-T) ||_F ; same number as spd_utils.py:196-197."""
    _, ci = _chol_inv(x)
    w = _sym(ci @ y @ np.swapaxes(ci, -1, -2))
    lam = np.linalg.eigvalsh(w)
    return np.sqrt(np.sum(np.log(lam) ** 2, axis=-1))


def egrad2rgrad(x, g):
    """X sym(G) X."""
    return x @ _sym(g) @ x


def exp(x, u):
    """X expm(X^-1 U), evaluated through the Cholesky whitening so the result is symmetric:
    L expm(L^-1 U L^-T) L^T  (identical in exact arithmetic to spd_utils.py:117-118)."""
    c, ci = _chol_inv(x)
    w = _sym(ci @ u @ np.swapaxes(ci, -1, -2))
    return c @ _eigfun(w, np.exp) @ np.swapaxes(c, -1, -2)


retr = exp


def log(x, y):
    """L logm(L^-1 Y L^-T) L^T  (spd_utils.py:136-137 in exact arithmetic)."""
    c, ci = _chol_inv(x)
    w = _sym(ci @ y @ np.swapaxes(ci, -1, -2))
    return c @ _eigfun(w, np.log) @ np.swapaxes(c, -1, -2)


def transp(x, y, u):
    """Identity transport (pymanopt 0.2.x PositiveDefinite.transp)."""
    return u


def parallel_transport(x1, x2, v):
    """True AI parallel transport E V E^T with E = (X2 X1^-1)^1/2 (spd_utils.py:200-213), computed
    as E = L (L^-1 X2 L^-T)^1/2 L^-1."""
    c, ci = _chol_inv(x1)
    w = _sym(ci @ x2 @ np.swapaxes(ci, -1, -2))
    e = c @ _eigfun(w, np.sqrt) @ ci
    return e @ v @ np.swapaxes(e, -1, -2)


def spd_sample(rng, n, d, min_eig=0.001, max_eig=5.0, max_cond=None):
    """Law of the reference's ``spd_sample`` (spd_utils.py:290-306): eigenvalues U[min,max], Q from QR of
    randn(d,d); optional rejection of cond > max_cond (spd_gaussian_kernel_parameters.py:91-96)."""
    out = np.empty((n, d, d))
    k = 0
    while k < n:
        lam = min_eig + (max_eig - min_eig) * rng.random(d)
        q, _ = np.linalg.qr(rng.standard_normal((d, d)))
        if max_cond is not None and lam.max() / lam.min() > max_cond:
            continue
        out[k] = (q * lam) @ q.T
        k += 1
    return _sym(out)


def ackley(x_mandel):
    """Ackley function on SPD(d) with base point 2I: test_functions_spd.py:34-69. (N, dv) -> (N,)."""
    xm = vector_to_symmetric_matrix_mandel(torch.as_tensor(x_mandel, dtype=torch.float64)).numpy()
    xm = xm.reshape((-1,) + xm.shape[-2:])
    d = xm.shape[-1]
    dv = d + d * (d - 1) // 2
    base = np.broadcast_to(2.0 * np.eye(d), xm.shape)
    xp = log(base, xm)
    v = symmetric_matrix_to_vector_mandel(torch.from_numpy(xp)).numpy()
    v[:, d:] /= SQRT2
    a, b, c = 20., 0.2, 2. * np.pi
    t1 = -a * np.exp(-b * np.sqrt(np.sum(v ** 2, axis=-1) / dv))
    t2 = -np.exp(np.sum(np.cos(c * v) / dv, axis=-1))
    return t1 + t2 + a + np.exp(1.)
