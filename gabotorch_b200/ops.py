"""Tensor-level wrappers of the C ABI (``include/gabo_b200.h``): torch CUDA tensors in, torch CUDA tensors out.

Every function here is a thin argument-marshalling layer: it checks shapes, makes the operands contiguous fp64 device
tensors (the reference's dtype), allocates the output and enqueues the library call on the current torch stream.
There is no arithmetic here and no CPU fallback: a CPU tensor is copied to the device, computed there, and the result
is returned on the device (the reference-facing classes in ``kernels.py`` / ``manifold_optimize.py`` move results
back to the caller's device).
"""
import ctypes
import math

import torch

from . import _lib

_DT = {torch.float32: _lib.GABO_F32, torch.float64: _lib.GABO_F64}


def device():
    """The CUDA device the library computes on (the current one).  Raises when no GPU is visible."""
    if not torch.cuda.is_available():
        raise _lib.GaboError('gabotorch_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def _p(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else None)


def to_dev64(x):
    """Contiguous fp64 tensor on the compute device (no copy when it already is one)."""
    x = torch.as_tensor(x)
    dev = x.device if x.is_cuda else device()
    return x.detach().to(device=dev, dtype=torch.float64).contiguous()


def _flat_batches(x, trailing):
    """View (..., *trailing-dims) as (B, *trailing-dims)."""
    lead = x.shape[:x.dim() - trailing]
    nb = 1
    for s in lead:
        nb *= int(s)
    return x.reshape((nb,) + tuple(x.shape[x.dim() - trailing:])), tuple(lead)


def _broadcast_lead(a, b, trailing):
    la, lb = a.shape[:a.dim() - trailing], b.shape[:b.dim() - trailing]
    lead = torch.broadcast_shapes(la, lb)
    if tuple(la) != tuple(lead):
        a = a.expand(tuple(lead) + tuple(a.shape[a.dim() - trailing:])).contiguous()
    if tuple(lb) != tuple(lead):
        b = b.expand(tuple(lead) + tuple(b.shape[b.dim() - trailing:])).contiguous()
    return a, b, tuple(lead)


# ----------------------------------------------------------------------------------------------------------------
# sphere Gram (G1 + G2)
# ----------------------------------------------------------------------------------------------------------------

def sphere_gram(x1, x2, param=0.0, kind=_lib.KIND_GAUSS, diag=False, out_dtype=torch.float64):
    """f(d(x1_i, x2_j)) for points (..., N, D) on the sphere.  Returns (..., N1, N2), or (..., N, 1) when diag."""
    lib = _lib.load()
    x1, x2 = to_dev64(x1), to_dev64(x2)
    if x1.dim() < 2 or x2.dim() < 2 or x1.shape[-1] != x2.shape[-1]:
        raise ValueError('sphere_gram: expected (..., N, D) inputs with equal D, got %s and %s'
                         % (tuple(x1.shape), tuple(x2.shape)))
    x1, x2, lead = _broadcast_lead(x1, x2, 2)
    dim = x1.shape[-1]
    b1, _ = _flat_batches(x1, 2)
    b2, _ = _flat_batches(x2, 2)
    n1, n2 = b1.shape[1], b2.shape[1]
    s = _lib.stream_ptr()
    if diag:
        if n1 != n2:
            raise ValueError('sphere_gram(diag=True) needs the same number of points in x1 and x2')
        out = torch.empty(b1.shape[0], n1, 1, dtype=out_dtype, device=x1.device)
        for b in range(b1.shape[0]):
            _lib.check(lib.gabo_sphere_gram_diag(_p(b1[b]), _p(b2[b]), n1, dim, float(param), kind, _p(out[b]),
                                                 _DT[out_dtype], s), 'gabo_sphere_gram_diag')
        return out.reshape(lead + (n1, 1))
    out = torch.empty(b1.shape[0], n1, n2, dtype=out_dtype, device=x1.device)
    for b in range(b1.shape[0]):
        _lib.check(lib.gabo_sphere_gram(_p(b1[b]), n1, _p(b2[b]), n2, dim, float(param), kind, _p(out[b]),
                                        _DT[out_dtype], n2, s), 'gabo_sphere_gram')
    return out.reshape(lead + (n1, n2))


# ----------------------------------------------------------------------------------------------------------------
# Mandel notation (G3)
# ----------------------------------------------------------------------------------------------------------------

def mandel_dim(d_vec):
    """Matrix size d for a Mandel vector of length d(d+1)/2 (spd_utils_torch.py:175)."""
    d = int((-1.0 + (1.0 + 8.0 * d_vec) ** 0.5) / 2.0)
    if d * (d + 1) // 2 != d_vec:
        raise ValueError('%d is not a Mandel vector length d(d+1)/2' % d_vec)
    return d


def mandel_unpack(vec):
    """(..., d(d+1)/2) -> (..., d, d); vector_to_symmetric_matrix_mandel_torch (spd_utils_torch.py:159-194)."""
    lib = _lib.load()
    v = to_dev64(vec)
    d = mandel_dim(v.shape[-1])
    flat = v.reshape(-1, v.shape[-1])
    out = torch.empty(flat.shape[0], d, d, dtype=torch.float64, device=v.device)
    _lib.check(lib.gabo_mandel_unpack(_p(flat), flat.shape[0], d, _p(out), _lib.stream_ptr()), 'gabo_mandel_unpack')
    return out.reshape(tuple(v.shape[:-1]) + (d, d))


def mandel_pack(mat):
    """(..., d, d) -> (..., d(d+1)/2); symmetric_matrix_to_vector_mandel_torch (spd_utils_torch.py:197-226)."""
    lib = _lib.load()
    m = to_dev64(mat)
    d = m.shape[-1]
    if m.dim() < 2 or m.shape[-2] != d:
        raise ValueError('mandel_pack: expected (..., d, d), got %s' % (tuple(m.shape),))
    flat = m.reshape(-1, d, d)
    out = torch.empty(flat.shape[0], d * (d + 1) // 2, dtype=torch.float64, device=m.device)
    _lib.check(lib.gabo_mandel_pack(_p(flat), flat.shape[0], d, _p(out), _lib.stream_ptr()), 'gabo_mandel_pack')
    return out.reshape(tuple(m.shape[:-2]) + (d * (d + 1) // 2,))


# ----------------------------------------------------------------------------------------------------------------
# SPD affine-invariant Gram (G4 + G5)
# ----------------------------------------------------------------------------------------------------------------

class NotPositiveDefiniteError(_lib.GaboError):
    """An input matrix has no Cholesky factor (the reference raises inside torch.cholesky, spd_utils_torch.py:87)."""


def spd_factor(x, d, is_mandel, check=True, flags=None):
    """Per-point factor records [L | L^-1] for n points given as (n, dv) Mandel vectors or (n, d, d) matrices.
    ``flags`` (one int32 on the device) lets a caller accumulate the not-positive-definite bit over several calls and
    test it once (``check_spd_flags``) instead of synchronising per call."""
    lib = _lib.load()
    x = to_dev64(x)
    n = x.shape[0]
    fs = lib.gabo_spd_factor_stride(d)
    if fs < 0:
        raise ValueError('SPD(%d): matrix size outside [1, %d]' % (d, _lib.MAX_SPD_DIM))
    fac = torch.empty(n, fs, dtype=torch.float64, device=x.device)
    own = flags is None
    if own:
        flags = torch.zeros(1, dtype=torch.int32, device=x.device)
    _lib.check(lib.gabo_spd_factor(_p(x), n, d, 1 if is_mandel else 0, _p(fac), _p(flags), _lib.stream_ptr()),
               'gabo_spd_factor')
    if check and own:
        check_spd_flags(flags)
    return fac


def spd_factor_pair(x1, x2, d, is_mandel, flags):
    """Factor records of both operands of a Gram build with ONE launch (gabo_spd_factor2)."""
    lib = _lib.load()
    x1, x2 = to_dev64(x1), to_dev64(x2)
    fs = lib.gabo_spd_factor_stride(d)
    if fs < 0:
        raise ValueError('SPD(%d): matrix size outside [1, %d]' % (d, _lib.MAX_SPD_DIM))
    f1 = torch.empty(x1.shape[0], fs, dtype=torch.float64, device=x1.device)
    f2 = torch.empty(x2.shape[0], fs, dtype=torch.float64, device=x1.device)
    _lib.check(lib.gabo_spd_factor2(_p(x1), x1.shape[0], _p(x2), x2.shape[0], d, 1 if is_mandel else 0, _p(f1), _p(f2),
                                    _p(flags), _lib.stream_ptr()), 'gabo_spd_factor2')
    return f1, f2


def check_spd_flags(flags):
    if int(flags.item()) != 0:
        raise NotPositiveDefiniteError('input contains a matrix that is not positive definite')


def spd_ai_gram_from_factors(fac1, fac2, d, param=0.0, kind=_lib.KIND_GAUSS, compute=_lib.GABO_F32, symmetric=False,
                             out_dtype=torch.float64, out=None):
    lib = _lib.load()
    n1, n2 = fac1.shape[0], fac2.shape[0]
    if out is None:
        out = torch.empty(n1, n2, dtype=out_dtype, device=fac1.device)
    _lib.check(lib.gabo_spd_ai_gram(_p(fac1), n1, _p(fac2), n2, d, float(param), kind, compute,
                                    1 if symmetric else 0, _p(out), _DT[out.dtype], out.stride(0),
                                    _lib.stream_ptr()), 'gabo_spd_ai_gram')
    return out


def spd_ai_gram(x1, x2, param=0.0, kind=_lib.KIND_GAUSS, is_mandel=True, compute=_lib.GABO_F32,
                out_dtype=torch.float64, check=True, host_out=False):
    """f(d_AI(X1_i, X2_j)).  x: (..., N, dv) Mandel vectors (is_mandel) or (..., N, d, d) matrices -> (..., N1, N2).
    ``host_out``: the per-pair kernel stores straight into a pinned (device-mapped) HOST tensor, so the PCIe transfer
    of the result overlaps the arithmetic instead of following it; the returned tensor is then a CPU tensor."""
    same = x1 is x2
    trailing = 2 if is_mandel else 3
    x1 = to_dev64(x1)
    x2 = x1 if same else to_dev64(x2)
    d = mandel_dim(x1.shape[-1]) if is_mandel else x1.shape[-1]
    if x1.shape[-1] != x2.shape[-1]:
        raise ValueError('spd_ai_gram: x1 and x2 live on different SPD manifolds')
    if not same:
        x1, x2, lead = _broadcast_lead(x1, x2, trailing)
    else:
        lead = tuple(x1.shape[:x1.dim() - trailing])
    b1, _ = _flat_batches(x1, trailing)
    b2 = b1 if same else _flat_batches(x2, trailing)[0]
    n1, n2 = b1.shape[1], b2.shape[1]
    if host_out:
        out = torch.empty(b1.shape[0], n1, n2, dtype=out_dtype, pin_memory=True)
    else:
        out = torch.empty(b1.shape[0], n1, n2, dtype=out_dtype, device=x1.device)
    flags = torch.zeros(1, dtype=torch.int32, device=x1.device)
    for b in range(b1.shape[0]):
        if same:
            f1 = f2 = spd_factor(b1[b], d, is_mandel, flags=flags)
        else:
            f1, f2 = spd_factor_pair(b1[b], b2[b], d, is_mandel, flags)
        # the mirrored (symmetric) form writes columns: fine in HBM, wrong over PCIe
        spd_ai_gram_from_factors(f1, f2, d, param, kind, compute, symmetric=same and not host_out, out=out[b])
    if check:
        check_spd_flags(flags)      # one synchronisation, after everything has been enqueued
    return out.reshape(lead + (n1, n2))


class HostGramWorkspace:
    """Device-side staging of ``spd_ai_gram_host`` for one (n1, n2, dv) problem shape: the input copies, the factor
    records and the not-positive-definite flag (device + its pinned host mirror) are allocated once per kernel object and
    reused, so a host-to-host Gram build costs two asynchronous H2D copies, three launches and ONE synchronisation."""

    def __init__(self):
        self.key = None

    def get(self, dev, n1, n2, dv, d, fs):
        key = (dev, n1, n2, dv)
        if self.key != key:
            self.key = key
            self.x1 = torch.empty(n1, dv, dtype=torch.float64, device=dev)
            self.x2 = torch.empty(n2, dv, dtype=torch.float64, device=dev)
            self.f1 = torch.empty(n1, fs, dtype=torch.float64, device=dev)
            self.f2 = torch.empty(n2, fs, dtype=torch.float64, device=dev)
            self.flags = torch.zeros(1, dtype=torch.int32, device=dev)
            self.flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        return self


def spd_ai_gram_host(ws, x1, x2, param=0.0, kind=_lib.KIND_GAUSS, compute=_lib.GABO_F32, out_dtype=torch.float64):
    """f(d_AI(X1_i, X2_j)) for HOST Mandel vectors (n1, dv), (n2, dv) -> HOST (n1, n2) tensor in pinned memory.  The
    per-pair kernel stores straight into the pinned (device-mapped) result, so the PCIe transfer of the Gram matrix
    overlaps its computation (measured on the B200 box: 52 GB/s for the zero-copy stores vs 57 GB/s for a copy-engine
    transfer that could only start after the kernel).  One stream synchronisation at the end; the not-positive-definite
    flag travels in a pinned mirror and is read from host memory after that synchronisation."""
    lib = _lib.load()
    dev = device()
    n1, n2, dv = x1.shape[0], x2.shape[0], x1.shape[1]
    if x2.shape[1] != dv:
        raise ValueError('spd_ai_gram: x1 and x2 live on different SPD manifolds')
    d = mandel_dim(dv)
    fs = lib.gabo_spd_factor_stride(d)
    if fs < 0:
        raise ValueError('SPD(%d): matrix size outside [1, %d]' % (d, _lib.MAX_SPD_DIM))
    w = ws.get(dev, n1, n2, dv, d, fs)
    w.x1.copy_(x1, non_blocking=True)
    w.x2.copy_(x2, non_blocking=True)
    w.flags.zero_()
    s = _lib.stream_ptr()
    _lib.check(lib.gabo_spd_factor2(_p(w.x1), n1, _p(w.x2), n2, d, 1, _p(w.f1), _p(w.f2), _p(w.flags), s),
               'gabo_spd_factor2')
    w.flag_host.copy_(w.flags, non_blocking=True)
    out = torch.empty(n1, n2, dtype=out_dtype, pin_memory=True)
    _lib.check(lib.gabo_spd_ai_gram(_p(w.f1), n1, _p(w.f2), n2, d, float(param), kind, compute, 0, _p(out),
                                    _DT[out_dtype], n2, s), 'gabo_spd_ai_gram')
    torch.cuda.current_stream().synchronize()
    if int(w.flag_host[0]) != 0:
        raise NotPositiveDefiniteError('input contains a matrix that is not positive definite')
    return out


def spd_ai_gram_backward(fac1, fac2, d, w, transpose_w=False, compute=_lib.GABO_F32):
    """sum_j w_ij grad_{X1_i} d_AI^2(X1_i, X2_j) as (n1, d, d) fp64 matrices (gabo_spd_ai_gram_backward)."""
    lib = _lib.load()
    w = to_dev64(w)
    n1, n2 = fac1.shape[0], fac2.shape[0]
    out = torch.empty(n1, d, d, dtype=torch.float64, device=fac1.device)
    _lib.check(lib.gabo_spd_ai_gram_backward(_p(fac1), n1, _p(fac2), n2, d, _p(w), w.stride(0), 1 if transpose_w else 0,
                                             compute, _p(out), _lib.stream_ptr()), 'gabo_spd_ai_gram_backward')
    return out


# ----------------------------------------------------------------------------------------------------------------
# Frobenius / log-Euclidean
# ----------------------------------------------------------------------------------------------------------------

def spd_logm(mat):
    """Batched logm_torch (spd_utils_torch.py:13-30): (..., d, d) -> (..., d, d)."""
    lib = _lib.load()
    m = to_dev64(mat)
    d = m.shape[-1]
    flat = m.reshape(-1, d, d)
    out = torch.empty_like(flat)
    _lib.check(lib.gabo_spd_logm(_p(flat), flat.shape[0], d, _p(out), _lib.stream_ptr()), 'gabo_spd_logm')
    return out.reshape(m.shape)


def spd_logm_backward(mat, grad_out):
    """Adjoint of the Frechet derivative of logm at the SPD matrices ``mat`` applied to ``grad_out`` ((n, d, d) each)."""
    lib = _lib.load()
    m, g = to_dev64(mat), to_dev64(grad_out)
    d = m.shape[-1]
    fm, fg = m.reshape(-1, d, d), g.reshape(-1, d, d)
    out = torch.empty_like(fm)
    _lib.check(lib.gabo_spd_logm_backward(_p(fm), _p(fg), fm.shape[0], d, _p(out), _lib.stream_ptr()),
               'gabo_spd_logm_backward')
    return out.reshape(m.shape)


def weighted_points_sum(g, b, transpose=False, dist=None):
    """out_i = sum_j W_ij b_j for the (n1, n2) weights ``g`` (``transpose``: W = g^T, the gradient of the second
    operand) and the points ``b`` (cols, k).  With ``dist`` (same shape as g) the weights are the sphere-distance
    backward ``-g / sin(dist)``, zero where the reference's clamp is active."""
    lib = _lib.load()
    g, b = to_dev64(g), to_dev64(b)
    n1, n2 = int(g.shape[0]), int(g.shape[1])
    rows = n2 if transpose else n1
    k = int(b.shape[1])
    if int(b.shape[0]) != (n1 if transpose else n2):
        raise ValueError('weighted_points_sum: %d points for %d weights' % (b.shape[0], n1 if transpose else n2))
    out = torch.empty(rows, k, dtype=torch.float64, device=g.device)
    dd = None if dist is None else to_dev64(dist)
    _lib.check(lib.gabo_weighted_points_sum(_p(g), None if dd is None else _p(dd), n1, n2, n2, 1 if transpose else 0,
                                            0 if dd is None else 1, _p(b), k, _p(out), _lib.stream_ptr()),
               'gabo_weighted_points_sum')
    return out


def nested_spd_project_backward(x_mat, w, grad_y, want_x=True, want_w=True):
    """Backward of Y_n = W^T X_n W: (grad_x (n, D, D) or None, grad_w (D, d) or None) from grad_y (n, d, d)."""
    lib = _lib.load()
    w = to_dev64(w)
    gy = to_dev64(grad_y)
    D, d = int(w.shape[0]), int(w.shape[1])
    n = int(gy.shape[0])
    x = None if x_mat is None else to_dev64(x_mat)
    gx = torch.empty(n, D, D, dtype=torch.float64, device=w.device) if want_x else None
    gw = torch.empty(D, d, dtype=torch.float64, device=w.device) if want_w else None
    _lib.check(lib.gabo_nested_spd_project_backward(None if x is None else _p(x), _p(w), _p(gy), n, D, d,
                                                    None if gx is None else _p(gx), None if gw is None else _p(gw),
                                                    _lib.stream_ptr()), 'gabo_nested_spd_project_backward')
    return gx, gw


def spd_sqrtm(mat):
    """Batched sqrtm_torch (spd_utils_torch.py:33-50): (..., d, d) -> (..., d, d)."""
    lib = _lib.load()
    m = to_dev64(mat)
    d = m.shape[-1]
    flat = m.reshape(-1, d, d)
    out = torch.empty_like(flat)
    _lib.check(lib.gabo_spd_sqrtm(_p(flat), flat.shape[0], d, _p(out), _lib.stream_ptr()), 'gabo_spd_sqrtm')
    return out.reshape(m.shape)


def sym_eig(mat, vectors=True):
    """Batched symmetric eigendecomposition, D <= 32 (gabo_sym_eig): (..., D, D) -> eigenvalues (..., D) and, with
    ``vectors``, eigenvectors (..., D, D) (column k belongs to eigenvalue k; eigenvalues are NOT sorted).  The convergence
    flag stays on the device and is returned as a 1-element int32 tensor (no synchronisation here)."""
    lib = _lib.load()
    m = to_dev64(mat)
    D = int(m.shape[-1])
    if m.dim() < 2 or m.shape[-2] != D:
        raise ValueError('sym_eig: expected (..., D, D), got %s' % (tuple(m.shape),))
    flat = m.reshape(-1, D, D)
    lam = torch.empty(flat.shape[0], D, dtype=torch.float64, device=m.device)
    vec = torch.empty_like(flat) if vectors else None
    flag = torch.zeros(1, dtype=torch.int32, device=m.device)
    _lib.check(lib.gabo_sym_eig(_p(flat), flat.shape[0], D, _p(lam), _p(vec), _p(flag), _lib.stream_ptr()),
               'gabo_sym_eig')
    lam = lam.reshape(tuple(m.shape[:-1]))
    return (lam, vec.reshape(m.shape), flag) if vectors else (lam, None, flag)


def nested_spd_reconstruct_pack(w, v, c, k):
    """Point-independent factors of projection_from_nested_spd_to_spd for fixed (W, V, C, K)."""
    lib = _lib.load()
    w, v, c, k = (to_dev64(t) for t in (w, v, c, k))
    D, d = int(w.shape[0]), int(w.shape[1])
    m = D - d
    if tuple(v.shape) != (D, m) or tuple(c.shape) != (m, m) or tuple(k.shape) != (d, m):
        raise ValueError('expected W (D, d), V (D, D-d), C (D-d, D-d), K (d, D-d)')
    size = lib.gabo_nested_spd_reconstruct_pack_size(D, d)
    pack = torch.empty(max(size, 1), dtype=torch.float64, device=w.device)
    flag = torch.zeros(1, dtype=torch.int32, device=w.device)
    _lib.check(lib.gabo_nested_spd_reconstruct_setup(_p(w), _p(v), _p(c), _p(k), D, d, _p(pack), _p(flag),
                                                     _lib.stream_ptr()), 'gabo_nested_spd_reconstruct_setup')
    if int(flag.item()) != 0:
        raise NotPositiveDefiniteError('bottom_spd_matrix is not positive definite')
    return pack


def nested_spd_reconstruct(y, D, pack):
    """(n, d, d) latent SPD matrices -> (n, D, D)."""
    lib = _lib.load()
    y = to_dev64(y)
    d = y.shape[-1]
    flat = y.reshape(-1, d, d)
    sq = spd_sqrtm(flat)
    x = torch.empty(flat.shape[0], D, D, dtype=torch.float64, device=y.device)
    _lib.check(lib.gabo_nested_spd_reconstruct(_p(flat), _p(sq), flat.shape[0], D, d, _p(pack), _p(x),
                                               _lib.stream_ptr()), 'gabo_nested_spd_reconstruct')
    return x.reshape(tuple(y.shape[:-2]) + (D, D))


def frobenius_gram(m1, m2, param=0.0, kind=_lib.KIND_GAUSS, out_dtype=torch.float64):
    """f(||M1_i - M2_j + 1e-15||_F) for (..., N, d, d) matrices (spd_utils_torch.py:124-156)."""
    lib = _lib.load()
    m1, m2 = to_dev64(m1), to_dev64(m2)
    m1, m2, lead = _broadcast_lead(m1, m2, 3)
    d = m1.shape[-1]
    b1, _ = _flat_batches(m1, 3)
    b2, _ = _flat_batches(m2, 3)
    n1, n2 = b1.shape[1], b2.shape[1]
    out = torch.empty(b1.shape[0], n1, n2, dtype=out_dtype, device=m1.device)
    for b in range(b1.shape[0]):
        _lib.check(lib.gabo_frobenius_gram(_p(b1[b]), n1, _p(b2[b]), n2, d, float(param), kind, _p(out[b]),
                                           _DT[out_dtype], n2, _lib.stream_ptr()), 'gabo_frobenius_gram')
    return out.reshape(lead + (n1, n2))


# ----------------------------------------------------------------------------------------------------------------
# batched manifold operations (M1, M2)
# ----------------------------------------------------------------------------------------------------------------

def sphere_op(op, a, b, c=None):
    lib = _lib.load()
    a, b = to_dev64(a), to_dev64(b)
    c = to_dev64(c) if c is not None else None
    dim = a.shape[-1]
    fa, fb = a.reshape(-1, dim), b.reshape(-1, dim)
    fc = c.reshape(-1, dim) if c is not None else None
    out = torch.empty_like(fa)
    _lib.check(lib.gabo_sphere_op(op, _p(fa), _p(fb), _p(fc), fa.shape[0], dim, _p(out), _lib.stream_ptr()),
               'gabo_sphere_op')
    return out.reshape(a.shape)


def sphere_dist(x, y):
    lib = _lib.load()
    x, y = to_dev64(x), to_dev64(y)
    dim = x.shape[-1]
    fx, fy = x.reshape(-1, dim), y.reshape(-1, dim)
    out = torch.empty(fx.shape[0], dtype=torch.float64, device=x.device)
    _lib.check(lib.gabo_sphere_dist(_p(fx), _p(fy), fx.shape[0], dim, _p(out), _lib.stream_ptr()),
               'gabo_sphere_dist')
    return out.reshape(x.shape[:-1])


def nested_sphere_project(x, axes, dists):
    """S^{D-1} -> S^{dl-1} through the nested-sphere chain: x (..., D); axes: list of (1, k) or (k,) unit vectors for
    k = D, D-1, ..., dl+1; dists: one distance to the axis per level.  Returns (..., dl) fp64 on the device."""
    lib = _lib.load()
    x = to_dev64(x)
    D = x.shape[-1]
    dl = D - len(axes)
    flat = x.reshape(-1, D)
    y = torch.empty(flat.shape[0], dl, dtype=torch.float64, device=x.device)
    if len(axes) > 0:
        ax = torch.cat([to_dev64(a).reshape(-1) for a in axes]).to(x.device)
        for k, a in zip(range(D, dl, -1), axes):
            if torch.as_tensor(a).numel() != k:
                raise ValueError('axis of level %d must have %d components' % (k, k))
        ds = torch.cat([to_dev64(r).reshape(-1)[:1] for r in dists]).to(x.device)
    else:
        ax = ds = None
    _lib.check(lib.gabo_nested_sphere_project(_p(flat), flat.shape[0], D, dl, _p(ax), _p(ds), _p(y), _lib.stream_ptr()),
               'gabo_nested_sphere_project')
    return y.reshape(tuple(x.shape[:-1]) + (dl,))


def _nested_axes(axes, dists, D, device):
    """Concatenated axes / distances of the levels D, D-1, ... (each axis (1, k) or (k,))."""
    dl = D - len(axes)
    if len(dists) != len(axes):
        raise ValueError('need one distance to the axis per level')
    for k, a in zip(range(D, dl, -1), axes):
        if torch.as_tensor(a).numel() != k:
            raise ValueError('axis of level %d must have %d components' % (k, k))
    ax = torch.cat([to_dev64(a).reshape(-1) for a in axes]).to(device)
    ds = torch.cat([to_dev64(r).reshape(-1)[:1] for r in dists]).to(device)
    return ax, ds


def _split_levels(buf, n, dims, lead):
    out, off = [], 0
    for k in dims:
        out.append(buf[off:off + n * k].reshape(tuple(lead) + (k,)))
        off += n * k
    return out


def nested_sphere_chain(x, axes, dists):
    """Every level of the projection chain: x (..., D) -> [(..., D-1), ..., (..., dl)] fp64 on the device."""
    lib = _lib.load()
    x = to_dev64(x)
    D = x.shape[-1]
    if len(axes) == 0:
        return []
    ax, ds = _nested_axes(axes, dists, D, x.device)
    dl = D - len(axes)
    flat = x.reshape(-1, D)
    n = flat.shape[0]
    dims = list(range(D - 1, dl - 1, -1))
    buf = torch.empty(n * sum(dims), dtype=torch.float64, device=x.device)
    _lib.check(lib.gabo_nested_sphere_chain(_p(flat), n, D, dl, _p(ax), _p(ds), _p(buf), _lib.stream_ptr()),
               'gabo_nested_sphere_chain')
    return _split_levels(buf, n, dims, x.shape[:-1])


def nested_sphere_to_nested(x, axis, dist):
    """Closest points of the nested sphere {p: d(p, axis) = dist}: (..., k) -> (..., k)."""
    lib = _lib.load()
    x = to_dev64(x)
    k = x.shape[-1]
    ax = to_dev64(axis).reshape(-1).to(x.device)
    if ax.numel() != k:
        raise ValueError('axis must have %d components' % k)
    flat = x.reshape(-1, k)
    y = torch.empty_like(flat)
    _lib.check(lib.gabo_nested_sphere_to_nested(_p(flat), flat.shape[0], k, _p(ax), float(torch.as_tensor(dist).reshape(-1)[0]),
                                                _p(y), _lib.stream_ptr()), 'gabo_nested_sphere_to_nested')
    return y.reshape(x.shape)


def nested_sphere_reconstruct(y, axes, dists):
    """Every level of the inverse chain: y (..., dl) -> [(..., dl+1), ..., (..., D)]; axes / dists ordered as for the
    projection (levels D, D-1, ..., dl+1)."""
    lib = _lib.load()
    y = to_dev64(y)
    dl = y.shape[-1]
    if len(axes) == 0:
        return []
    D = dl + len(axes)
    ax, ds = _nested_axes(axes, dists, D, y.device)
    flat = y.reshape(-1, dl)
    n = flat.shape[0]
    dims = list(range(dl + 1, D + 1))
    buf = torch.empty(n * sum(dims), dtype=torch.float64, device=y.device)
    _lib.check(lib.gabo_nested_sphere_reconstruct(_p(flat), n, dl, D, _p(ax), _p(ds), _p(buf), _lib.stream_ptr()),
               'gabo_nested_sphere_reconstruct')
    return _split_levels(buf, n, dims, y.shape[:-1])


def spd_op(op, a, b, c=None):
    lib = _lib.load()
    a, b = to_dev64(a), to_dev64(b)
    c = to_dev64(c) if c is not None else None
    d = a.shape[-1]
    fa, fb = a.reshape(-1, d, d), b.reshape(-1, d, d)
    fc = c.reshape(-1, d, d) if c is not None else None
    out = torch.empty_like(fa)
    _lib.check(lib.gabo_spd_op(op, _p(fa), _p(fb), _p(fc), fa.shape[0], d, _p(out), _lib.stream_ptr()),
               'gabo_spd_op')
    return out.reshape(a.shape)


def spd_scalar(what, x, b, c=None):
    lib = _lib.load()
    x, b = to_dev64(x), to_dev64(b)
    c = to_dev64(c) if c is not None else None
    d = x.shape[-1]
    fx, fb = x.reshape(-1, d, d), b.reshape(-1, d, d)
    fc = c.reshape(-1, d, d) if c is not None else None
    out = torch.empty(fx.shape[0], dtype=torch.float64, device=x.device)
    _lib.check(lib.gabo_spd_scalar(what, _p(fx), _p(fb), _p(fc), fx.shape[0], d, _p(out), _lib.stream_ptr()),
               'gabo_spd_scalar')
    return out.reshape(x.shape[:-2])


# ----------------------------------------------------------------------------------------------------------------
# acquisition (A1, A3, A4)
# ----------------------------------------------------------------------------------------------------------------

class DeviceGP:
    """Device-resident description of the GP behind the acquisition function (struct gabo_gp_desc)."""

    def __init__(self, manifold, dim, x_train, alpha, minv, mean, outputscale, beta, best_f, kxx=1.0,
                 compute=_lib.GABO_F32):
        self.manifold = manifold
        self.dim = int(dim)
        self.x_train = x_train          # sphere: (n, D) fp64; spd: (n, factor_stride) fp64 factor records
        self.alpha = to_dev64(alpha)
        self.minv = to_dev64(minv)
        self.n_train = int(self.alpha.shape[0])
        if self.n_train > _lib.MAX_TRAIN:
            raise ValueError('at most %d training points are supported by the optimiser kernels' % _lib.MAX_TRAIN)
        self.desc = _lib.GpDesc(manifold, self.dim, self.n_train, compute, x_train.data_ptr(),
                                self.alpha.data_ptr(), self.minv.data_ptr(), float(mean), float(outputscale),
                                float(beta), float(best_f), float(kxx))

    @property
    def point_shape(self):
        return (self.dim,) if self.manifold == _lib.SPHERE else (self.dim, self.dim)

    def with_compute(self, compute):
        """The same GP (shared device arrays) evaluated in another arithmetic type (GABO_F32 | GABO_F64)."""
        if int(self.desc.compute) == int(compute):
            return self
        d = self.desc
        return DeviceGP(self.manifold, self.dim, self.x_train, self.alpha, self.minv, d.mean, d.outputscale, d.beta,
                        d.best_f, d.kxx, compute)


class TensorGP:
    """GP behind the acquisition function for the SPD kernels the solver kernels do not evaluate in closed form (the
    log-Euclidean and Frobenius Gaussian kernels: ``hd_gabo_spd.py`` builds its latent model on
    ``SpdLogEuclideanGaussianKernel``).  EI and its Riemannian gradient are differentiable DEVICE tensor code for all
    restarts at once: ``logm`` of the iterates and its adjoint are the ``gabo_spd_logm`` / ``gabo_spd_logm_backward`` kernels,
    the rest (n_train <= 128 kernel values, posterior mean / variance, EI) is a handful of small fp64 tensor operations;
    the gradient comes from one autograd pass.  Same quantities and conventions as ``gabo_ei_eval`` (botorch analytic EI,
    variance floor 1e-9, targets already sign-flipped for maximisation).  Used by the lock-step trust-region drivers."""

    is_tensor_gp = True
    manifold = _lib.SPD

    def __init__(self, dim, s_train, alpha, minv, mean, outputscale, inv_ls2, best_f, use_log):
        self.dim = int(dim)
        self.s_train = to_dev64(s_train)     # (n, d, d): logm of the training matrices (or the matrices themselves)
        self.alpha = to_dev64(alpha)
        self.minv = to_dev64(minv)
        self.n_train = int(self.alpha.shape[0])
        self.mean, self.outputscale, self.inv_ls2, self.best_f = float(mean), float(outputscale), float(inv_ls2), float(best_f)
        self.use_log = bool(use_log)
        d2self = self.dim * self.dim * 1e-30   # ||0 + 1e-15||_F^2: the reference adds 1e-15 to every entry of the difference
        self.kxx = math.exp(-d2self * self.inv_ls2)

    @property
    def point_shape(self):
        return (self.dim, self.dim)

    def with_compute(self, compute):
        return self

    #: evaluations of a fixed batch size are captured in a CUDA graph after this many eager calls (the lock-step solvers call
    #: the evaluator hundreds of times with all restarts at once; one evaluation is ~30 small launches)
    graph_after = 3

    def ei(self, x, want_grad=False):
        x = to_dev64(x)
        key = (int(x.shape[0]), bool(want_grad))
        cache = self.__dict__.setdefault('_graphs', {})
        entry = cache.get(key)
        if entry is None:
            entry = cache[key] = {'calls': 0, 'graph': None}
        if entry['graph'] is None and entry['calls'] >= 0 and x.is_cuda and 2 <= key[0] <= 4096:
            entry['calls'] += 1
            if entry['calls'] > self.graph_after:
                try:
                    static_x = x.clone()
                    side = torch.cuda.Stream(device=x.device)
                    side.wait_stream(torch.cuda.current_stream(x.device))
                    with torch.cuda.stream(side):
                        self._ei_eager(static_x, want_grad)
                    torch.cuda.current_stream(x.device).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        out = self._ei_eager(static_x, want_grad)
                    entry.update(graph=graph, x=static_x, out=out)
                except Exception:                       # noqa: BLE001 -- capture not possible here: stay eager
                    entry['calls'] = -1
        if entry['graph'] is not None:
            entry['x'].copy_(x)
            entry['graph'].replay()
            out = entry['out']
            return (out[0].clone(), out[1].clone()) if want_grad else out.clone()
        return self._ei_eager(x, want_grad)

    def _ei_eager(self, x, want_grad=False):
        from .kernel_utils import _SpdLogm
        xs = x.clone().requires_grad_(True) if want_grad else x
        with torch.enable_grad() if want_grad else torch.no_grad():
            s = _SpdLogm.apply(xs) if self.use_log else xs
            diff = s.unsqueeze(1) - self.s_train.unsqueeze(0) + 1e-15
            k = self.outputscale * torch.exp(-(diff * diff).sum((-1, -2)) * self.inv_ls2)       # (r, n)
            mk = k @ self.minv
            mu = self.mean + k @ self.alpha
            var = (self.outputscale * self.kxx - (mk * k).sum(-1)).clamp_min(1e-9)
            sigma = var.sqrt()
            u = (self.best_f - mu) / sigma
            pdf = torch.exp(-0.5 * u * u) * 0.3989422804014327
            cdf = 0.5 * torch.erfc(-u * 0.7071067811865476)
            ei = sigma * (pdf + u * cdf)
            if not want_grad:
                return ei
            e, = torch.autograd.grad(ei.sum(), xs)
        return ei.detach(), spd_op(_lib.OP_EGRAD2RGRAD, x, e)


def ei_eval(gp, x, want_grad=False):
    """EI (and its Riemannian gradient) at r points; x: (r, D) or (r, d, d)."""
    if getattr(gp, 'is_tensor_gp', False):
        return gp.ei(x, want_grad)
    lib = _lib.load()
    x = to_dev64(x)
    r = x.shape[0]
    ei = torch.empty(r, dtype=torch.float64, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    _lib.check(lib.gabo_ei_eval(ctypes.byref(gp.desc), _p(x), r, _p(ei), _p(grad), _lib.stream_ptr()),
               'gabo_ei_eval')
    return (ei, grad) if want_grad else ei


def acq_rcg(gp, x0, maxiter=1000, mingradnorm=1e-6, minstepsize=1e-10, ls_maxiter=10, contraction=0.5,
            suff_decr=0.5, initial_stepsize=1.0):
    """Multi-start Riemannian CG on -EI.  Returns (candidates, values, iters, reasons)."""
    lib = _lib.load()
    x = to_dev64(x0).clone()
    r = x.shape[0]
    val = torch.empty(r, dtype=torch.float64, device=x.device)
    iters = torch.empty(r, dtype=torch.int32, device=x.device)
    reason = torch.empty(r, dtype=torch.int32, device=x.device)
    opts = _lib.RcgOpts(int(maxiter), int(ls_maxiter), float(mingradnorm), float(minstepsize), float(contraction),
                        float(suff_decr), float(initial_stepsize))
    _lib.check(lib.gabo_acq_rcg(ctypes.byref(gp.desc), _p(x), r, ctypes.byref(opts), _p(val), _p(iters), _p(reason),
                                _lib.stream_ptr()), 'gabo_acq_rcg')
    return x, val, iters, reason


def acq_rtr(gp, x0, maxiter=1000, mingradnorm=1e-6, kappa=0.1, theta=1.0, rho_prime=0.1, rho_regularization=1e3,
            mininner=1, maxinner=None, delta_bar=None, delta0=None):
    """Multi-start Riemannian trust regions (tCG, finite-difference Hessian) on -EI: spheres the register kernel covers
    and SPD(d) (fp64, one warp per restart).  Returns (candidates, values, iters, reasons)."""
    lib = _lib.load()
    x = to_dev64(x0).clone()
    r = x.shape[0]
    val = torch.empty(r, dtype=torch.float64, device=x.device)
    iters = torch.empty(r, dtype=torch.int32, device=x.device)
    reason = torch.empty(r, dtype=torch.int32, device=x.device)
    opts = _lib.RtrOpts(int(maxiter), int(mininner), int(maxinner or 0), 0, float(mingradnorm), float(kappa),
                        float(theta), float(rho_prime), float(rho_regularization), float(delta_bar or 0.0),
                        float(delta0 or 0.0))
    _lib.check(lib.gabo_acq_rtr(ctypes.byref(gp.desc), _p(x), r, ctypes.byref(opts), _p(val), _p(iters), _p(reason),
                                _lib.stream_ptr()), 'gabo_acq_rtr')
    return x, val, iters, reason


def acq_ctr(gp, x0, constraints=(), strict=False, delta_cons=1e-6, maxiter=1000, mingradnorm=1e-6, kappa=0.1,
            theta=1.0, rho_prime=0.1, rho_regularization=1e3, mininner=1, maxinner=None, delta_bar=None, delta0=None):
    """Multi-start [Strict]ConstrainedTrustRegions on SPD(d) in one launch (``gabo_acq_ctr``).  ``constraints``: up to
    two ``('max' | 'min', bound)`` eigenvalue inequality constraints; none = plain TrustRegions.  Returns (candidates,
    values, iters, reasons)."""
    lib = _lib.load()
    x = to_dev64(x0).clone()
    r = x.shape[0]
    val = torch.empty(r, dtype=torch.float64, device=x.device)
    iters = torch.empty(r, dtype=torch.int32, device=x.device)
    reason = torch.empty(r, dtype=torch.int32, device=x.device)
    constraints = list(constraints)
    if len(constraints) > 2:
        raise ValueError('gabo_acq_ctr takes at most two eigenvalue constraints')
    opts = _lib.CtrOpts()
    opts.tr = _lib.RtrOpts(int(maxiter), int(mininner), int(maxinner or 0), 0, float(mingradnorm), float(kappa),
                           float(theta), float(rho_prime), float(rho_regularization), float(delta_bar or 0.0),
                           float(delta0 or 0.0))
    opts.n_constraints, opts.strict, opts.delta_cons = len(constraints), int(bool(strict)), float(delta_cons)
    for i, (kind, bound) in enumerate(constraints):
        opts.kind[i] = _lib.CONS_MAX_EIG if kind == 'max' else _lib.CONS_MIN_EIG
        opts.bound[i] = float(bound)
    _lib.check(lib.gabo_acq_ctr(ctypes.byref(gp.desc), _p(x), r, ctypes.byref(opts), _p(val), _p(iters), _p(reason),
                                _lib.stream_ptr()), 'gabo_acq_ctr')
    return x, val, iters, reason


def gp_mll(dmat, y, theta, want_grad=True, want_factors=False):
    """log N(y | m, s exp(-beta dmat) + noise I) for a batch of theta = (beta, s, noise, m) rows (``gabo_gp_mll``).
    Returns (ll (B,), grad (B, 4) or None, alpha (B, n) or None, kinv (B, n, n) or None, flags (B,) int32), on device."""
    lib = _lib.load()
    dmat = to_dev64(dmat)
    y = to_dev64(y).reshape(-1)
    theta = to_dev64(theta).reshape(-1, 4)
    n, B = y.shape[0], theta.shape[0]
    if tuple(dmat.shape) != (n, n):
        raise ValueError('dmat must be (n, n) with n = len(y)')
    dev = dmat.device
    ll = torch.empty(B, dtype=torch.float64, device=dev)
    grad = torch.empty(B, 4, dtype=torch.float64, device=dev) if want_grad else None
    alpha = torch.empty(B, n, dtype=torch.float64, device=dev) if want_factors else None
    kinv = torch.empty(B, n, n, dtype=torch.float64, device=dev) if want_factors else None
    flags = torch.zeros(B, dtype=torch.int32, device=dev)
    _lib.check(lib.gabo_gp_mll(_p(dmat), n, _p(y), _p(theta), B, _p(ll), _p(grad), _p(alpha), _p(kinv), _p(flags),
                               _lib.stream_ptr()), 'gabo_gp_mll')
    return ll, grad, alpha, kinv, flags


def gp_fit(dmat, y, raw0, beta_min, noise_min, priors, fixed, maxiter=15000, pgtol=1e-5, ftol=2.220446049250313e-09):
    """``gabo_gp_fit``: BFGS fits of (raw_beta, raw_outputscale, raw_noise, mean) from the starts ``raw0`` (B, 4), all in
    one launch.  ``priors``: 6 floats (Gamma concentration, rate for beta, outputscale, noise; concentration <= 0: none);
    ``fixed``: 4 ints.  Returns host numpy (raw (B, 4), f (B,), info (B, 3) = status, iterations, evaluations) after ONE
    read-back."""
    lib = _lib.load()
    dmat = to_dev64(dmat)
    y = to_dev64(y).reshape(-1)
    raw0 = to_dev64(raw0).reshape(-1, 4)
    n, B = y.shape[0], raw0.shape[0]
    if tuple(dmat.shape) != (n, n):
        raise ValueError('dmat must be (n, n) with n = len(y)')
    dev = dmat.device
    pri = torch.tensor([float(v) for v in priors], dtype=torch.float64)
    fix = torch.tensor([int(v) for v in fixed], dtype=torch.int32)
    # one packed result buffer: raw (4), f (1), info as doubles (3) per start -> a single device-to-host copy
    out_raw = torch.empty(B, 4, dtype=torch.float64, device=dev)
    out_f = torch.empty(B, dtype=torch.float64, device=dev)
    out_info = torch.empty(B, 3, dtype=torch.int32, device=dev)
    _lib.check(lib.gabo_gp_fit(_p(dmat), n, _p(y), _p(raw0), B, float(beta_min), float(noise_min), _p(pri), _p(fix),
                               int(maxiter), float(pgtol), float(ftol), _p(out_raw), _p(out_f), _p(out_info),
                               _lib.stream_ptr()), 'gabo_gp_fit')
    packed = torch.cat([out_raw, out_f[:, None], out_info.double()], dim=1).cpu().numpy()
    return packed[:, :4], packed[:, 4], packed[:, 5:].astype(int)


def gp_factor(kmat, y, outputscale, noise, mean):
    """alpha = (s K + noise I)^-1 (y - m) and (s K + noise I)^-1 from a base-kernel matrix (``gabo_gp_factor``)."""
    lib = _lib.load()
    kmat = to_dev64(kmat)
    y = to_dev64(y).reshape(-1)
    n = y.shape[0]
    if tuple(kmat.shape) != (n, n):
        raise ValueError('kmat must be (n, n) with n = len(y)')
    alpha = torch.empty(n, dtype=torch.float64, device=kmat.device)
    kinv = torch.empty(n, n, dtype=torch.float64, device=kmat.device)
    flag = torch.zeros(1, dtype=torch.int32, device=kmat.device)
    _lib.check(lib.gabo_gp_factor(_p(kmat), n, _p(y), float(outputscale), float(noise), float(mean), _p(alpha),
                                  _p(kinv), _p(flag), _lib.stream_ptr()), 'gabo_gp_factor')
    if int(flag.item()) != 0:
        raise NotPositiveDefiniteError('the GP covariance s K + noise I is not positive definite')
    return alpha, kinv


def argmax_records(values, gidx=None):
    """(slot, value) of the best record: highest value, lowest global index on ties, NaN = -inf."""
    lib = _lib.load()
    v = to_dev64(values).reshape(-1)
    g = None
    if gidx is not None:
        g = torch.as_tensor(gidx).to(device=v.device, dtype=torch.int64).contiguous()
    slot = torch.empty(1, dtype=torch.int64, device=v.device)
    best = torch.empty(1, dtype=torch.float64, device=v.device)
    _lib.check(lib.gabo_argmax_records(_p(v), _p(g), v.shape[0], _p(slot), _p(best), _lib.stream_ptr()),
               'gabo_argmax_records')
    return slot, best


# ----------------------------------------------------------------------------------------------------------------
# nested projection (P1)
# ----------------------------------------------------------------------------------------------------------------

def nested_projection_matrix(w):
    """Packed tf32 hi/lo Mandel projection operator for a (D, d) projection W (nested_spd_utils.py:13-48 in Mandel form)."""
    lib = _lib.load()
    w = to_dev64(w)
    D, d = w.shape
    size = lib.gabo_nested_projection_pack_size(D, d)
    if size < 0:
        raise ValueError('nested projection SPD(%d) -> SPD(%d) is not supported (d <= %d)' % (D, d, _lib.MAX_SPD_DIM))
    p = torch.empty(size, dtype=torch.float32, device=w.device)
    _lib.check(lib.gabo_nested_projection_matrix(_p(w), D, d, _p(p), _lib.stream_ptr()),
               'gabo_nested_projection_matrix')
    return p


def nested_spd_project_f64(x_mandel, w):
    """Mandel(W^T X W) in fp64 for (..., D(D+1)/2) Mandel vectors and a (D, d) projection matrix."""
    lib = _lib.load()
    x = to_dev64(x_mandel)
    w = to_dev64(w).to(x.device)
    D, d = int(w.shape[0]), int(w.shape[1])
    if mandel_dim(x.shape[-1]) != D:
        raise ValueError('Mandel length %d does not match a %d x %d matrix' % (x.shape[-1], D, D))
    flat = x.reshape(-1, x.shape[-1])
    y = torch.empty(flat.shape[0], d * (d + 1) // 2, dtype=torch.float64, device=x.device)
    _lib.check(lib.gabo_nested_spd_project_f64(_p(flat), flat.shape[0], D, d, _p(w), _p(y), _lib.stream_ptr()),
               'gabo_nested_spd_project_f64')
    return y.reshape(tuple(x.shape[:-1]) + (y.shape[-1],))


def nested_spd_project(x_mandel, D, d, p_padded):
    """y_mandel (n, dvl) = x_mandel (n, dvh) P^T on the tensor cores; fp32 in / out."""
    lib = _lib.load()
    x = torch.as_tensor(x_mandel)
    if not x.is_cuda:
        x = x.to(device())
    x = x.to(torch.float32).contiguous()
    n = x.shape[0]
    y = torch.empty(n, d * (d + 1) // 2, dtype=torch.float32, device=x.device)
    _lib.check(lib.gabo_nested_spd_project(_p(x), n, D, d, _p(p_padded), _p(y), _lib.stream_ptr()),
               'gabo_nested_spd_project')
    return y
