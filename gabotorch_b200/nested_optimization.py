"""Reconstruction-parameter fit of the nested-sphere mapping.

Mirrors ``BoManifolds/nested_mappings/nested_spheres_optimization.py`` of the reference:
``min_error_reconstruction_cost`` (:20-38) and ``optimize_reconstruction_parameters_nested_sphere`` (:41-98) -- the step
of hd_gabo_sphere.py (:196-199) between the latent acquisition optimiser and ``projection_from_subsphere_to_sphere``: the
distances-to-axis r_k in (0, pi) of every level are chosen so that the data reconstructed from their latent projections are
as close as possible (sum of squared geodesic distances) to the data.  The reference treats the r_k through a sigmoid
(gpytorch ``Interval(0, pi)``) as a product of ``Euclidean(1)`` manifolds, screens ``nb_init_candidates`` random draws and
hands the best to a pymanopt solver (``TrustRegions()`` in the example).

Here the cost and its gradient are device work: the inverse chain (rotation of the north pole to each axis applied in O(k),
nested_spheres_utils.py:149-213) and the row-wise geodesic distance run as fp64 tensor code on the B200 for ALL candidates at
once (one pass, one read-back), and for the solver's evaluations with torch.autograd over the same code; the solver itself is
the host loop of ``manifold_gp_fit`` (pymanopt ``TrustRegions`` / ``ConjugateGradient`` semantics, finite-difference
Hessian), a handful of scalars.

The SPD counterpart mirrors ``BoManifolds/nested_mappings/nested_spd_optimization.py``:
``min_affine_invariant_distance_reconstruction_cost`` (:22-55), ``min_log_euclidean_distance_reconstruction_cost`` (:58-92)
and ``optimize_reconstruction_parameters_nested_spd`` (:95-186, the step of hd_gabo_spd.py:230-233): the complement V of the
projection, the bottom block C and the contraction K = sigmoid(s) * k (|k| = 1) are fitted on
Grassmann(D, D - d) x SPD(D - d) x Sphere(d (D - d)) x R by the augmented Lagrangian method under W^T V = 0.  The data
matrices are D x D (10 .. 20 in the reference's examples), beyond the d <= 8 register kernels: every cost evaluation is ONE
launch of ``gabo_sym_eig`` (a warp per matrix, Jacobi in shared memory) over all N reconstructed matrices -- the reference
runs N ``torch.symeig`` calls in a Python loop -- wrapped in autograd Functions whose backward is the Daleckii-Krein formula
on the saved eigenpairs; the small products around it are batched fp64 device matmuls.  All ``nb_init_candidates`` random
starts are screened in one batched pass.
"""
import math

import numpy as np
import torch

from . import ops
from .kernel_utils import _dev64_keep_grad
from .manifold_gp_fit import (EuclideanParam, GrassmannParam, ProductParam, SphereParam, SpdParam, riemannian_alm,
                              solve_on_manifold)


def _rotate(v, p_from, p_to):
    """R v for the rotation R that moves the unit vector ``p_from`` to ``p_to`` along their geodesic
    (rotation_from_sphere_points_torch, sphere_utils_torch.py:58-93), applied without forming R; v: (..., N, k)."""
    ct = (p_from * p_to).sum().clamp(-1.0 + 1e-15, 1.0 - 1e-15)
    c = p_from - p_to * ct
    c = c / c.norm()
    st = torch.sin(torch.acos(ct))
    vc, vy = (v * c).sum(-1, keepdim=True), (v * p_to).sum(-1, keepdim=True)
    return v + st * (vc * p_to - vy * c) + (ct - 1.0) * (vy * p_to + vc * c)


def _reconstruct(x_sub, axes, dists):
    """projection_from_subsphere_to_sphere (nested_spheres_utils.py:184-213), last level only; ``dists``: (..., r) so that a
    leading candidate dimension broadcasts: x_sub (N, d) -> (..., N, D)."""
    x = x_sub
    nb = len(axes)
    for s in range(nb):
        axis = axes[nb - s - 1]
        r = dists[..., nb - s - 1][..., None, None]
        north = torch.zeros_like(axis)
        north[-1] = 1.0
        lifted = torch.cat([torch.sin(r) * x, torch.cos(r) * torch.ones_like(x[..., :1])], dim=-1)
        x = _rotate(lifted, north, axis)
    return x


def _cost_from_distances(x_data, x_sub, axes, dists):
    xr = _reconstruct(x_sub, axes, dists)
    inner = (x_data * xr).sum(-1).clamp(-1.0 + 1e-15, 1.0 - 1e-15)      # sphere_distance_torch(diag=True)
    d = torch.acos(inner)
    return (d * d).sum(-1)


def min_error_reconstruction_cost(x_data, x_subsphere, sphere_axes, sphere_distances):
    """Sum of squared geodesic distances between the data and their reconstruction from the subsphere
    (nested_spheres_optimization.py:20-38).  Differentiable device code; returns a 0-d fp64 tensor."""
    xd = _dev64_keep_grad(x_data)
    xs = _dev64_keep_grad(x_subsphere).to(xd.device)
    axes = [_dev64_keep_grad(a).reshape(-1).to(xd.device) for a in sphere_axes]
    dists = torch.stack([_dev64_keep_grad(r).reshape(()).to(xd.device) for r in sphere_distances])
    return _cost_from_distances(xd, xs, axes, dists)


def optimize_reconstruction_parameters_nested_sphere(x_data, x_subsphere, sphere_axes, solver, nb_init_candidates=100):
    """Distances-to-axis of ``projection_from_subsphere_to_sphere`` that minimise the reconstruction error
    (nested_spheres_optimization.py:41-98).  Returns a list of (1,) float32 tensors like the reference
    (``radius_constraint.transform(torch.Tensor(distance))``)."""
    xd = ops.to_dev64(x_data)
    xs = ops.to_dev64(x_subsphere)
    axes = [ops.to_dev64(a).reshape(-1) for a in sphere_axes]
    nlev = len(axes)
    if xd.shape[-1] - xs.shape[-1] != nlev:
        raise ValueError('need one axis per level: data on S^%d, subsphere data on S^%d, %d axes'
                         % (xd.shape[-1] - 1, xs.shape[-1] - 1, nlev))
    manifold = ProductParam([EuclideanParam(1) for _ in range(nlev)])

    def transform(p):                                   # gpytorch Interval(0, pi): sigmoid(raw) * pi
        return torch.sigmoid(p) * math.pi

    # candidate screening: every candidate in one pass over the device code, one read-back
    cands = [manifold.rand() for _ in range(int(nb_init_candidates))]
    raw = torch.from_numpy(np.array([[float(c[0]) for c in cand] for cand in cands])).to(xd.device)
    with torch.no_grad():
        vals = _cost_from_distances(xd, xs, axes, transform(raw)).cpu().numpy()
    x0 = cands[int(np.argmin(vals))]

    def cost(x):
        with torch.no_grad():
            p = torch.tensor([float(v[0]) for v in x], dtype=torch.float64, device=xd.device)
            return float(_cost_from_distances(xd, xs, axes, transform(p)))

    def cost_grad(x):
        p = torch.tensor([float(v[0]) for v in x], dtype=torch.float64, device=xd.device, requires_grad=True)
        with torch.enable_grad():
            f = _cost_from_distances(xd, xs, axes, transform(p))
            f.backward()
        g = p.grad.cpu().numpy()
        return float(f.detach()), [np.array([gi]) for gi in g]

    opt, log = solve_on_manifold(manifold, cost, cost_grad, x0, solver)
    out = [transform(torch.tensor(np.asarray(v), dtype=torch.float32).reshape(1)) for v in opt]
    optimize_reconstruction_parameters_nested_sphere.last_log = dict(log, start_cost=float(vals.min()))
    return out


# ----------------------------------------------------------------------------------------------------------------
# nested SPD mapping
# ----------------------------------------------------------------------------------------------------------------

class _SpectralFn(torch.autograd.Function):
    """U f(lambda) U^T for a batch of symmetric D x D matrices (D <= 32) from ONE ``gabo_sym_eig`` launch; ``kind`` 0: f = log
    (logm_torch, spd_utils_torch.py:13-30), 1: f = sqrt (sqrtm_torch, :33-50).  Backward: Daleckii-Krein,
    U [(U^T sym(G) U) o F] U^T with F_ij the divided differences of f (log: log1p form, sqrt: 1 / (f_i + f_j) -- no
    cancellation for close eigenvalues)."""

    @staticmethod
    def forward(ctx, mat, kind):
        lam, vec, _ = ops.sym_eig(mat)
        f = torch.log(lam) if kind == 0 else torch.sqrt(lam)
        ctx.save_for_backward(lam, vec, f)
        ctx.kind = kind
        return (vec * f.unsqueeze(-2)) @ vec.transpose(-1, -2)

    @staticmethod
    def backward(ctx, g):
        lam, vec, f = ctx.saved_tensors
        gs = 0.5 * (g + g.transpose(-1, -2))
        inner = vec.transpose(-1, -2) @ gs @ vec
        if ctx.kind == 0:
            lj = lam.unsqueeze(-2)
            x = (lam.unsqueeze(-1) - lj) / lj
            small = x.abs() < 1e-9
            dd = torch.where(small, 1.0 - 0.5 * x, torch.log1p(x) / torch.where(small, torch.ones_like(x), x)) / lj
        else:
            dd = 1.0 / (f.unsqueeze(-1) + f.unsqueeze(-2))
        return vec @ (inner * dd) @ vec.transpose(-1, -2), None


class _LogEigSumSq(torch.autograd.Function):
    """sum_k log^2 lambda_k(M) for a batch of symmetric positive definite matrices (the squared affine-invariant distance
    when M = L^-1 X L^-T, spd_utils_torch.py:100-120); backward U diag(2 log(lambda) / lambda) U^T."""

    @staticmethod
    def forward(ctx, mat):
        lam, vec, _ = ops.sym_eig(mat)
        ll = torch.log(lam)
        ctx.save_for_backward(lam, vec, ll)
        return (ll * ll).sum(-1)

    @staticmethod
    def backward(ctx, g):
        lam, vec, ll = ctx.saved_tensors
        w = (2.0 * ll / lam) * g.unsqueeze(-1)
        return (vec * w.unsqueeze(-2)) @ vec.transpose(-1, -2)


def _reconstruct_spd(y, y_sqrt, w, v, c, k):
    """projection_from_nested_spd_to_spd (nested_spd_utils.py:51-118) as differentiable device code: y, y_sqrt (N, d, d);
    w (D, d); v (..., D, m), c (..., m, m), k (..., d, m) with optional leading candidate dimensions -> (..., N, D, D)."""
    lead = v.shape[:-2]
    n = y.shape[0]
    c_sqrt = _SpectralFn.apply(c, 1)
    side = y_sqrt @ (k @ c_sqrt).unsqueeze(-3)                               # (..., N, d, m)
    top = torch.cat([y.expand(lead + tuple(y.shape)), side], dim=-1)
    bottom = torch.cat([side.transpose(-1, -2), c.unsqueeze(-3).expand(lead + (n,) + tuple(c.shape[-2:]))], dim=-1)
    xr = torch.cat([top, bottom], dim=-2)
    r = torch.cat([w.expand(lead + tuple(w.shape)), v], dim=-1).unsqueeze(-3)
    return r @ xr @ r.transpose(-1, -2)


class _SpdReconstructionCost:
    """The two reconstruction costs with the data-only terms (sqrt of the latent matrices; inverse Cholesky factors or
    logarithms of the data) computed once.  ``kind``: 'affine_invariant' or 'log_euclidean'."""

    def __init__(self, x_data, x_data_projected, projection_matrix, kind):
        self.kind = kind
        self.x = ops.to_dev64(x_data)
        self.y = ops.to_dev64(x_data_projected).to(self.x.device)
        self.w = ops.to_dev64(projection_matrix).to(self.x.device)
        if self.x.dim() != 3 or self.y.dim() != 3 or self.x.shape[0] != self.y.shape[0]:
            raise ValueError('expected x_data (N, D, D) and x_data_projected (N, d, d)')
        self.D, self.d = int(self.x.shape[-1]), int(self.y.shape[-1])
        if tuple(self.w.shape) != (self.D, self.d):
            raise ValueError('projection_matrix must be (D, d) = (%d, %d)' % (self.D, self.d))
        with torch.no_grad():
            self.y_sqrt = _SpectralFn.apply(self.y, 1)
            if kind == 'affine_invariant':
                # X = L L^T from the eigenpairs: L^-1 may be ANY factor with L^-1 X L^-T = I (the eigenvalues of
                # L^-1 Xrec L^-T do not depend on the choice); X^(-1/2) = U diag(lambda^-1/2) U^T
                lam, vec, flag = ops.sym_eig(self.x)
                if int(flag.item()) != 0 or not bool((lam > 0).all()):         # one read-back, at construction only
                    raise ops.NotPositiveDefiniteError('x_data contains a matrix that is not symmetric positive definite')
                self.linv = (vec * lam.rsqrt().unsqueeze(-2)) @ vec.transpose(-1, -2)
            elif kind == 'log_euclidean':
                self.logx = _SpectralFn.apply(self.x, 0)
            else:
                raise ValueError('kind must be affine_invariant or log_euclidean')

    def __call__(self, v, c, k):
        xrec = _reconstruct_spd(self.y, self.y_sqrt, self.w, v, c, k)
        if self.kind == 'affine_invariant':
            m = self.linv @ xrec @ self.linv
            return (_LogEigSumSq.apply(m) + 1e-15).sum(-1)
        diff = self.logx - _SpectralFn.apply(xrec, 0) + 1e-15                 # frobenius_distance_torch, :156
        return (diff * diff).sum((-1, -2)).sum(-1)


def _spd_cost(kind, x_data, x_data_projected, projection_matrix, projection_complement_matrix, bottom_spd_matrix,
              contraction_matrix):
    fn = _SpdReconstructionCost(x_data, x_data_projected, projection_matrix, kind)
    dev = fn.x.device
    return fn(_dev64_keep_grad(projection_complement_matrix).to(dev), _dev64_keep_grad(bottom_spd_matrix).to(dev),
              _dev64_keep_grad(contraction_matrix).to(dev))


def min_affine_invariant_distance_reconstruction_cost(x_data, x_data_projected, projection_matrix,
                                                      projection_complement_matrix, bottom_spd_matrix,
                                                      contraction_matrix):
    """Sum of squared affine-invariant distances between the SPD data and their reconstruction from the projections
    Y = W^T X W (nested_spd_optimization.py:22-55).  Differentiable device code, 0-d fp64 tensor."""
    return _spd_cost('affine_invariant', x_data, x_data_projected, projection_matrix, projection_complement_matrix,
                     bottom_spd_matrix, contraction_matrix)


def min_log_euclidean_distance_reconstruction_cost(x_data, x_data_projected, projection_matrix,
                                                   projection_complement_matrix, bottom_spd_matrix, contraction_matrix):
    """Same with the log-Euclidean distance ||logm X - logm Xrec + 1e-15||_F (nested_spd_optimization.py:58-92)."""
    return _spd_cost('log_euclidean', x_data, x_data_projected, projection_matrix, projection_complement_matrix,
                     bottom_spd_matrix, contraction_matrix)


def optimize_reconstruction_parameters_nested_spd(x_data, x_data_projected, projection_matrix, inner_solver,
                                                  cost_function=min_affine_invariant_distance_reconstruction_cost,
                                                  nb_init_candidates=100, maxiter=50, use_cuda_graph=True):
    """Parameters (V, C, K) of ``projection_from_nested_spd_to_spd`` that minimise the reconstruction error of the data
    (nested_spd_optimization.py:95-186): augmented Lagrangian (``lambdas_fact=0.05``, ``maxiter`` outer iterations) around
    ``inner_solver`` on Grassmann(D, D - d) x SPD(D - d) x Sphere(d (D - d)) x R under ||V^T W|| = 0, started from the best
    of ``nb_init_candidates`` random points.  Returns fp64 CPU tensors (D, D - d), (D - d, D - d), (d, D - d)."""
    from .manifold_optimization import AugmentedLagrangeMethod
    if cost_function is min_affine_invariant_distance_reconstruction_cost:
        kind = 'affine_invariant'
    elif cost_function is min_log_euclidean_distance_reconstruction_cost:
        kind = 'log_euclidean'
    else:
        raise NotImplementedError('cost_function must be one of the two reconstruction costs of this module')
    fn = _SpdReconstructionCost(x_data, x_data_projected, projection_matrix, kind)
    dev, D, d = fn.x.device, fn.D, fn.d
    m = D - d
    if not 1 <= m or D > 32:
        raise ValueError('need d < D <= 32, got D=%d, d=%d' % (D, d))
    manifold = ProductParam([GrassmannParam(D, m), SpdParam(m), SphereParam(d * m), EuclideanParam(1)])
    w_host = fn.w.cpu().numpy()

    def to_device(points, requires_grad=False):
        """list of manifold points (or a list of such lists) -> V, C, k, s device tensors (leading candidate dimension)."""
        batch = isinstance(points[0], (list, tuple))
        cols = list(zip(*points)) if batch else [[p] for p in points]
        ts = [torch.from_numpy(np.ascontiguousarray(np.array(col, dtype=np.float64))).to(dev) for col in cols]
        if not batch:
            ts = [t[0] for t in ts]
        return [t.requires_grad_(requires_grad) for t in ts]

    def evaluate(v, c, kvec, s):
        kmat = torch.sigmoid(s).unsqueeze(-1) * kvec.reshape(kvec.shape[:-1] + (d, m))   # Interval(0, 1) of the norm
        return fn(v, c, kmat)

    # candidate screening: all candidates in one batched pass (the reference evaluates them one after the other)
    cands = [manifold.rand() for _ in range(int(nb_init_candidates))]
    with torch.no_grad():
        vals = evaluate(*to_device(cands)).cpu().numpy()
    x0 = cands[int(np.nanargmin(vals))]

    # One cost + gradient evaluation is ~60 small launches (eigensolver, products, element-wise steps) on tiny operands:
    # launch-bound.  The solver calls it hundreds of times with operands of fixed shape, so the whole forward + backward
    # pass is captured ONCE in a CUDA graph over static input / output tensors and replayed per evaluation (one packed
    # host-to-device copy in, one read-back out).  Any capture problem falls back to the eager evaluation.
    sizes = [D * m, m * m, d * m, 1]
    shapes = [(D, m), (m, m), (d * m,), (1,)]
    graph = None
    if dev.type == 'cuda' and use_cuda_graph:
        try:
            packed_in = torch.zeros(sum(sizes), dtype=torch.float64, device=dev)
            stat = [packed_in[sum(sizes[:i]):sum(sizes[:i + 1])].view(shapes[i]).requires_grad_(True) for i in range(4)]
            x0_dev = torch.from_numpy(np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in x0])).to(dev)
            with torch.no_grad():
                packed_in.copy_(x0_dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):               # warm-up outside the capture (lazy initialisations)
                for _ in range(2):
                    torch.autograd.grad(evaluate(*stat), stat)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                f_stat = evaluate(*stat)
                g_stat = torch.autograd.grad(f_stat, stat)
                packed_out = torch.cat([f_stat.detach().reshape(1)] + [gi.detach().reshape(-1) for gi in g_stat])
        except Exception:                               # noqa: BLE001 -- e.g. capture-unsafe allocator state: eager path
            graph = None

    def pack_point(x):
        return torch.from_numpy(np.concatenate([np.asarray(a, dtype=np.float64).ravel() for a in x]))

    def cost_grad(x):
        if graph is not None:
            with torch.no_grad():
                packed_in.copy_(pack_point(x))
            graph.replay()
            out = packed_out.detach().cpu().numpy()
            grads, at = [], 1
            for size, shape in zip(sizes, shapes):
                grads.append(out[at:at + size].reshape(shape).copy())
                at += size
            return float(out[0]), grads
        ts = to_device(x, requires_grad=True)
        with torch.enable_grad():
            f = evaluate(*ts)
            f.backward()
        return float(f.detach()), [t.grad.cpu().numpy() for t in ts]

    def cost(x):
        if graph is not None:
            return cost_grad(x)[0]
        with torch.no_grad():
            return float(evaluate(*to_device(x)))

    def orthogonality(x):                                   # ||V^T W||_F and its gradient (zero at the feasible point)
        vtw = x[0].T @ w_host
        val = float(np.linalg.norm(vtw))
        gv = (w_host @ vtw.T) / val if val > 0 else np.zeros_like(x[0])
        return val, [gv, np.zeros_like(x[1]), np.zeros_like(x[2]), np.zeros_like(x[3])]

    solver = AugmentedLagrangeMethod(maxiter=maxiter, inner_solver=inner_solver, lambdas_fact=0.05)
    opt, log = riemannian_alm(manifold, cost, cost_grad, x0, solver, eq_constraints=[orthogonality])
    v = torch.from_numpy(np.array(opt[0], dtype=np.float64))
    c = torch.from_numpy(np.array(opt[1], dtype=np.float64))
    norm = torch.sigmoid(torch.from_numpy(np.array(opt[3], dtype=np.float64)))
    kmat = norm * torch.from_numpy(np.array(opt[2], dtype=np.float64)).view(d, m)
    optimize_reconstruction_parameters_nested_spd.last_log = dict(log, start_cost=float(np.nanmin(vals)),
                                                                  cuda_graph=graph is not None)
    return v, c, kmat
