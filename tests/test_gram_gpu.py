"""Parity of the fused Gram kernels (through the C ABI) with the oracle and with the golden vectors generated from the
reference's own code.  Tolerances: distances |d - d_ref| <= 1e-5 d_ref + 1e-6; kernel values relative
1e-5 on entries >= 1e-6 for the sphere and for SPD with fp64 Jacobi; the fp32 Jacobi path is held to the bound the
distance tolerance implies through the exponential, 1e-5 * max(1, 2 beta d^2) (the reference itself only carries
float32 eigenvalues, spd_utils_torch.py:108)."""
import math

import numpy as np
import pytest
import torch

from gabotorch_b200 import _lib, ops
from oracle import spd as ospd
from oracle import sphere as osph

pytestmark = pytest.mark.gpu

DIST_RTOL, DIST_ATOL, K_RTOL, K_FLOOR = 1e-5, 1e-6, 1e-5, 1e-6


def check_dist(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    bad = np.abs(got - ref) - (DIST_RTOL * np.abs(ref) + DIST_ATOL)
    assert bad.max() <= 0, 'distance off by %.3e' % np.abs(got - ref).max()


def check_kernel(got, ref, rtol=K_RTOL, amplification=None):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    m = ref >= K_FLOOR
    rel = np.abs(got - ref)[m] / ref[m]
    tol = rtol if amplification is None else rtol * np.maximum(1.0, amplification[m])
    assert np.all(rel <= tol), 'kernel rel err %.3e' % rel.max()
    assert np.abs(got - ref)[~m].max(initial=0.0) <= 2e-6 * rtol / K_RTOL


@pytest.mark.parametrize('name', ['s2_n256', 's5_n1024', 's8_n96'])
@pytest.mark.parametrize('out_dtype', [torch.float32, torch.float64])
def test_sphere_gram_golden(golden, name, out_dtype):
    x = torch.from_numpy(golden[name + '_x'])
    st = int(golden[name + '_stride'])
    beta = float(golden[name + '_beta'])
    d = ops.sphere_gram(x, x, kind=_lib.KIND_DIST, out_dtype=out_dtype).cpu().numpy()
    k = ops.sphere_gram(x, x, beta, _lib.KIND_GAUSS, out_dtype=out_dtype).cpu().numpy()
    check_dist(d[::st, ::st], golden[name + '_d'])
    check_kernel(k[::st, ::st], golden[name + '_k'])
    assert abs(k.astype(np.float64).sum() - float(golden[name + '_ksum'])) <= 1e-6 * k.size   # checksum over ALL pairs
    dd = ops.sphere_gram(x, x, kind=_lib.KIND_DIST, diag=True).cpu().numpy()
    assert dd.shape == (x.shape[0], 1)
    check_dist(dd, golden[name + '_ddiag'])


@pytest.mark.parametrize('D,n1,n2', [(2, 5, 7), (3, 300, 313), (4, 37, 53), (6, 257, 1029), (7, 64, 64), (9, 130, 77),
                                     (12, 33, 600), (13, 40, 41), (20, 70, 83), (64, 19, 23), (128, 9, 11)])
def test_sphere_gram_vs_oracle_ragged(D, n1, n2):
    rng = np.random.default_rng(100 + D)
    x, y = osph.rand(rng, n1, D), osph.rand(rng, n2, D)
    y[0] = x[0]                                              # identical
    y[1] = -x[1]                                             # antipodal
    t = x[2] + 1e-5 * rng.standard_normal(D)
    y[2] = t / np.linalg.norm(t)                             # nearly identical
    t = -x[3] + 1e-4 * rng.standard_normal(D)
    y[3] = t / np.linalg.norm(t)                             # nearly antipodal
    beta = 1.0 + math.log(2.0)
    check_dist(ops.sphere_gram(x, y, kind=_lib.KIND_DIST).cpu(), osph.sphere_distance(x, y))
    check_kernel(ops.sphere_gram(x, y, beta, _lib.KIND_GAUSS).cpu(), osph.sphere_gaussian_kernel(x, y, beta))
    check_kernel(ops.sphere_gram(x, y, 1.0 / 0.49, _lib.KIND_LAPLACE).cpu(), osph.sphere_laplace_kernel(x, y, 0.7))


def test_sphere_gram_golden_rect_and_edge_sizes(golden):
    a, b = golden['s3_rect_a'], golden['s3_rect_b']
    check_dist(ops.sphere_gram(a, b, kind=_lib.KIND_DIST).cpu(), golden['s3_rect_d'])
    # empty inputs
    e = ops.sphere_gram(torch.zeros(0, 3, dtype=torch.float64), torch.from_numpy(a[:, :3] * 0 + 1), 1.0)
    assert tuple(e.shape) == (0, a.shape[0])
    # single point, batch dimensions, non-contiguous input
    x = torch.from_numpy(a)
    one = ops.sphere_gram(x[:1], x[:1], kind=_lib.KIND_DIST).cpu()
    assert tuple(one.shape) == (1, 1) and float(one) < 1e-6
    xb = torch.stack([x, x.flip(0)])                          # (2, 37, 4)
    kb = ops.sphere_gram(xb, torch.from_numpy(b), 1.3).cpu()
    assert tuple(kb.shape) == (2, 37, 53)
    np.testing.assert_allclose(kb[1].numpy(), osph.sphere_gaussian_kernel(a[::-1].copy(), b, 1.3).numpy(), rtol=1e-5)
    xt = torch.from_numpy(np.ascontiguousarray(a.T)).T        # non-contiguous view
    check_dist(ops.sphere_gram(xt, b, kind=_lib.KIND_DIST).cpu(), golden['s3_rect_d'])


def test_sphere_gram_batched_odd_sizes_and_offset_views():
    # the reference accepts b1 x ... x N x D (sphere_utils_torch.py:12-55): batch slices of a contiguous buffer sit at
    # b*N*D*8 bytes, which is only 8-byte aligned when N*D is odd; so is a row-offset view x[1:] with odd D
    rng = np.random.default_rng(17)
    xb = osph.rand(rng, 10, 3).reshape(2, 5, 3)                  # N*D = 15
    yb = osph.rand(rng, 14, 3).reshape(2, 7, 3)                  # N*D = 21
    got = ops.sphere_gram(xb, yb, kind=_lib.KIND_DIST).cpu()
    assert tuple(got.shape) == (2, 5, 7)
    for b in range(2):
        check_dist(got[b], osph.sphere_distance(xb[b], yb[b]))
    # gpytorch's posterior shape: b x 1 x D test points against N x D training points
    xt, tr = osph.rand(rng, 6, 3).reshape(6, 1, 3), osph.rand(rng, 9, 3)
    kb = ops.sphere_gram(xt, tr, 2.0, _lib.KIND_GAUSS).cpu()
    assert tuple(kb.shape) == (6, 1, 9)
    check_kernel(kb[:, 0], osph.sphere_gaussian_kernel(xt[:, 0], tr, 2.0))
    # contiguous device views that start 8 (not 16) bytes into the allocation, both operands, both output dtypes
    x = torch.from_numpy(osph.rand(rng, 301, 3)).cuda()
    y = torch.from_numpy(osph.rand(rng, 1031, 5)[:, :3].copy())
    y = (y / y.norm(dim=-1, keepdim=True)).cuda()
    for od in (torch.float32, torch.float64):
        full = ops.sphere_gram(x, y, 1.3, _lib.KIND_GAUSS, out_dtype=od)
        view = ops.sphere_gram(x[1:], y[1:], 1.3, _lib.KIND_GAUSS, out_dtype=od)
        assert x[1:].data_ptr() % 16 == 8
        assert torch.equal(view, full[1:, 1:])
    dd = ops.sphere_gram(xb, xb.copy(), kind=_lib.KIND_DIST, diag=True).cpu()
    assert tuple(dd.shape) == (2, 5, 1) and float(dd.max()) < 1e-6


def test_sphere_gram_non_unit_inputs_follow_the_reference_formula():
    # the reference never normalises: it clamps the raw inner product (sphere_utils_torch.py:53)
    rng = np.random.default_rng(9)
    x = osph.rand(rng, 50, 3) * (1.0 + 0.05 * rng.standard_normal((50, 1)))
    y = osph.rand(rng, 60, 3) * 0.97
    check_dist(ops.sphere_gram(x, y, kind=_lib.KIND_DIST).cpu(), osph.sphere_distance(x, y))


def test_sphere_gram_full_size_properties():
    # BASELINE-size checks through size-independent properties: symmetry, unit diagonal, row-block consistency
    N, D, beta = 16384, 3, 6.5 + math.log(2.0)
    g = torch.Generator().manual_seed(1234)
    x = torch.nn.functional.normalize(torch.randn(N, D, dtype=torch.float64, generator=g), dim=-1)
    k = ops.sphere_gram(x, x, beta, _lib.KIND_GAUSS, out_dtype=torch.float32)
    assert float((k - k.T).abs().max()) <= 2e-6
    assert float((k.diagonal() - 1).abs().max()) <= 1e-6
    blk = ops.sphere_gram(x[5000:5300], x, beta, _lib.KIND_GAUSS, out_dtype=torch.float32)
    assert torch.equal(blk, k[5000:5300])                     # row-block sharding reproduces the same bits
    sub = osph.sphere_gaussian_kernel(x[5000:5064].numpy(), x[:4096].numpy(), beta)
    check_kernel(k[5000:5064, :4096].cpu(), sub)


@pytest.mark.parametrize('name', ['spd3_n128', 'spd8_n64', 'spd2_n40', 'spd5_n48'])
@pytest.mark.parametrize('compute', [_lib.GABO_F32, _lib.GABO_F64])
def test_spd_gram_golden(golden, name, compute):
    v = torch.from_numpy(golden[name + '_vec'])
    beta = float(golden[name + '_beta'])
    dref, kref = golden[name + '_d'], golden[name + '_k']
    d = ops.spd_ai_gram(v, v, kind=_lib.KIND_DIST, compute=compute).cpu().numpy()
    check_dist(d, dref)
    k = ops.spd_ai_gram(v, v, beta, _lib.KIND_GAUSS, compute=compute).cpu().numpy()
    if compute == _lib.GABO_F64:
        check_kernel(k, kref)
    else:
        check_kernel(k, kref, amplification=2.0 * beta * dref * dref)
    # x1 is not x2 -> general (non-mirrored) tile path must give the same numbers as the symmetric path
    k2 = ops.spd_ai_gram(v, v.clone(), beta, _lib.KIND_GAUSS, compute=compute).cpu().numpy()
    np.testing.assert_allclose(np.triu(k2), np.triu(k), rtol=2e-5, atol=1e-9)


@pytest.mark.parametrize('d,n1,n2', [(1, 9, 11), (2, 100, 131), (3, 200, 257), (4, 33, 129), (5, 96, 40), (6, 70, 65),
                                     (7, 20, 150), (8, 80, 90)])
def test_spd_gram_vs_oracle_ragged(d, n1, n2):
    rng = np.random.default_rng(200 + d)
    X, Y = ospd.spd_sample(rng, n1, d, max_cond=100.0), ospd.spd_sample(rng, n2, d, max_cond=100.0)
    Y[0] = X[0]
    beta = 0.3 + math.log(2.0)
    dref = ospd.affine_invariant_distance(X, Y).numpy()
    for compute in (_lib.GABO_F32, _lib.GABO_F64):
        dg = ops.spd_ai_gram(X, Y, kind=_lib.KIND_DIST, is_mandel=False, compute=compute).cpu().numpy()
        check_dist(dg, dref)
        kg = ops.spd_ai_gram(X, Y, beta, _lib.KIND_GAUSS, is_mandel=False, compute=compute).cpu().numpy()
        kref = np.exp(-beta * dref * dref)
        check_kernel(kg, kref, amplification=None if compute == _lib.GABO_F64 else 2.0 * beta * dref * dref)
        lg = ops.spd_ai_gram(X, Y, beta, _lib.KIND_LAPLACE, is_mandel=False, compute=compute).cpu().numpy()
        check_kernel(lg, np.exp(-beta * dref))
    assert abs(dg[0, 0] - math.sqrt(1e-15)) < 1e-6            # d(X,X) = sqrt(1e-15), spd_utils_torch.py:120


def test_spd_gram_rect_golden_and_errors(golden):
    a, b = golden['spd3_rect_a'], golden['spd3_rect_b']
    check_dist(ops.spd_ai_gram(a, b, kind=_lib.KIND_DIST, is_mandel=False).cpu(), golden['spd3_rect_d'])
    bad = a.copy()
    bad[3] = -bad[3]                                          # not positive definite: torch.cholesky raises in the reference
    with pytest.raises(ops.NotPositiveDefiniteError):
        ops.spd_ai_gram(bad, b, kind=_lib.KIND_DIST, is_mandel=False)
    with pytest.raises(ValueError):
        ops.spd_ai_gram(torch.zeros(4, 5, dtype=torch.float64), torch.zeros(4, 5, dtype=torch.float64))


def test_spd_gram_full_size_properties():
    # BASELINE config 2: SPD(3), N = 2048
    rng = np.random.default_rng(1234)
    N, d, beta = 2048, 3, 0.5 + math.log(2.0)
    X = ospd.spd_sample(rng, N, d, max_cond=100.0)
    v = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(X))
    k = ops.spd_ai_gram(v, v, beta, _lib.KIND_GAUSS, out_dtype=torch.float32)
    assert float((k - k.T).abs().max()) == 0.0                 # mirrored tiles: exactly symmetric
    assert float((k.diagonal() - 1).abs().max()) <= 1e-6
    d_full = ops.spd_ai_gram(v, v, kind=_lib.KIND_DIST, compute=_lib.GABO_F64)
    # affine invariance: d(A X A^T, A Y A^T) = d(X, Y)
    A = rng.standard_normal((d, d)) + 2 * np.eye(d)
    Xc = A @ X[:256] @ A.T
    dc = ops.spd_ai_gram(Xc, Xc, kind=_lib.KIND_DIST, is_mandel=False, compute=_lib.GABO_F64)
    assert float((dc - d_full[:256, :256]).abs().max()) <= 2e-5
    # inversion invariance and a sample against the oracle
    di = ops.spd_ai_gram(np.linalg.inv(X[:256]), np.linalg.inv(X[:256]), kind=_lib.KIND_DIST, is_mandel=False,
                         compute=_lib.GABO_F64)
    assert float((di - d_full[:256, :256]).abs().max()) <= 2e-5
    check_dist(d_full[:96, 1000:1100].cpu(), ospd.affine_invariant_distance(X[:96], X[1000:1100]))
    lam_min = torch.linalg.eigvalsh(k[:512, :512].double()).min()
    assert float(lam_min) > -5e-7                              # PD at beta >= beta_min (spd_gaussian_kernel_parameters.py:50-53)


@pytest.mark.parametrize('d', [2, 3])
def test_spd_gram_closed_form_hard_cases(d):
    # d = 2, 3 in fp32 take closed-form eigenvalues (largest of W and of W^-1, the middle one from det W): the cases
    # where a trigonometric / quadratic formula is fragile, all pairs against the oracle
    rng = np.random.default_rng(77 + d)
    sets = {}
    base = ospd.spd_sample(rng, 1, d, max_cond=100.0)[0]
    near = [base]
    for eps in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-9, 1e-12):
        for _ in range(12):
            e = rng.standard_normal((d, d))
            near.append(base + 0.5 * (e + e.T) * eps)
    sets['near-identical'] = np.stack(near)
    q, _ = np.linalg.qr(rng.standard_normal((96, d, d)))
    lam = []
    for a in (0.01, 0.1, 1.0, 3.0):
        for b in (0.011, 0.5, 1.0001, 1.0 + 1e-7, 5.0):
            lam += [[a] * (d - 1) + [b], [b] + [a] * (d - 1), [a] * d]
    lam = np.array(lam)[:96]
    sets['double / triple eigenvalues'] = (q[:len(lam)] * lam[:, None, :]) @ np.swapaxes(q[:len(lam)], -1, -2)
    sets['cond up to 5000'] = ospd.spd_sample(rng, 160, d, max_cond=1e9)
    sets['diagonal (commuting)'] = np.stack([np.diag(rng.uniform(0.001, 5, d)) for _ in range(64)])
    sets['scaled identities'] = np.stack([np.eye(d) * s for s in (1e-3, 0.5, 1.0, 2.0, 1e3)])
    # scales far apart: intermediate quantities leave the window of the closed form -> the Jacobi route takes over
    far = ospd.spd_sample(rng, 40, d, max_cond=100.0)
    sets['scales far apart'] = far * np.repeat([1e-7, 1e-3, 1.0, 1e3, 1e7], 8)[:, None, None]
    beta = 0.5 + math.log(2.0)
    for name, X in sets.items():
        X = 0.5 * (X + np.swapaxes(X, -1, -2))
        dref = ospd.affine_invariant_distance(X, X).numpy()
        dg = ops.spd_ai_gram(X, X.copy(), kind=_lib.KIND_DIST, is_mandel=False).cpu().numpy()
        err = np.abs(dg - dref) - (1e-5 * dref + 1e-6)
        assert err.max() <= 0, '%s: distance bound violated by %.3e' % (name, err.max())
        kg = ops.spd_ai_gram(X, X.copy(), beta, _lib.KIND_GAUSS, is_mandel=False).cpu().numpy()
        check_kernel(kg, np.exp(-beta * dref * dref), amplification=2.0 * beta * dref * dref)
        assert np.abs(np.diag(dg) - math.sqrt(1e-15)).max() < 1e-6


@pytest.mark.parametrize('d,beta_min', [(3, 0.5), (8, 0.22)])
def test_spd_gram_full_matrix_parity_at_the_benchmarked_size(d, beta_min):
    # BASELINE configs[1] (SPD(3), N = 2048: the configuration bench.py times) and the SPD(8) extra: EVERY one of the
    # 4.2 M pairs against the vectorised oracle (faithful to spd_utils_torch.py:87-120 incl. the float32 eigenvalues).
    # The achieved errors are printed (pytest -s) and returned to bench.py / smoke() through the same helper.
    import bench
    N, beta = 2048, beta_min + math.log(2.0)
    v = bench.spd_sample_mandel(np.random.default_rng(1234), N, d)       # bench.py's input law and seed
    rep, rep64 = bench.spd_parity_report(v, beta, compute=[_lib.GABO_F32, _lib.GABO_F64])
    print('SPD(%d) N=%d fp32 compute: %s' % (d, N, rep))
    print('SPD(%d) N=%d fp64 compute: %s' % (d, N, rep64))
    assert rep['pairs_checked'] == N * N
    assert rep['max_dist_margin'] <= 0.0                       # |d - d_ref| <= 1e-5 d_ref + 1e-6 on all pairs
    assert rep['max_rel_err_K_amplified'] <= 1.0               # rel err <= 1e-5 max(1, 2 beta d^2) on K >= 1e-6
    assert rep64['max_dist_margin'] <= 0.0
    assert rep64['max_rel_err_K'] <= 1e-5                      # flat north_star bound with the fp64 eigen-solve


@pytest.mark.parametrize('name', ['spd3_n128', 'spd8_n64', 'spd5_n48'])
def test_mandel_frobenius_logm_golden(golden, name):
    m, v = golden[name + '_mat'], golden[name + '_vec']
    assert torch.equal(ops.mandel_pack(m).cpu(), torch.from_numpy(v))                      # bit-exact
    assert torch.equal(ops.mandel_unpack(v).cpu(), torch.from_numpy(golden[name + '_unpacked']))
    back = golden[name + '_unpacked']
    np.testing.assert_allclose(ops.frobenius_gram(back, back, kind=_lib.KIND_DIST).cpu().numpy(), golden[name + '_frob'],
                               rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(ops.spd_logm(back).cpu().numpy(), golden[name + '_logm'], rtol=0, atol=1e-10)
    ls = 1.7
    np.testing.assert_allclose(ops.frobenius_gram(back, back, 1.0 / ls ** 2, _lib.KIND_GAUSS).cpu().numpy(),
                               ospd.spd_frobenius_gaussian_kernel(v, v, ls).numpy(), rtol=1e-12)
    lg = ops.spd_logm(back)
    np.testing.assert_allclose(ops.frobenius_gram(lg, lg, 1.0 / ls ** 2, _lib.KIND_GAUSS).cpu().numpy(),
                               ospd.spd_log_euclidean_gaussian_kernel(v, v, ls).numpy(), rtol=1e-9, atol=1e-12)


def test_mandel_large_dim_roundtrip():
    rng = np.random.default_rng(2)
    m = rng.standard_normal((33, 20, 20))
    m = m + np.swapaxes(m, -1, -2)
    v = ops.mandel_pack(m)
    assert tuple(v.shape) == (33, 210)
    assert torch.equal(v.cpu(), ospd.symmetric_matrix_to_vector_mandel(m))
    np.testing.assert_allclose(ops.mandel_unpack(v).cpu().numpy(), m, rtol=0, atol=4e-16 * np.abs(m).max())
