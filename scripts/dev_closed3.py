"""CPU prototype (numpy float32 emulation) of the 'bilateral closed form' for the SPD(3) / SPD(2) affine-invariant
distance: largest eigenvalue of W = G G^T and of W^-1 = H^T H (H = G^-1 = A_j L_i, both from the factor records) by the
trigonometric formula (the inverse through the adjugate of the triangular G), the middle one from det W = (g00 g11 g22)^2.  Checks the distance bound of SURVEY 8(d) against the
fp64 oracle on the benchmark law and on hard cases.  Development aid, not part of the product."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from oracle import spd as ospd

f32 = np.float32


def _fit_g(deg=8):
    """cos(acos(r)/3) = g(s), s = sqrt((1 + r)/2) = cos(3 phi / 2): g(s) = cos(2/3 acos(s)) is analytic on [0, 1]
    (nearest singularity s = -1), so a short polynomial in s does it.  Chebyshev-node least squares."""
    from numpy.polynomial import chebyshev as C
    k = np.arange(8 * deg); u = np.cos(np.pi * (k + 0.5) / (8 * deg)); sn = (u + 1) / 2
    return np.polynomial.polynomial.polyfit(sn, np.cos((2.0 / 3.0) * np.arccos(sn)), deg)

G_COEF = _fit_g()


def cos_third_acos(r):
    """float32 emulation of the device code: s = sqrt(0.5 r + 0.5), Horner in s."""
    s = np.sqrt(np.clip(r * f32(0.5) + f32(0.5), f32(0), f32(1))).astype(f32)
    acc = np.full_like(s, f32(G_COEF[-1]))
    for c in G_COEF[-2::-1]:
        acc = (acc * s + f32(c)).astype(f32)
    return acc


def lam_max_sym3(m00, m11, m22, m01, m02, m12, g=cos_third_acos):
    third = f32(1.0 / 3.0)
    q = (m00 + m11 + m22) * third
    b00, b11, b22 = m00 - q, m11 - q, m22 - q
    p2 = (b00 * b00 + b11 * b11 + b22 * b22 + f32(2) * (m01 * m01 + m02 * m02 + m12 * m12)) * f32(1.0 / 6.0)
    p2 = np.maximum(p2, f32(1e-37))
    ip = (f32(1) / np.sqrt(p2)).astype(f32)
    det = b00 * (b11 * b22 - m12 * m12) + m01 * (m12 * m02 - m01 * b22) + m02 * (m01 * m12 - b11 * m02)   # unscaled B
    r = np.clip(((det * ip) * ip) * (ip * f32(0.5)), f32(-1), f32(1))
    p = p2 * ip
    return q + f32(2) * p * g(r)


def dist3(fac_L, fac_A, i_idx, j_idx):
    """fac_L, fac_A: (N,3,3) fp64 lower-triangular.  Returns d for the pairs (i_idx, j_idx)."""
    G = (fac_A[i_idx] @ fac_L[j_idx]).astype(f32)          # lower triangular, rounded to fp32 as on the device
    G64 = fac_A[i_idx] @ fac_L[j_idx]
    x = (G64[:, 1, 0] * G64[:, 2, 1] - G64[:, 2, 0] * G64[:, 1, 1]).astype(f32)   # the one cancelling adjugate entry, fp64
    H = np.zeros_like(G)                                   # D adj(G) D, D = diag(1,-1,1): five products + x
    H[:, 0, 0] = G[:, 1, 1] * G[:, 2, 2]; H[:, 1, 1] = G[:, 0, 0] * G[:, 2, 2]; H[:, 2, 2] = G[:, 0, 0] * G[:, 1, 1]
    H[:, 1, 0] = G[:, 1, 0] * G[:, 2, 2]; H[:, 2, 1] = G[:, 2, 1] * G[:, 0, 0]; H[:, 2, 0] = x
    def gram_rows(G):   # M = G G^T for lower-triangular G
        g00, g10, g11, g20, g21, g22 = G[:, 0, 0], G[:, 1, 0], G[:, 1, 1], G[:, 2, 0], G[:, 2, 1], G[:, 2, 2]
        return (g00 * g00, g10 * g10 + g11 * g11, g20 * g20 + g21 * g21 + g22 * g22,
                g00 * g10, g00 * g20, g10 * g20 + g11 * g21)
    def gram_cols(G):   # M = G^T G
        g00, g10, g11, g20, g21, g22 = G[:, 0, 0], G[:, 1, 0], G[:, 1, 1], G[:, 2, 0], G[:, 2, 1], G[:, 2, 2]
        return (g00 * g00 + g10 * g10 + g20 * g20, g11 * g11 + g21 * g21, g22 * g22,
                g10 * g11 + g20 * g21, g20 * g22, g21 * g22)
    l1 = lam_max_sym3(*gram_rows(G))
    m1 = lam_max_sym3(*gram_cols(H))
    ldet = f32(2) * np.log2(G[:, 0, 0] * G[:, 1, 1] * G[:, 2, 2])
    a = np.log2(l1)
    c = ldet - np.log2(m1)                                 # m1 = det W / lambda_min
    b = ldet - a - c
    s = (a * a + b * b + c * c) * f32(0.48045301391820142) + f32(1e-15)
    return np.sqrt(s).astype(np.float64)


def check(mats, name):
    mats = torch.as_tensor(mats)
    n = mats.shape[0]
    L = torch.linalg.cholesky(mats)
    A = torch.inverse(L)
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    d = dist3(L.numpy(), A.numpy(), ii.ravel(), jj.ravel()).reshape(n, n)
    dref = ospd.affine_invariant_distance(mats, mats).numpy()
    err = np.abs(d - dref)
    margin = (err - (1e-5 * dref + 1e-6)).max()
    big = dref > 1e-3
    print('%-34s n=%4d  max|d-dref| %.3e  margin %.3e (<=0 ok)  max rel (d>1e-3) %.3e  diag %.3e' %
          (name, n, err.max(), margin, (err[big] / dref[big]).max() if big.any() else 0.0, d.diagonal().max()))
    return margin


rng = np.random.default_rng(1234)
v = bench.spd_sample_mandel(rng, 512, 3)
m = ospd.vector_to_symmetric_matrix_mandel(torch.from_numpy(v))
check(m, 'bench law (cond<=100)')
# near-identical matrices
base = m[:1].numpy()
pert = []
for eps in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7):
    for _ in range(20):
        e = rng.standard_normal((3, 3)); e = 0.5 * (e + e.T) * eps
        pert.append(base[0] + e)
check(np.stack([base[0]] + pert), 'near-identical')
# isotropic and double eigenvalues
q, _ = np.linalg.qr(rng.standard_normal((60, 3, 3)))
lam = np.stack([np.array([1.0, 1.0, 1.0]) * s for s in (0.5, 1, 2)] + [np.array([a, a, b]) for a in (0.1, 1, 3) for b in (0.01, 0.5, 1.0001, 5)]
               + [np.array([a, b, b]) for a in (0.1, 1, 3) for b in (0.011, 0.5, 1.0001, 5)])
lam = np.concatenate([lam, lam])[:60]
check((q[:lam.shape[0]] * lam[:, None, :]) @ np.swapaxes(q[:lam.shape[0]], -1, -2), 'double / triple eigenvalues')
# extreme conditioning: cond up to 5000 per matrix
v2 = bench.spd_sample_mandel(rng, 256, 3, max_cond=1e9)
check(ospd.vector_to_symmetric_matrix_mandel(torch.from_numpy(v2)), 'no cond filter (<=5000)')
# diagonal matrices (commuting)
check(np.stack([np.diag(rng.uniform(0.001, 5, 3)) for _ in range(128)]), 'diagonal')

print('g coefficients (s^0 ..):', ', '.join('%.9ef' % c for c in G_COEF))
