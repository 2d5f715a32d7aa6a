"""Nested SPD projection Y = W^T X W (P1): tensor-core Mandel contraction against the reference's bmm outputs."""
import numpy as np
import pytest
import torch

from gabotorch_b200 import nested_mappings as nm
from gabotorch_b200 import ops
from oracle import nested as onest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['proj_20_5', 'proj_5_2'])
def test_projection_golden(golden, name):
    x, w = golden[name + '_x'], golden[name + '_w']
    y = nm.projection_from_spd_to_nested_spd(torch.from_numpy(x), torch.from_numpy(w))
    assert y.dtype == torch.float64 and not y.is_cuda
    scale = np.abs(golden[name + '_y']).max()
    np.testing.assert_allclose(y.numpy(), golden[name + '_y'], rtol=0, atol=1e-5 * scale)
    ym = nm.projection_mandel(torch.from_numpy(golden[name + '_xvec']), w).cpu().numpy()
    np.testing.assert_allclose(ym, golden[name + '_yvec'], rtol=0, atol=1e-5 * np.abs(golden[name + '_yvec']).max())


@pytest.mark.parametrize('D,d,n', [(20, 5, 1), (20, 5, 63), (20, 5, 64), (20, 5, 65), (20, 5, 30000), (5, 2, 1001),
                                   (6, 3, 777), (12, 8, 500), (8, 8, 130), (3, 1, 50), (22, 4, 257)])
def test_projection_vs_oracle_ragged(D, d, n):
    rng = np.random.default_rng(D * 100 + d)
    dvh = D * (D + 1) // 2
    xv = rng.standard_normal((n, dvh))
    w = onest.grassmann_rand(rng, D, d)
    P = onest.mandel_projection_matrix(w)
    ref = xv.astype(np.float32).astype(np.float64) @ P.T
    got = nm.projection_mandel(torch.from_numpy(xv), w).cpu().numpy()
    assert got.shape == ref.shape and got.dtype == np.float32
    np.testing.assert_allclose(got, ref, rtol=0, atol=3e-6 * np.abs(ref).max())


def test_projection_linearity_and_spd_preservation():
    rng = np.random.default_rng(3)
    D, d, n = 20, 5, 4096
    proj = nm.NestedSpdProjection(onest.grassmann_rand(rng, D, d))
    a = torch.from_numpy(rng.standard_normal((n, 210))).float()
    b = torch.from_numpy(rng.standard_normal((n, 210))).float()
    lhs = proj.mandel(2.0 * a + b)
    rhs = 2.0 * proj.mandel(a) + proj.mandel(b)
    assert float((lhs - rhs).abs().max()) <= 2e-5 * float(rhs.abs().max())
    m = rng.standard_normal((256, D, D))
    x = m @ np.swapaxes(m, -1, -2) + 0.1 * np.eye(D)
    y = nm.projection_from_spd_to_nested_spd(torch.from_numpy(x), proj_matrix := onest.grassmann_rand(rng, D, d))
    assert np.linalg.eigvalsh(y.numpy()).min() > 0
    np.testing.assert_allclose(y.numpy(), onest.projection_from_spd_to_nested_spd(x, proj_matrix).numpy(), rtol=0,
                               atol=1e-5 * np.abs(x).max())
