"""Batched manifold operations (through the C ABI) against the oracle and the golden outputs of the reference's numpy
formulas (Riemannian_utils/sphere_utils.py, spd_utils.py).  fp64 kernels: tolerance 1e-9."""
import numpy as np
import pytest

import gabotorch_b200 as g
from oracle import spd as ospd
from oracle import sphere as osph

pytestmark = pytest.mark.gpu


def test_sphere_ops_golden(golden):
    man = g.Sphere(6)
    x, y, v = golden['man_sphere_x'], golden['man_sphere_y'], golden['man_sphere_v']
    np.testing.assert_allclose(man.log(x, y), golden['man_sphere_log'], atol=1e-12)
    np.testing.assert_allclose(man.exp(x, 0.7 * golden['man_sphere_log']), golden['man_sphere_exp07'], atol=1e-12)
    np.testing.assert_allclose(man.dist(x, y), golden['man_sphere_dist'], atol=1e-12)
    np.testing.assert_allclose(man.parallel_transport(x, y, v), golden['man_sphere_pt'], atol=1e-12)


@pytest.mark.parametrize('D', [2, 3, 6, 17, 128])
def test_sphere_ops_vs_oracle(D):
    rng = np.random.default_rng(D)
    man = g.Sphere(D)
    x, y = osph.rand(rng, 200, D), osph.rand(rng, 200, D)
    h = rng.standard_normal((200, D))
    u = osph.proj(x, h)
    np.testing.assert_allclose(man.proj(x, h), u, atol=1e-13)
    np.testing.assert_allclose(man.egrad2rgrad(x, h), u, atol=1e-13)
    np.testing.assert_allclose(man.retr(x, 0.3 * u), osph.retr(x, 0.3 * u), atol=1e-13)
    np.testing.assert_allclose(man.exp(x, u), osph.exp(x, u), atol=1e-12)
    np.testing.assert_allclose(man.exp(x, 1e-5 * u), osph.exp(x, 1e-5 * u), atol=1e-13)     # tiny-step branch
    np.testing.assert_allclose(man.log(x, y), osph.log(x, y), atol=1e-12)
    np.testing.assert_allclose(man.transp(x, y, u), osph.transp(x, y, u), atol=1e-13)
    np.testing.assert_allclose(man.dist(x, y), osph.dist(x, y), atol=1e-12)
    np.testing.assert_allclose(man.exp(x, man.log(x, y)), y, atol=1e-11)                     # exp o log = id
    assert np.abs(np.sum(man.parallel_transport(x, y, u) * y, axis=-1)).max() < 1e-12       # lands in T_y
    # single point, numpy in -> numpy out (pymanopt calling convention)
    one = man.exp(x[0], u[0])
    assert isinstance(one, np.ndarray) and one.shape == (D,)
    np.testing.assert_allclose(man.inner(x, u, u), np.sum(u * u, axis=-1), rtol=1e-13)


@pytest.mark.parametrize('d', [3, 8])
def test_spd_ops_golden(golden, d):
    man = g.PositiveDefinite(d)
    tag = 'man_spd%d_' % d
    x, y, u = golden[tag + 'x'], golden[tag + 'y'], golden[tag + 'u']
    np.testing.assert_allclose(man.log(x, y), golden[tag + 'log'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(man.exp(x, 0.5 * golden[tag + 'log']), golden[tag + 'exp05'], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(man.dist(x, y), golden[tag + 'dist'], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(man.parallel_transport(x, y, u), golden[tag + 'pt'], rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize('d', [1, 2, 3, 4, 5, 6, 7, 8])
def test_spd_ops_vs_oracle(d):
    rng = np.random.default_rng(50 + d)
    man = g.PositiveDefinite(d)
    n = 64
    X, Y = ospd.spd_sample(rng, n, d, max_cond=100.0), ospd.spd_sample(rng, n, d, max_cond=100.0)
    G = rng.standard_normal((n, d, d))
    U = 0.5 * (G + np.swapaxes(G, -1, -2))
    V = ospd.egrad2rgrad(X, rng.standard_normal((n, d, d)))
    np.testing.assert_allclose(man.egrad2rgrad(X, G), ospd.egrad2rgrad(X, G), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(man.exp(X, U), ospd.exp(X, U), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(man.retr(X, 0.1 * U), ospd.retr(X, 0.1 * U), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(man.log(X, Y), ospd.log(X, Y), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(man.dist(X, Y), ospd.dist(X, Y), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(man.norm(X, U), ospd.norm(X, U), rtol=1e-10)
    np.testing.assert_allclose(man.inner(X, U, V), ospd.inner(X, U, V), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(man.transp(X, Y, U), U, atol=0)                              # identity transport
    np.testing.assert_allclose(man.parallel_transport(X, Y, U), ospd.parallel_transport(X, Y, U), rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(man.exp(X, man.log(X, Y)), Y, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(man.norm(X, man.log(X, Y)), man.dist(X, Y), rtol=1e-8)
    bad = X.copy()
    bad[0] = -bad[0]
    assert np.isnan(man.dist(bad, Y)[0]) and not np.isnan(man.dist(bad, Y)[1:]).any()
