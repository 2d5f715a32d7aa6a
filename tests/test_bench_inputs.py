"""bench.py synthesises its inputs without the oracle (the oracle may only be the checker); these CPU tests pin the
synthetic-input helpers on the oracle's restatement of the reference's sampling law and test functions."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

import bench
from oracle import spd as ospd
from oracle import sphere as osph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spd_sample_mandel_follows_the_reference_law():
    rng = np.random.default_rng(5)
    for d in (2, 3, 5, 8):
        v = bench.spd_sample_mandel(rng, 200, d)
        assert v.shape == (200, d * (d + 1) // 2) and v.dtype == np.float64
        m = ospd.vector_to_symmetric_matrix_mandel(torch.from_numpy(v)).numpy()          # spd_utils_torch.py:159-194
        np.testing.assert_allclose(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(m)).numpy(), v, atol=1e-14)
        lam = np.linalg.eigvalsh(m)
        assert lam.min() >= 0.001 - 1e-12 and lam.max() <= 5.0 + 1e-12                    # spd_utils.py:298-305
        assert (lam.max(1) / lam.min(1)).max() <= 100.0 * (1 + 1e-9)                      # cond filter of the examples


def test_ackley_sphere_matches_the_oracle_restatement():
    rng = np.random.default_rng(6)
    x = bench.sphere_sample(rng, 64, 6)
    np.testing.assert_allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(bench.ackley_sphere(x), osph.ackley(x), rtol=1e-12, atol=1e-12)


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1', '--ref-rows', '1'], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC and d['unit'] == bench.UNIT
    assert d['cpu_baseline']['kind'] == 'port' and d['e2e']['h2d_bytes_per_step'] == 0 and d['value'] > 0
