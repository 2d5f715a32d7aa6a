"""Golden values of the nested SPD reconstruction costs FROM THE REFERENCE'S OWN CODE (build container only):

    python tests/golden/make_golden_recon_cost.py

Records inputs and outputs of ``min_affine_invariant_distance_reconstruction_cost`` and
``min_log_euclidean_distance_reconstruction_cost`` (nested_mappings/nested_spd_optimization.py:22-92) -> recon_cost_vectors.npz.
The reference accumulates the per-sample distances in a float32 tensor (``cost = torch.zeros(n_data)``, :49, :86): the
values carry float32 rounding.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_loader  # noqa: E402
from make_golden import spd_points  # noqa: E402


def main():
    mod = reference_loader.load_nested_spd_optimization()
    proj = reference_loader.load().nested_spd_utils.projection_from_spd_to_nested_spd
    rng = np.random.default_rng(4321)
    out = {}
    for name, n, D, d in (('rc_6_2', 9, 6, 2), ('rc_10_3', 7, 10, 3), ('rc_20_5', 5, 20, 5)):
        q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        w, v = q[:, :d].copy(), q[:, d:].copy()
        c = spd_points(rng, 1, D - d, max_cond=50.0)[0]
        k = rng.standard_normal((d, D - d))
        k = 0.6 * k / np.linalg.norm(k)
        x = spd_points(rng, n, D, max_cond=1e9)
        xt, wt = torch.from_numpy(x), torch.from_numpy(w)
        y = proj(xt, wt).double()
        args = (xt, y, wt, torch.from_numpy(v), torch.from_numpy(c), torch.from_numpy(k))
        out[name + '_x'], out[name + '_y'], out[name + '_w'] = x, y.numpy(), w
        out[name + '_v'], out[name + '_c'], out[name + '_k'] = v, c, k
        out[name + '_ai'] = np.float64(mod.min_affine_invariant_distance_reconstruction_cost(*args).item())
        out[name + '_le'] = np.float64(mod.min_log_euclidean_distance_reconstruction_cost(*args).item())
        print(name, out[name + '_ai'], out[name + '_le'])
    path = os.path.join(HERE, 'recon_cost_vectors.npz')
    np.savez_compressed(path, **out)
    print('wrote %s: %d arrays, %.1f KiB' % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
