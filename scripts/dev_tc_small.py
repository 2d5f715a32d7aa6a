import sys
import numpy as np, torch
sys.path.insert(0, '.')
from gabotorch_b200 import ops, nested_mappings as nm
from oracle import nested as onest
D, d, n = 20, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(1)
xv = rng.standard_normal((n, D * (D + 1) // 2))
w = onest.grassmann_rand(rng, D, d)
P = onest.mandel_projection_matrix(w)
ref = xv.astype(np.float32).astype(np.float64) @ P.T
got = nm.projection_mandel(torch.from_numpy(xv), w).cpu().numpy()
print('max err / scale', np.abs(got - ref).max() / np.abs(ref).max())
