"""tcgen05 projection kernel: parity against the fp64 operator product, then timing at N = 2^20 (GABO_PROJECT_KERNEL=mma
=tc in the environment selects the tcgen05 kernel instead of the default mma.sync one)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from gabotorch_b200 import ops, nested_mappings as nm
from oracle import nested as onest
print('kernel:', os.environ.get('GABO_PROJECT_KERNEL', 'mma.sync (default)'))
for D, d, n in ((20, 5, 64), (20, 5, 65), (20, 5, 128), (20, 5, 30000), (5, 2, 1001), (6, 3, 777), (3, 1, 500)):
    rng = np.random.default_rng(D * 100 + d)
    dvh = D * (D + 1) // 2
    xv = rng.standard_normal((n, dvh))
    w = onest.grassmann_rand(rng, D, d)
    P = onest.mandel_projection_matrix(w)
    ref = xv.astype(np.float32).astype(np.float64) @ P.T
    got = nm.projection_mandel(torch.from_numpy(xv), w).cpu().numpy()
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print('SPD(%d)->SPD(%d) n=%d: max err / scale %.3e %s' % (D, d, n, err, 'OK' if err <= 3e-6 else 'FAIL'), flush=True)
    if err > 3e-6:
        bad = np.argwhere(np.abs(got - ref) > 3e-6 * np.abs(ref).max())
        print('  first bad entries', bad[:8].tolist(), got[tuple(bad[0])], ref[tuple(bad[0])])
if '--parity-only' in sys.argv:
    sys.exit(0)
N, D, d = 1 << 20, 20, 5
x = torch.randn(N, 210, device='cuda')
pack = ops.nested_projection_matrix(torch.from_numpy(onest.grassmann_rand(np.random.default_rng(0), D, d)))
for _ in range(3):
    y = ops.nested_spd_project(x, D, d, pack)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for _ in range(10):
    y = ops.nested_spd_project(x, D, d, pack)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print('N=2^20: %.4f ms, %.0f GB/s' % (ms, N * 4 * 225 / ms / 1e6))
