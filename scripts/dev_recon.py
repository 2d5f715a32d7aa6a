"""Reconstruction kernel timing (SPD(5) -> SPD(20), N = 65536) and a parity check against the fp64 composition in torch;
GABO_RECONSTRUCT_KERNEL=old selects the pre-DMMA kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gabotorch_b200 import nested_mappings as nm, ops
print('kernel:', os.environ.get('GABO_RECONSTRUCT_KERNEL', 'dmma (default)'))
for (D, d, n) in ((20, 5, 65536), (20, 5, 1000), (10, 3, 4097), (5, 2, 333), (12, 4, 100), (32, 5, 64)):
    rng = np.random.default_rng(9)
    q, _ = np.linalg.qr(rng.standard_normal((D, D)))
    a = rng.standard_normal((D - d, D - d)); c = a @ a.T + np.eye(D - d)
    k = rng.standard_normal((d, D - d)); k = 0.7 * k / np.linalg.norm(k, 2)
    w, v = torch.from_numpy(q[:, :d].copy()).cuda(), torch.from_numpy(q[:, d:].copy()).cuda()
    tc, tk = torch.from_numpy(c).cuda(), torch.from_numpy(k).cuda()
    rec = nm.NestedSpdReconstruction(w, v, tc, tk)
    b = torch.randn(n, d, d, dtype=torch.float64, device='cuda')
    y = b @ b.transpose(-1, -2) + torch.eye(d, dtype=torch.float64, device='cuda')
    x = rec(y)
    lam, vec, _ = ops.sym_eig(y); ys = (vec * lam.sqrt().unsqueeze(-2)) @ vec.transpose(-1, -2)
    lc, vc, _ = ops.sym_eig(tc); cs = (vc * lc.sqrt()) @ vc.T
    side = ys @ tk @ cs
    xr = torch.cat([torch.cat([y, side], -1), torch.cat([side.transpose(-1, -2), tc.expand(n, -1, -1)], -1)], -2)
    r = torch.cat([w, v], 1)
    ref = r @ xr @ r.T
    err = float((x.to(ref.device) - ref).abs().max() / ref.abs().max())
    print('SPD(%d)->SPD(%d) n=%d: max err / scale %.2e %s' % (d, D, n, err, 'OK' if err < 1e-12 else 'FAIL'), flush=True)
    if n == 65536:
        for _ in range(3): rec(y)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): rec(y)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nbytes = n * 8 * (2 * 2 * d * d + D * D)
        print('  N=65536: %.4f ms per call (sqrtm + contraction), %.0f GB/s' % (ms, nbytes / ms / 1e6))
