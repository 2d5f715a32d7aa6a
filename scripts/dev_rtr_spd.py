"""Timing of the lock-step trust-region driver on SPD(d) (wall clock between synchronisations; launch-bound path)."""
import math, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from gabotorch_b200 import _lib, ops, manifold_optimization as mo
from oracle import gp as ogp, spd as ospd
for d, R in ((3, 512), (8, 128)):
    rng = np.random.default_rng(5)
    xt = ospd.spd_sample(rng, 32, d, max_cond=100.0)
    y = ospd.ackley(ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(xt)))
    gp = ogp.make_gp('spd', xt, y, beta=0.5 + math.log(2.0), noise=1e-2)
    dgp = ops.DeviceGP(_lib.SPD, d, ops.spd_factor(gp.x_train, d, False), gp.alpha, gp.minv, gp.mean, gp.outputscale,
                       gp.beta, gp.best_f, gp.kxx, _lib.GABO_F32)
    x0 = ospd.spd_sample(rng, R, d, max_cond=50.0)
    mo.batched_trust_regions(dgp, x0, maxiter=3)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    X, val, it, why = mo.batched_trust_regions(dgp, x0)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('SPD(%d) R=%d fp32: %.1f ms, mean outer iters %.1f, max %d, reasons %s, solves/s %.3g'
          % (d, R, dt * 1e3, it.double().mean().item(), int(it.max()), np.unique(why.cpu().numpy()).tolist(), R / dt))
