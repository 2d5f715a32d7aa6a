import sys, functools; sys.path.insert(0, '.')
import numpy as np, torch
import bench
import gabotorch_b200 as g
from gabotorch_b200 import ops, _lib, manifold_optimization as mo, riemannian_utils as ru
for d, R in ((3, 16), (5, 16), (8, 16)):
    rng = np.random.default_rng(31)
    xv = bench.spd_sample_mandel(rng, 32, d)
    y = np.array([float(np.sum(v * v)) for v in xv]); y = (y - y.mean()) / (y.std() + 1e-12)
    model = g.ManifoldGP(torch.from_numpy(xv), torch.from_numpy(y), g.ScaleKernel(g.SpdAffineInvariantGaussianKernel(beta_min=0.5)), noise=1e-2)
    model.covar_module.outputscale = 1.0
    gp = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False, compute='f64').device_gp()
    x0 = ops.mandel_unpack(torch.from_numpy(bench.spd_sample_mandel(np.random.default_rng(32), R, d, min_eig=0.5, max_eig=2.5)))
    out = ops.acq_ctr(gp, x0, [('max', 3.0)], maxiter=100, mingradnorm=1e-4)
    cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=3.0)]
    ref = mo.batched_trust_regions(gp, x0, maxiter=100, mingradnorm=1e-4, ineq_constraints=mo.batched_constraints(cons, _lib.SPD))
    print('d=%d kernel iters' % d, out[2].cpu().tolist())
    print('     lockstep    ', ref[2].cpu().tolist())
    print('     values k', [round(float(v), 6) for v in out[1].cpu()[:6]], 'l', [round(float(v), 6) for v in ref[1].cpu()[:6]])
    ei0 = ops.ei_eval(gp, x0)
    print('     EI at starts', [round(float(v), 6) for v in ei0.cpu()[:6]])
