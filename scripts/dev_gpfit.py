import sys, math; sys.path.insert(0, '.')
import numpy as np, torch
import bench
from gabotorch_b200 import ops, gp_fit
import gabotorch_b200 as g
rng = np.random.default_rng(11)
xt = bench.sphere_sample(rng, 32, 3); y = bench.ackley_sphere(xt)
dmat, _ = gp_fit.kernel_distance_matrix(g.SphereGaussianKernel(beta_min=6.5), torch.from_numpy(xt))
obj = gp_fit.MarginalLogLikelihood(dmat, y, 6.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))
raw0 = obj.inverse_transform((6.5 + math.log(2.0), math.log(2.0), 2.0, 0.0))
for _ in range(3):
    r = ops.gp_fit(dmat, torch.from_numpy(y), torch.from_numpy(raw0[None]), 6.5, 1e-8, [0, 0, 2.0, 0.15, 1.1, 0.05], [0, 0, 0, 0])
print(r)
