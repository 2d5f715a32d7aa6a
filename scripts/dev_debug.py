import sys, math; sys.path.insert(0, '.')
import numpy as np, torch
import gabotorch_b200 as g
from gabotorch_b200 import ops, _lib, kernel_utils as ku
from oracle import sphere as osph
rng = np.random.default_rng(8)
D, d = 6, 3
x, y = osph.rand(rng, 21, D), osph.rand(rng, 17, D)
k = g.NestedSphereGaussianKernel(D, d, beta_min=1.0)
a = torch.tensor(x, requires_grad=True); b = torch.tensor(y, requires_grad=True)
q1 = ku._nested_sphere_project_autograd(a, k.axes, k.distances_to_axis)
q2 = ku._nested_sphere_project_autograd(b, k.axes, k.distances_to_axis)
print('q nan', q1.isnan().any().item(), q2.isnan().any().item(), q1.dtype, q1.is_contiguous(), q1.shape)
q1d = q1.detach().clone().requires_grad_(True); q2d = q2.detach().clone().requires_grad_(True)
dd = ku._SphereDistance.apply(q1d, q2d)
print('d nan', dd.isnan().any().item(), float(dd.min()), float(dd.max()))
gup = torch.randn_like(dd)
dd.backward(gup)
print('grad q nan', q1d.grad.isnan().any().item(), q2d.grad.isnan().any().item())
ref1 = torch.where((dd > 1e-7) & (dd < math.pi - 1e-7), -gup / torch.sin(dd), torch.zeros_like(dd)).detach() @ q2.detach()
print('vs matmul', float((q1d.grad - ref1).abs().max()))
q1.sum().backward(retain_graph=True)
print('chain grad nan', a.grad.isnan().any().item(), [p.grad.isnan().any().item() for p in k.axes])
