"""Wall-clock breakdown of one HD-GaBO SPD iteration (examples/hd_gabo_spd.py) by monkey-patched timers."""
import importlib.util, os, sys, time, collections
import torch
sys.path.insert(0, '.')
import gabotorch_b200 as g
acc = collections.OrderedDict()
def timed(name, fn):
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return out
    return wrap
g.fit_gpytorch_manifold = timed('fit_gpytorch_manifold', g.fit_gpytorch_manifold)
g.optimize_reconstruction_parameters_nested_spd = timed('reconstruction fit', g.optimize_reconstruction_parameters_nested_spd)
g.joint_optimize_manifold = timed('joint_optimize_manifold', g.joint_optimize_manifold)
spec = importlib.util.spec_from_file_location('ex', 'examples/hd_gabo_spd.py')
ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
t0 = time.perf_counter()
ex.run(n_iters=4, verbose=False)
tot = time.perf_counter() - t0
print('4 iterations: %.2f s' % tot)
for k, v in acc.items(): print('  %-28s %.2f s (%.0f %%)' % (k, v, 100 * v / tot))
log = getattr(g.fit_gpytorch_manifold, '__wrapped__', None)
