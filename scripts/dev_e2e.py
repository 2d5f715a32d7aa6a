"""Developer check (GPU): host-to-host Gram through the API, device-out + pinned copy vs zero-copy mapped output."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gabotorch_b200 import ops, _lib
from gabotorch_b200.kernel_utils import _finish

rng = np.random.default_rng(1234)
res = {}
for n in (2048, 8192):
    xm = bench.spd_sample_mandel(rng, n, 3)
    x1 = torch.from_numpy(xm).pin_memory(); x2 = torch.from_numpy(xm.copy()).pin_memory()
    def dev_copy():
        return _finish(ops.spd_ai_gram(x1, x2, 1.19), x1)
    def pageable():
        return ops.spd_ai_gram(x1, x2, 1.19).cpu()
    def zero_copy():
        k = ops.spd_ai_gram(x1, x2, 1.19, host_out=True)
        torch.cuda.synchronize()
        return k
    outs = {}
    for name, fn in (('pageable', pageable), ('pinned_copy', dev_copy), ('zero_copy', zero_copy)):
        for _ in range(3): outs[name] = fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): outs[name] = fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
        res['n%d_%s' % (n, name)] = dict(ms=dt * 1e3, pairs_per_s=n * n / dt, GBs=n * n * 8 / dt / 1e9)
        print(n, name, '%.3f ms  %.3e pairs/s  %.1f GB/s' % (dt * 1e3, n * n / dt, n * n * 8 / dt / 1e9), flush=True)
    assert torch.equal(outs['pageable'], outs['zero_copy']) and torch.equal(outs['pageable'], outs['pinned_copy'])
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/dev_e2e.json', 'w'), indent=1)
