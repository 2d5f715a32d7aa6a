"""Same-box A/B timing of the reconstruction contraction kernel for several builds of the library:
    python scripts/dev_recon_ab.py lib1.so lib2.so ...   (each path is copied over gabotorch_b200/lib/libgabo_b200.so in a subprocess)"""
import os, shutil, subprocess, sys
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from gabotorch_b200 import ops, _lib
D, d, n = 20, 5, 65536
rng = np.random.default_rng(9)
q, _ = np.linalg.qr(rng.standard_normal((D, D)))
a = rng.standard_normal((D - d, D - d)); c = a @ a.T + np.eye(D - d)
k = rng.standard_normal((d, D - d)); k = 0.7 * k / np.linalg.norm(k, 2)
pack = ops.nested_spd_reconstruct_pack(torch.from_numpy(q[:, :d].copy()), torch.from_numpy(q[:, d:].copy()), torch.from_numpy(c), torch.from_numpy(k))
b = torch.randn(n, d, d, dtype=torch.float64, device='cuda')
y = b @ b.transpose(-1, -2) + torch.eye(d, dtype=torch.float64, device='cuda')
sq = ops.spd_sqrtm(y)
x = torch.empty(n, D, D, dtype=torch.float64, device='cuda')
lib = _lib.load()
def call():
    _lib.check(lib.gabo_nested_spd_reconstruct(ops._p(y), ops._p(sq), n, D, d, ops._p(pack), ops._p(x), _lib.stream_ptr()), 'recon')
for _ in range(5): call()
ts = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 20)
print('contraction only: best %%.4f ms, median %%.4f ms' %% (min(ts), sorted(ts)[2]))
''' % HERE
target = os.path.join(HERE, 'gabotorch_b200', 'lib', 'libgabo_b200.so')
keep = target + '.keep'
shutil.copy(target, keep)
try:
    for rnd in range(2):
        for path in sys.argv[1:]:
            shutil.copy(path, target)
            out = subprocess.run([sys.executable, '-c', CHILD], capture_output=True, text=True, timeout=300)
            print(os.path.basename(path), (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1], flush=True)
finally:
    shutil.move(keep, target)
