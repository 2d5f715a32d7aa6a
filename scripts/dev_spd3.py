"""Timing + all-pairs parity of the SPD(2) / SPD(3) / SPD(8) Gram kernels (device-resident, CUDA events)."""
import ctypes, math, sys
import numpy as np, torch
sys.path.insert(0, '.')
import bench
from gabotorch_b200 import _lib, ops
lib = _lib.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())
for d, N, beta in ((3, 2048, bench.BETA_SPD3), (3, 8192, bench.BETA_SPD3), (2, 2048, 1.0), (8, 2048, 0.22 + math.log(2.0))):
    v = bench.spd_sample_mandel(np.random.default_rng(1234), N, d)
    x = torch.from_numpy(v).cuda()
    fs = lib.gabo_spd_factor_stride(d)
    f1 = torch.empty(N, fs, dtype=torch.float64, device='cuda'); f2 = torch.empty_like(f1)
    flags = torch.zeros(1, dtype=torch.int32, device='cuda')
    out = torch.empty(N, N, dtype=torch.float64, device='cuda')
    _lib.check(lib.gabo_spd_factor2(p(x), N, p(x), N, d, 1, p(f1), p(f2), p(flags), _lib.stream_ptr()), 'factor')
    def run():
        _lib.check(lib.gabo_spd_ai_gram(p(f1), N, p(f2), N, d, beta, _lib.KIND_GAUSS, _lib.GABO_F32, 0, p(out), _lib.GABO_F64, N,
                                        _lib.stream_ptr()), 'gram')
    for _ in range(3): run()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print('SPD(%d) N=%d fp32: %.4f ms, %.3e pairs/s' % (d, N, ms, N * N / ms * 1e3))
    if N <= 2048:
        print('   parity', bench.spd_parity_report(v, beta, _lib.GABO_F32))
