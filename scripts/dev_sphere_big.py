"""One launch each of sphere_gram_kernel<3,float> and <3,double> at N = 32768 (the launches that carry the HBM-fraction
claims) -- target of the ncu captures in scripts/gpu_r02a.sh.  Prints the event-timed GB/s as well."""
import ctypes, math, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gabotorch_b200 import _lib
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
for D in (3, 9):
    x = np.random.default_rng(1).standard_normal((N, D))
    x = torch.from_numpy(x / np.linalg.norm(x, axis=1, keepdims=True)).cuda()
    for dt, code in ((torch.float32, _lib.GABO_F32), (torch.float64, _lib.GABO_F64)):
        out = torch.empty(N, N, dtype=dt, device='cuda')
        def run():
            _lib.check(lib.gabo_sphere_gram(ctypes.c_void_p(x.data_ptr()), N, ctypes.c_void_p(x.data_ptr()), N, D,
                                            6.5 + math.log(2.0), 0, ctypes.c_void_p(out.data_ptr()), code, N,
                                            _lib.stream_ptr()), 'gabo_sphere_gram')
        for _ in range(2):
            run()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(5):
            run()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print('D=%d %s: %.3f ms, %.0f GB/s' % (D, dt, ms, N * N * out.element_size() / ms / 1e6))
        del out
