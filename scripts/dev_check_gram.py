"""Developer check (GPU): sphere / SPD Gram vs the oracle + first timings.  Not part of the test suite."""
import ctypes, json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gabotorch_b200 import _lib
from oracle import sphere as osph, spd as ospd

lib = _lib.load()
dev = torch.device('cuda:0')
P = lambda t: ctypes.c_void_p(t.data_ptr())
S = _lib.stream_ptr

def sphere_gram(x1, x2, beta, kind=0, out_dtype=torch.float32):
    out = torch.empty(x1.shape[0], x2.shape[0], dtype=out_dtype, device=dev)
    _lib.check(lib.gabo_sphere_gram(P(x1), x1.shape[0], P(x2), x2.shape[0], x1.shape[1], beta, kind, P(out),
                                    0 if out_dtype == torch.float32 else 1, out.stride(0), S()), 'sphere_gram')
    return out

def spd_factor(xm, d):
    fs = lib.gabo_spd_factor_stride(d)
    fac = torch.empty(xm.shape[0], fs, dtype=torch.float64, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.gabo_spd_factor(P(xm), xm.shape[0], d, 1, P(fac), P(flags), S()), 'factor')
    return fac, flags

def spd_gram(f1, f2, d, beta, kind=0, compute=0, symmetric=0, out_dtype=torch.float32):
    out = torch.empty(f1.shape[0], f2.shape[0], dtype=out_dtype, device=dev)
    _lib.check(lib.gabo_spd_ai_gram(P(f1), f1.shape[0], P(f2), f2.shape[0], d, beta, kind, compute, symmetric, P(out),
                                    0 if out_dtype == torch.float32 else 1, out.stride(0), S()), 'spd_gram')
    return out

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

res = {}
rng = np.random.default_rng(1234)
# ---- sphere parity
for D, N, beta in ((3, 300, 6.5 + np.log(2)), (6, 257, 1.0 + np.log(2)), (9, 130, 0.6 + np.log(2)), (20, 70, 0.35 + np.log(2))):
    x = osph.rand(rng, N, D); y = osph.rand(rng, N + 13, D)
    y[0] = x[0]; y[1] = -x[1]
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    for kind, name in ((0, 'gauss'), (2, 'dist'), (1, 'laplace')):
        got = sphere_gram(xt, yt, beta, kind).double().cpu()
        if kind == 0: ref = osph.sphere_gaussian_kernel(x, y, beta)
        elif kind == 2: ref = osph.sphere_distance(x, y)
        else: ref = osph.sphere_laplace_kernel(x, y, 1.0 / np.sqrt(beta))
        if kind == 2:
            err = ((got - ref).abs() / (1e-5 * ref + 1e-6)).max().item()
        else:
            m = ref >= 1e-6
            err = ((got - ref).abs() / ref)[m].max().item()
        res['sphere_D%d_%s' % (D, name)] = err
        print('sphere D=%d %s: err metric %.3e' % (D, name, err), flush=True)
# ---- sphere timing
for D, N in ((3, 2048), (3, 16384), (3, 32768), (6, 32768), (9, 32768)):
    x = torch.from_numpy(osph.rand(rng, N, D)).to(dev)
    out = torch.empty(N, N, dtype=torch.float32, device=dev)
    def f():
        _lib.check(lib.gabo_sphere_gram(P(x), N, P(x), N, D, 7.19, 0, P(out), 0, N, S()))
    med, best = timeit(f)
    gbs = N * N * 4 / (best * 1e-3) / 1e9
    res['sphere_time_D%d_N%d' % (D, N)] = dict(ms_med=med, ms_best=best, pairs_per_s=N * N / (best * 1e-3), GBs=gbs)
    print('sphere D=%d N=%d: %.3f ms (best %.3f) -> %.3e pairs/s, %.0f GB/s' % (D, N, med, best, N * N / (best * 1e-3), gbs), flush=True)
# ---- SPD parity
for d, N, beta in ((3, 200, 0.5 + np.log(2)), (2, 100, 0.6 + np.log(2)), (5, 96, 0.25 + np.log(2)), (8, 80, 0.22 + np.log(2))):
    X = ospd.spd_sample(rng, N, d, max_cond=100.0)
    xm = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(X))
    ref_d = ospd.affine_invariant_distance(torch.from_numpy(X), torch.from_numpy(X))
    ref_k = torch.exp(-ref_d * ref_d * beta)
    fac, flags = spd_factor(xm.to(dev), d)
    for compute in (0, 1):
        for sym in (0, 1):
            gd = spd_gram(fac, fac, d, beta, 2, compute, sym).double().cpu()
            gk = spd_gram(fac, fac, d, beta, 0, compute, sym).double().cpu()
            ed = ((gd - ref_d).abs() / (1e-5 * ref_d + 1e-6)).max().item()
            m = ref_k >= 1e-6
            ek = ((gk - ref_k).abs() / ref_k)[m].max().item()
            res['spd_d%d_c%d_s%d' % (d, compute, sym)] = dict(dist_crit=ed, k_rel=ek)
            print('spd d=%d compute=%d sym=%d: dist crit %.3e  K rel %.3e flags %d' % (d, compute, sym, ed, ek, flags.item()), flush=True)
# ---- SPD timing
for d, N in ((3, 2048), (3, 8192), (3, 16384), (5, 8192), (8, 2048), (8, 8192)):
    X = ospd.spd_sample(rng, N, d, max_cond=100.0)
    xm = ospd.symmetric_matrix_to_vector_mandel(torch.from_numpy(X)).to(dev)
    fs = lib.gabo_spd_factor_stride(d)
    fac = torch.empty(N, fs, dtype=torch.float64, device=dev); flags = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty(N, N, dtype=torch.float32, device=dev)
    for compute in (0, 1):
        for sym in (0, 1):
            def f():
                _lib.check(lib.gabo_spd_factor(P(xm), N, d, 1, P(fac), P(flags), S()))
                _lib.check(lib.gabo_spd_ai_gram(P(fac), N, P(fac), N, d, 1.19, 0, compute, sym, P(out), 0, N, S()))
            med, best = timeit(f, n=5, warm=2)
            res['spd_time_d%d_N%d_c%d_s%d' % (d, N, compute, sym)] = dict(ms_med=med, ms_best=best, pairs_per_s=N * N / (best * 1e-3))
            print('spd d=%d N=%d compute=%d sym=%d: %.3f ms (best %.3f) -> %.3e pairs/s' % (d, N, compute, sym, med, best, N * N / (best * 1e-3)), flush=True)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/dev_check_gram.json', 'w'), indent=1)
