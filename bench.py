#!/usr/bin/env python
"""bench.py -- the headline measurement of the GaBOtorch hot path on B200 (contract: see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]

Headline workload (BASELINE.json configs[1]): the SpdAffineInvariantGaussianKernel Gram matrix on SPD(3), N = 2048 points
given as Mandel vectors, K = exp(-beta d_AI^2), fp64 output like the reference (kernels_spd.py:72-100).
One "step" = one full Gram build = per-point factorisation of both operands (gabo_spd_factor2, one launch) + the per-pair kernel
(gabo_spd_ai_gram, all N x N pairs -- the symmetric shortcut of the product API is NOT used for the headline value).

  value        pairs/s, inputs resident in HBM, CUDA-event time of the K steps (L2 flushed between steps), max over ranks
  e2e          the same metric through the reference-facing API SpdAffineInvariantGaussianKernel.forward(x1, x2) with
               pinned HOST tensors in and a HOST float64 tensor out (H2D + D2H inside the timed region)
  roofline     the dominant kernel (spd_ai_gram_kernel): algorithmic bytes per launch / event time of that launch alone
  cpu_baseline the oracle port of the reference's own per-pair loop (spd_utils_torch.py:92-110) on a bounded row sample
  extra        the other BASELINE configs measured the same way (sphere Gram, SPD(8) Gram, acquisition optimiser
               candidates/s with the NCCL all-gather of the argmax records, nested projection)

Multi-GPU (torchrun, one rank per GPU): ONE Gram K(X1, X2) with X1 = 2048 G points and X2 = 2048 points is
built in ROW BLOCKS -- rank g owns rows [2048 g, 2048 (g+1)) of X1, X2 is replicated (48 KB), no data-path collective;
per-GPU work is fixed (weak scaling) and equals configs[1].  The extras carry the other sharded paths: a fixed
8192 x 8192 SPD(3) Gram and a fixed 32768 x 32768 sphere Gram split in row blocks (strong scaling, with and without the
all-gather of the blocks), BASELINE configs[3] exactly (SPD(8), 4096 restarts in total sharded by `shard_range`, ONE
all-gather of (value, global index, candidate) records, winner ASSERTED bit-identical to the single-GPU winner of the
same starts) and configs[4] (projection, N sharded).

`--impl reference` times the reference's CPU algorithm (oracle port; the reference is pure Python whose kernel classes
need gpytorch, absent from this image, so it cannot be pip-installed -- DESIGN.md) on the host cores, rank 0 only.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'geodesic_kernel_pairs_per_s'
UNIT = 'pairs/s'
WORKLOAD = 'SpdAffineInvariantGaussianKernel Gram, SPD(3), N=2048 (BASELINE configs[1])'
N_POINTS, SPD_D = 2048, 3
BETA_SPD3 = 0.5 + math.log(2.0)          # beta_min(d=3) = 0.5 (gabo_spd.py:151-162) + softplus(0)


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs.  Not the oracle: plain numpy sampling of the reference's input laws.
# ----------------------------------------------------------------------------------------------------------------

def spd_sample_mandel(rng, n, d, min_eig=0.001, max_eig=5.0, max_cond=100.0):
    """Law of spd_sample (Riemannian_utils/spd_utils.py:290-306) with the cond <= 100 filter of
    spd_gaussian_kernel_parameters.py:91-96, returned as Mandel vectors (n, d(d+1)/2) fp64."""
    mats = np.empty((n, d, d))
    k = 0
    while k < n:
        m = min(4 * (n - k) + 16, 1 << 16)
        lam = min_eig + (max_eig - min_eig) * rng.random((m, d))
        keep = lam.max(1) / lam.min(1) <= max_cond
        lam = lam[keep][:n - k]
        q, _ = np.linalg.qr(rng.standard_normal((lam.shape[0], d, d)))
        mats[k:k + lam.shape[0]] = (q * lam[:, None, :]) @ np.swapaxes(q, -1, -2)
        k += lam.shape[0]
    mats = 0.5 * (mats + np.swapaxes(mats, -1, -2))
    cols = []
    for off in range(d):
        diag = np.stack([mats[:, i, i + off] for i in range(d - off)], axis=1)
        cols.append(diag if off == 0 else diag * math.sqrt(2.0))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def sphere_sample(rng, n, D):
    x = rng.standard_normal((n, D))
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def ackley_sphere(x):
    """Ackley on S^{D-1} in the tangent space at e_1 (test_functions_sphere.py:34-65), only used to synthesise GP targets."""
    D = x.shape[1]
    base = np.zeros(D)
    base[0] = 1.0
    c = np.clip(x @ base, -1.0, 1.0)
    th = np.arccos(c)
    p = x - c[:, None] * base
    pn = np.linalg.norm(p, axis=1, keepdims=True)
    u = np.where(pn > 1e-12, p / np.maximum(pn, 1e-300) * th[:, None], 0.0)[:, 1:]
    a, b, cc = 20.0, 0.2, 2.0 * np.pi
    dd = u.shape[1]
    return (-a * np.exp(-b * np.sqrt((u ** 2).sum(1) / dd)) - np.exp(np.cos(cc * u).sum(1) / dd) + a + np.e)


# ----------------------------------------------------------------------------------------------------------------
# clocks: NVML sampler thread running DURING the timed regions
# ----------------------------------------------------------------------------------------------------------------

class ClockSampler:
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap',
               0x80: 'hw_power_brake', 0x2: 'applications_clocks_setting'}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                uuid = 'GPU-' + str(torch.cuda.get_device_properties(device_index).uuid)
                self._h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: the clocks object says so instead of inventing numbers
            self.error = repr(e)

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    try:
                        bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.REASONS.items():
                        if bits & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def timed(self, on):
        (self._active.set if on else self._active.clear)()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)

    def summary(self):
        if self._h is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvml_unavailable']}
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------------

def cpu_reference_step(x1_mandel, x2_mandel, beta):
    """One pass of the reference's CPU algorithm (oracle port of kernels_spd.py:90-100 with the per-pair eigen-solve
    loop of spd_utils_torch.py:92-110) over len(x1) x len(x2) pairs.  Returns seconds."""
    import torch
    from oracle import spd as ospd
    a, b = torch.from_numpy(x1_mandel), torch.from_numpy(x2_mandel)
    t0 = time.perf_counter()
    k = ospd.spd_affine_invariant_gaussian_kernel(a, b, beta, loop=True)
    dt = time.perf_counter() - t0
    assert k.shape == (a.shape[0], b.shape[0])
    return dt


def cpu_vectorised_step(x1_mandel, x2_mandel, beta):
    """'Fair CPU': the same arithmetic with batched cholesky / eigvalsh on all cores (no Python per-pair loop)."""
    import torch
    from oracle import spd as ospd
    a, b = torch.from_numpy(x1_mandel), torch.from_numpy(x2_mandel)
    t0 = time.perf_counter()
    ospd.spd_affine_invariant_gaussian_kernel(a, b, beta, loop=False)
    return time.perf_counter() - t0


def spd_parity_report(x_mandel, beta, compute=0, rows_per_chunk=256):
    """Checker (oracle used as the CHECKER, never as the thing measured): the GPU Gram of ALL pairs of `x_mandel` against
    the vectorised oracle restatement of spd_utils_torch.py:87-120 + kernels_spd.py:96-98.  Returns the achieved
      max_rel_err_d            max |d - d_ref| / d_ref over pairs with d_ref >= 1e-3
      max_dist_margin          max (|d - d_ref| - (1e-5 d_ref + 1e-6))  -- <= 0 means the SURVEY 8(d) distance bound holds
      max_rel_err_K            max |K - K_ref| / K_ref over entries with K_ref >= 1e-6   (the flat north_star figure)
      max_rel_err_K_amplified  max of the same error divided by 1e-5 max(1, 2 beta d^2)  (<= 1: the bound the distance
                               tolerance implies through exp; the reference's own float32 eigenvalues have this
                               amplification too).
    `compute` may be a list of arithmetic types: one report per type, the oracle evaluated once."""
    import torch
    from gabotorch_b200 import _lib, ops
    from oracle import spd as ospd
    v = torch.as_tensor(x_mandel)
    n = v.shape[0]
    modes = list(compute) if isinstance(compute, (list, tuple)) else [compute]
    got = [(ops.spd_ai_gram(v, v.clone(), kind=_lib.KIND_DIST, compute=c).cpu(),
            ops.spd_ai_gram(v, v.clone(), beta, _lib.KIND_GAUSS, compute=c).cpu()) for c in modes]
    mats = ospd.vector_to_symmetric_matrix_mandel(v)
    reps = [{'pairs_checked': 0, 'max_rel_err_d': 0.0, 'max_dist_margin': -1.0, 'max_rel_err_K': 0.0,
             'max_rel_err_K_amplified': 0.0} for _ in modes]
    for lo in range(0, n, rows_per_chunk):
        dref = ospd.affine_invariant_distance(mats[lo:lo + rows_per_chunk], mats)
        kref = torch.exp(-dref * dref * beta)
        big, m = dref >= 1e-3, kref >= 1e-6
        for rep, (dg, kg) in zip(reps, got):
            d, k = dg[lo:lo + rows_per_chunk], kg[lo:lo + rows_per_chunk]
            ed = (d - dref).abs()
            rep['max_dist_margin'] = max(rep['max_dist_margin'], float((ed - (1e-5 * dref + 1e-6)).max()))
            if bool(big.any()):
                rep['max_rel_err_d'] = max(rep['max_rel_err_d'], float((ed[big] / dref[big]).max()))
            if bool(m.any()):
                rel = (k - kref).abs()[m] / kref[m]
                rep['max_rel_err_K'] = max(rep['max_rel_err_K'], float(rel.max()))
                amp = 1e-5 * torch.clamp(2.0 * beta * dref[m] ** 2, min=1.0)
                rep['max_rel_err_K_amplified'] = max(rep['max_rel_err_K_amplified'], float((rel / amp).max()))
            rep['pairs_checked'] += int(dref.numel())
    return reps if isinstance(compute, (list, tuple)) else reps[0]


def cpu_acq_baseline(kind, gp_inputs, x0, T, budget_s=8.0, max_restarts=256):
    """CPU baseline of the acquisition optimiser: the oracle's per-restart Riemannian CG
    (pymanopt 0.2.x ConjugateGradient restated, parity UNPINNED) run SERIALLY over a subset of the restarts as the
    reference does (manifold_optimize.py:207-221), exactly T iterations each; the rate is extrapolated linearly.
    The cost here is the oracle's closed-form numpy EI -- the reference additionally pays a gpytorch posterior and a
    torch.autograd graph per cost call, so this number is GENEROUS to the reference."""
    from oracle import gp as ogp, rcg as orcg
    xt, y, beta, noise = gp_inputs
    gp = ogp.make_gp(kind, xt, y, beta=beta, outputscale=1.0, noise=noise)
    opts = orcg.CGOptions(maxiter=T + 1, mingradnorm=0.0, minstepsize=-1.0)
    done, iters, t0 = 0, 0, time.perf_counter()
    with np.errstate(all='ignore'):
        for x in x0[:max_restarts]:
            _, _, it, _ = orcg.solve_cg(gp, x, opts)
            done += 1
            iters += it
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return {'value': iters / dt, 'unit': 'restart-iterations/s', 'cores': 1, 'kind': 'port',
            'sample': '%d of the restarts x %d CG iterations solved serially by oracle.rcg.solve_cg in %.1f s, '
                      'extrapolated linearly; closed-form numpy EI (no autograd / gpytorch posterior): generous to '
                      'the reference; pymanopt semantics restated, unpinned' % (done, T, dt)}


def run_reference_arm(args):
    """--impl reference: rank 0 times the reference CPU algorithm on a bounded sample of the headline workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rng = np.random.default_rng(1234)
    x = spd_sample_mandel(rng, N_POINTS, SPD_D)
    rows = args.ref_rows
    # bound the whole run to ~2 minutes whatever K the driver asks for: one probe row gives the per-pair cost
    per_pair = cpu_reference_step(x[:1], x, BETA_SPD3) / N_POINTS
    budget_rows = int(120.0 / max(per_pair * N_POINTS * (args.warmup + args.steps), 1e-9))
    rows = max(1, min(rows, budget_rows))
    times = []
    for s in range(args.warmup + args.steps):
        lo = (s * rows) % N_POINTS
        dt = cpu_reference_step(x[lo:lo + rows], x, BETA_SPD3)
        if s >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    value = rows * N_POINTS * len(times) / total
    sample = '%d rows x %d columns of the N=%d Gram per step (%d pairs); per-pair Python loop as spd_utils_torch.py:109' \
             % (rows, N_POINTS, N_POINTS, rows * N_POINTS)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'n_points': N_POINTS, 'spd_dim': SPD_D, 'beta': BETA_SPD3,
                   'note': 'reference is pure Python (not pip-installable: needs gpytorch/botorch/pymanopt, absent); '
                           'timed through the oracle port of its algorithm on the host cores'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------

class Bench:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args = args
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: gabotorch_b200 has no CPU fallback '
                             '(use --impl reference for the CPU arm)')
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=self.dev)
        from gabotorch_b200 import _lib, ops
        self._lib, self.ops = _lib, ops
        self.lib = _lib.load()
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # 256 MB > 126 MB L2
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                self.peaks = json.load(f)
        except Exception:
            pass
        self.hbm_peak = float(self.peaks.get('hbm_gbs', 6650.0))
        self.peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in self.peaks else 'fallback (B200_PROFILING.md)'
        self.clocks = ClockSampler(self.local_rank)
        self.clocks.start()

    # -- helpers ---------------------------------------------------------------------------------------------
    def ncu_traffic(self, kernel):
        """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu capture of the same
        command (profiles/rNN_traffic.json, written by scripts/make_profiles.py); None when there is no capture.
        For the N = 2048 Gram it is far BELOW the algorithmic bytes: the 33.5 MB result stays in the 126 MB L2."""
        try:
            files = sorted(f for f in os.listdir(os.path.join(ROOT, 'profiles')) if f.endswith('_traffic.json'))
            with open(os.path.join(ROOT, 'profiles', files[-1])) as f:
                t = json.load(f)[kernel]
            return t['dram_bytes_read'] + t['dram_bytes_write']
        except Exception:
            return None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        self.flush_buf.fill_(1)

    def time_steps(self, fn, steps, warmup, flush=True):
        """Device time of `steps` calls of fn (CUDA events on the launching stream, L2 flushed between steps, the flush
        outside the event pairs).  Returns total milliseconds, max over ranks."""
        torch = self.torch
        for _ in range(max(warmup, 0)):
            fn()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        self.clocks.timed(True)
        for a, b in evs:
            if flush:
                self.flush_l2()
            a.record()
            fn()
            b.record()
        self.barrier()
        self.clocks.timed(False)
        total = sum(a.elapsed_time(b) for a, b in evs)
        return self.max_over_ranks(total)

    def p(self, t):
        import ctypes
        return ctypes.c_void_p(t.data_ptr())

    # -- headline: SPD(3) Gram ------------------------------------------------------------------------------
    def spd_gram_setup(self, n, d, seed, sharded=False):
        """Operands of one Gram build.  `sharded`: the GLOBAL point set has n * world points (same seed on every rank);
        this rank owns rows [n rank, n (rank+1)) as x1, x2 = the first n points, replicated (SURVEY 8e row-block
        partition; on one GPU x1 and x2 hold the same n points, which is BASELINE configs[1])."""
        torch = self.torch
        rng = np.random.default_rng(seed)
        if sharded:
            allx = spd_sample_mandel(rng, n * self.world, d)
            xm, x2m = np.ascontiguousarray(allx[n * self.rank:n * (self.rank + 1)]), np.ascontiguousarray(allx[:n])
        else:
            xm = spd_sample_mandel(rng, n, d)
            x2m = xm.copy()
        fs = self.lib.gabo_spd_factor_stride(d)
        st = {
            'n': n, 'd': d, 'x_host': xm, 'x2_host': x2m,
            'x1': torch.from_numpy(xm).to(self.dev), 'x2': torch.from_numpy(x2m).to(self.dev),
            'fac1': torch.empty(n, fs, dtype=torch.float64, device=self.dev),
            'fac2': torch.empty(n, fs, dtype=torch.float64, device=self.dev),
            'flags': torch.zeros(1, dtype=torch.int32, device=self.dev),
            'out': torch.empty(n, n, dtype=torch.float64, device=self.dev),
        }
        return st

    def spd_gram_step(self, st, beta, symmetric=False):
        lib, lb, p, s = self.lib, self._lib, self.p, self._lib.stream_ptr()
        n, d = st['n'], st['d']
        if symmetric:
            f2 = st['fac1']
            lb.check(lib.gabo_spd_factor(p(st['x1']), n, d, 1, p(st['fac1']), p(st['flags']), s), 'gabo_spd_factor')
        else:
            f2 = st['fac2']
            lb.check(lib.gabo_spd_factor2(p(st['x1']), n, p(st['x2']), n, d, 1, p(st['fac1']), p(f2), p(st['flags']), s),
                     'gabo_spd_factor2')
        lb.check(lib.gabo_spd_ai_gram(p(st['fac1']), n, p(f2), n, d, beta, lb.KIND_GAUSS, lb.GABO_F32,
                                      1 if symmetric else 0, p(st['out']), lb.GABO_F64, n, s), 'gabo_spd_ai_gram')

    def spd_gram_only(self, st, beta):
        lib, lb, p, s = self.lib, self._lib, self.p, self._lib.stream_ptr()
        n, d = st['n'], st['d']
        lb.check(lib.gabo_spd_ai_gram(p(st['fac1']), n, p(st['fac2']), n, d, beta, lb.KIND_GAUSS, lb.GABO_F32, 0,
                                      p(st['out']), lb.GABO_F64, n, s), 'gabo_spd_ai_gram')

    def headline(self):
        torch, args = self.torch, self.args
        st = self.spd_gram_setup(N_POINTS, SPD_D, 1234, sharded=True)
        pairs = N_POINTS * N_POINTS
        total_ms = self.time_steps(lambda: self.spd_gram_step(st, BETA_SPD3), args.steps, args.warmup)
        launches = 2 * args.steps                              # spd_factor_kernel (both operands) + spd_ai_gram_kernel per step
        assert int(st['flags'].item()) == 0
        ms_per_step = total_ms / args.steps
        value = self.world * pairs / (ms_per_step * 1e-3)

        # roofline of the dominant kernel alone (fp64 Gram out + the factor records in)
        nroof = max(10, min(args.steps, 200))
        roof_ms = self.time_steps(lambda: self.spd_gram_only(st, BETA_SPD3), nroof, 3) / nroof
        fs = self.lib.gabo_spd_factor_stride(SPD_D)
        alg_bytes = pairs * 8 + 2 * N_POINTS * fs * 8
        achieved = alg_bytes / (roof_ms * 1e-3) / 1e9
        roofline = {'kernel': 'spd_ai_gram_kernel<3,float,double,GAUSS>', 'bound': 'hbm', 'achieved': achieved,
                    'peak': self.hbm_peak, 'unit': 'GB/s', 'frac': achieved / self.hbm_peak,
                    'traffic': self.ncu_traffic('spd_ai_gram_kernel<3, float, double, 0>'),
                    'peak_source': self.peak_src, 'ms_per_launch': roof_ms, 'algorithmic_bytes_per_launch': alg_bytes,
                    'note': 'the per-pair eigenvalue solve is instruction-bound (about 115 FLOP per 8 output bytes): '
                            'see compute_roofline; the HBM fraction is reported because the contract asks for it'}
        # Executed arithmetic of the closed-form SPD(3) solve (ncu opcode mix of the N = 8192 launch, profiles/r02*):
        # per pair 28 FFMA + 21 FMUL + 5 FADD (packed two pairs per instruction) + 11 DFMA/DMUL + 8 MUFU = ~115 FLOP.
        # Round 1's Jacobi sweeps needed ~620 algorithmic (~1200 executed) FLOP per pair for the same result.
        flop_per_pair = 115.0
        sm_mhz = self.peaks.get('sm_max_mhz', 1965.0)
        fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        compute = {'bound': 'fp32', 'achieved': pairs * flop_per_pair / (roof_ms * 1e-3) / 1e12, 'peak': fp32_peak,
                   'unit': 'TFLOP/s', 'flop_per_pair': flop_per_pair,
                   'jacobi_equivalent_tflops': pairs * 620.0 / (roof_ms * 1e-3) / 1e12,
                   'note': 'closed-form eigenvalues (largest of W and of adj W by the trigonometric formula, middle one '
                           'from det W): 5x fewer FLOP per pair than the Jacobi sweeps of round 1, so the FP32 fraction '
                           'falls while the time per Gram drops 3x; ncu: issue slots 57 %, FMA pipe 55 %, XU pipe 58 % '
                           '-- no single pipe binds, the mix does'}
        compute['frac'] = compute['achieved'] / fp32_peak

        # e2e through the reference-facing API: pinned host tensors in, host float64 Gram out
        import gabotorch_b200 as g
        kern = g.SpdAffineInvariantGaussianKernel(beta_min=0.5)
        x1h = torch.from_numpy(st['x_host']).pin_memory()
        x2h = torch.from_numpy(st['x2_host'].copy()).pin_memory()
        e2e_steps = max(3, min(args.steps, 50))
        with torch.no_grad():
            for _ in range(3):
                kh = kern.forward(x1h, x2h)
            assert kh.device.type == 'cpu' and kh.dtype == torch.float64 and tuple(kh.shape) == (N_POINTS, N_POINTS)
            self.barrier()
            self.clocks.timed(True)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                kh = kern.forward(x1h, x2h)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            self.clocks.timed(False)
        dt = self.max_over_ranks(dt)
        e2e = {'value': self.world * pairs * e2e_steps / dt, 'unit': UNIT,
               'h2d_bytes_per_step': int(x1h.numel() * 8 + x2h.numel() * 8), 'd2h_bytes_per_step': int(kh.numel() * 8),
               'ms_per_step': 1e3 * dt / e2e_steps, 'steps': e2e_steps,
               'api': 'SpdAffineInvariantGaussianKernel.forward(x1_host, x2_host) -> host float64'}
        # keep the device result honest: e2e output equals the device-resident one
        self.spd_gram_step(st, float(kern.beta.detach()))
        torch.cuda.synchronize()
        assert float((st['out'].cpu() - kh).abs().max()) < 1e-12
        return dict(value=value, ms_per_step=ms_per_step, launches=launches, roofline=roofline, compute=compute, e2e=e2e)

    # -- extras ---------------------------------------------------------------------------------------------
    def extra_spd(self, n, d, beta, symmetric, steps=10):
        st = self.spd_gram_setup(n, d, 4321 + d)
        ms = self.time_steps(lambda: self.spd_gram_step(st, beta, symmetric), steps, 3) / steps
        return {'workload': 'spd_ai_gram SPD(%d) N=%d%s fp64 out' % (d, n, ' symmetric (x1 is x2)' if symmetric else ''),
                'pairs_per_s': self.world * n * n / (ms * 1e-3), 'ms_per_step': ms}

    def extra_sphere(self, n, D, beta, out_dtype, steps=10):
        torch, lb, p = self.torch, self._lib, self.p
        rng = np.random.default_rng(99 + D + self.rank)
        x = torch.from_numpy(sphere_sample(rng, n, D)).to(self.dev)
        out = torch.empty(n, n, dtype=out_dtype, device=self.dev)
        code = lb.GABO_F32 if out_dtype == torch.float32 else lb.GABO_F64

        def step():
            lb.check(self.lib.gabo_sphere_gram(p(x), n, p(x), n, D, beta, lb.KIND_GAUSS, p(out), code, n,
                                               lb.stream_ptr()), 'gabo_sphere_gram')
        ms = self.time_steps(step, steps, 3) / steps
        nbytes = n * n * out.element_size() + 2 * n * D * 8
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {'workload': 'sphere_gram S^%d (D=%d) N=%d %s out' % (D - 1, D, n, 'fp32' if code == 0 else 'fp64'),
                'pairs_per_s': self.world * n * n / (ms * 1e-3), 'ms_per_step': ms,
                'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': self.hbm_peak, 'unit': 'GB/s',
                             'frac': gbs / self.hbm_peak, 'algorithmic_bytes_per_launch': nbytes}}

    def extra_acq_sphere(self, R, T, D=6, n_train=32, noise=1e-2, steps=5, forced=True):
        """BASELINE configs[2]: Ackley on S^5, R restarts x T CG steps per rank, then ONE all-gather of the best records."""
        import gabotorch_b200 as g
        from gabotorch_b200 import manifold_optimization as mo
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(2024)                      # the GP is replicated: same data on every rank
        xt = sphere_sample(rng, n_train, D)
        y = ackley_sphere(xt)
        base = g.SphereGaussianKernel(beta_min=1.0)            # D = 6: beta_min 1.0 (hd_gabo_sphere.py:122-123)
        model = g.ManifoldGP(torch.from_numpy(xt), torch.from_numpy(y), g.ScaleKernel(base), noise=noise)
        model.covar_module.outputscale = 1.0
        acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
        gp = acq.device_gp()
        rs = np.random.default_rng(777)
        x0_all = sphere_sample(rs, R * self.world, D)          # starts keyed by GLOBAL restart index
        lo = self.rank * R
        x0 = torch.from_numpy(x0_all[lo:lo + R]).to(self.dev)
        gidx = torch.arange(lo, lo + R, device=self.dev)
        res = {}

        def step():
            if forced:   # maxiter = T + 1 with the stopping tolerances disabled: every restart runs exactly T iterations
                cand, val, iters, _ = ops.acq_rcg(gp, x0, maxiter=T + 1, mingradnorm=0.0, minstepsize=-1.0)
            else:        # pymanopt's defaults (mingradnorm 1e-6, minstepsize 1e-10): restarts stop when converged
                cand, val, iters, _ = ops.acq_rcg(gp, x0, maxiter=T)
            slot, best = ops.argmax_records(val, gidx)
            if self.world > 1:
                s = slot                                                   # device-side gather of the winner
                v, gi, c = mo.allgather_records(best.reshape(()), gidx[s].reshape(()), cand[s].reshape(-1))
                win, _ = ops.argmax_records(v, gi)
                res['best'] = (v, gi, win)
            else:
                res['best'] = (best, gidx[slot], slot)
            res['iters'] = iters
        ms = self.time_steps(step, steps, 3, flush=False) / steps
        iters = res['iters'].double().mean().item()
        if not forced:
            return {'workload': 'acq RCG on EI, S^%d, %d restarts/GPU, ConjugateGradient(maxiter=%d) with pymanopt default '
                                'stopping rules, n_train=%d, noise=%g' % (D - 1, R, T, n_train, noise),
                    'solves_per_s': self.world * R / (ms * 1e-3), 'executed_iterations_per_s': self.world * R * iters / (ms * 1e-3),
                    'ms_per_step': ms, 'mean_iters': iters}
        out = {'workload': 'acq RCG on EI, S^%d, %d restarts/GPU x %d CG steps, n_train=%d, noise=%g'
                           % (D - 1, R, T, n_train, noise),
               'candidates_per_s': self.world * R * T / (ms * 1e-3), 'ms_per_step': ms, 'mean_iters': iters,
               'collective': 'one all_gather of (value, gidx, candidate) per solve' if self.world > 1 else 'none (1 GPU)'}
        if self.rank == 0 and self.world == 1 and not self.args.no_cpu_baseline:
            out['cpu_baseline'] = cpu_acq_baseline('sphere', (xt, y, float(base.beta.detach()), noise), x0_all, T)
        return out

    def extra_ei_screen(self, R=1 << 20, D=6, n_train=32, noise=1e-2, steps=5):
        """A2 / A4: raw-sample screening -- EI at R random points of S^5 in one launch (manifold_optimize.py:297-309)."""
        import gabotorch_b200 as g
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(2024)
        xt = sphere_sample(rng, n_train, D)
        y = ackley_sphere(xt)
        base = g.SphereGaussianKernel(beta_min=1.0)
        model = g.ManifoldGP(torch.from_numpy(xt), torch.from_numpy(y), g.ScaleKernel(base), noise=noise)
        model.covar_module.outputscale = 1.0
        gp = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False).device_gp()
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(17 + self.rank)
        x = torch.randn(R, D, dtype=torch.float64, device=self.dev, generator=gen)
        x = x / x.norm(dim=-1, keepdim=True)
        ms = self.time_steps(lambda: ops.ei_eval(gp, x), steps, 3, flush=False) / steps
        return {'workload': 'EI screening of %d raw samples on S^%d, n_train=%d (one launch)' % (R, D - 1, n_train),
                'ei_evals_per_s': self.world * R / (ms * 1e-3), 'ms_per_step': ms}

    def extra_acq_spd(self, R_total, T, d=8, n_train=32, noise=1e-2, steps=3, check=True):
        """BASELINE configs[3]: Ackley on SPD(8), R_total restarts IN TOTAL sharded over the ranks by `shard_range`
        (global restart index), T CG steps each, ONE all-gather of (value, global index, candidate) records and the
        same deterministic argmax on every rank.  `check`: rank 0 also solves ALL R_total starts alone (outside the
        timed region) and the gathered winner must be bit-identical."""
        import gabotorch_b200 as g
        from gabotorch_b200 import manifold_optimization as mo
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(2025)
        xm = spd_sample_mandel(rng, n_train, d)                # GP inputs: Mandel vectors (replicated on every rank)
        mats = ops.mandel_unpack(torch.from_numpy(xm)).cpu().numpy()
        lam, q = np.linalg.eigh(mats / 2.0)                    # Ackley in the tangent space at 2 I (test_functions_spd.py:34-69)
        logm = 2.0 * (q * np.log(lam)[:, None, :]) @ np.swapaxes(q, -1, -2)
        iu = np.triu_indices(d)
        v = logm[:, iu[0], iu[1]]
        dv = v.shape[1]
        y = (-20.0 * np.exp(-0.2 * np.sqrt((v ** 2).sum(1) / dv)) - np.exp(np.cos(2 * np.pi * v).sum(1) / dv) + 20.0 + np.e)
        base = g.SpdAffineInvariantGaussianKernel(beta_min=0.22)   # d = 8: nearest tabulated beta_min (gabo_spd.py:151-162)
        model = g.ManifoldGP(torch.from_numpy(xm), torch.from_numpy(y), g.ScaleKernel(base), noise=noise)
        model.covar_module.outputscale = 1.0
        acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
        gp = acq.device_gp()
        rs = np.random.default_rng(778)
        x0_all = spd_sample_mandel(rs, R_total, d)             # starts keyed by GLOBAL restart index
        lo, hi = mo.shard_range(R_total, self.rank, self.world)
        x0 = ops.mandel_unpack(torch.from_numpy(x0_all[lo:hi]).to(self.dev))
        gidx = torch.arange(lo, hi, device=self.dev)
        res = {}

        def step():
            cand, val, iters, _ = ops.acq_rcg(gp, x0, maxiter=T + 1, mingradnorm=0.0, minstepsize=-1.0)
            slot, best = ops.argmax_records(val, gidx)
            if self.world > 1:
                v_, gi, c = mo.allgather_records(best.reshape(()), gidx[slot].reshape(()), cand[slot].reshape(-1))
                win, bv = ops.argmax_records(v_, gi)
                res['winner'] = (gi[win].reshape(()), bv.reshape(()), c[win].reshape(-1))
            else:
                res['winner'] = (gidx[slot].reshape(()), best.reshape(()), cand[slot].reshape(-1))
            res['iters'] = iters
        ms = self.time_steps(step, steps, 3, flush=False) / steps
        out = {'workload': 'acq RCG on EI, SPD(%d), %d restarts in total (%d per GPU, shard_range) x %d CG steps, '
                           'n_train=%d, noise=%g' % (d, R_total, hi - lo, T, n_train, noise),
               'candidates_per_s': R_total * T / (ms * 1e-3), 'ms_per_step': ms,
               'mean_iters': res['iters'].double().mean().item(),
               'collective': 'one all_gather of (value, gidx, candidate) per solve' if self.world > 1 else 'none (1 GPU)'}
        wg, wv, wc = (t.cpu() for t in res['winner'])
        out['winner_gidx'], out['winner_value'] = int(wg), float(wv)
        if check and self.world > 1:
            # every rank must hold the same winner; rank 0 re-solves all the starts alone and compares bit for bit
            box = torch.stack([wg.double().reshape(()), wv.reshape(())]).to(self.dev)
            gathered = [torch.empty_like(box) for _ in range(self.world)]
            self.dist.all_gather(gathered, box)
            same_everywhere = all(torch.equal(t, gathered[0]) for t in gathered)
            ok = None
            if self.rank == 0:
                xa = ops.mandel_unpack(torch.from_numpy(x0_all).to(self.dev))
                ca, va, _, _ = ops.acq_rcg(gp, xa, maxiter=T + 1, mingradnorm=0.0, minstepsize=-1.0)
                s1, b1 = ops.argmax_records(va, torch.arange(R_total, device=self.dev))
                ok = bool(int(s1) == int(wg) and float(b1) == float(wv)
                          and torch.equal(ca[int(s1)].reshape(-1).cpu(), wc))
                out['single_gpu_winner_gidx'], out['single_gpu_winner_value'] = int(s1), float(b1)
            out['winner_identical_on_every_rank'] = bool(same_everywhere)
            out['winner_matches_single_gpu'] = ok
            if self.rank == 0:
                assert same_everywhere and ok, 'sharded argmax differs from the single-GPU argmax: %r' % (out,)
        if self.rank == 0 and self.world == 1 and not self.args.no_cpu_baseline:
            x0m = ops.mandel_unpack(torch.from_numpy(x0_all[:64])).cpu().numpy()
            out['cpu_baseline'] = cpu_acq_baseline('spd', (mats, y, float(base.beta.detach()), noise), x0m, T,
                                                   max_restarts=64)
        return out

    def extra_spd_strong(self, n=8192, d=3, steps=5):
        """ONE fixed n x n SPD(d) Gram split in row blocks over the ranks (strong scaling, no collective)."""
        torch, lb, p = self.torch, self._lib, self.p
        from gabotorch_b200 import manifold_optimization as mo
        xm = spd_sample_mandel(np.random.default_rng(4321 + d), n, d)
        lo, hi = mo.shard_range(n, self.rank, self.world)
        rows = hi - lo
        fs = self.lib.gabo_spd_factor_stride(d)
        x1 = torch.from_numpy(np.ascontiguousarray(xm[lo:hi])).to(self.dev)
        x2 = torch.from_numpy(xm).to(self.dev)
        f1 = torch.empty(rows, fs, dtype=torch.float64, device=self.dev)
        f2 = torch.empty(n, fs, dtype=torch.float64, device=self.dev)
        flags = torch.zeros(1, dtype=torch.int32, device=self.dev)
        out = torch.empty(rows, n, dtype=torch.float64, device=self.dev)
        beta = BETA_SPD3

        def step():
            s = lb.stream_ptr()
            lb.check(self.lib.gabo_spd_factor2(p(x1), rows, p(x2), n, d, 1, p(f1), p(f2), p(flags), s), 'gabo_spd_factor2')
            lb.check(self.lib.gabo_spd_ai_gram(p(f1), rows, p(f2), n, d, beta, lb.KIND_GAUSS, lb.GABO_F32, 0, p(out),
                                               lb.GABO_F64, n, s), 'gabo_spd_ai_gram')
        ms = self.time_steps(step, steps, 3) / steps
        return {'workload': 'ONE SPD(%d) Gram N=%d fp64 out, row blocks of %d rows per GPU (strong scaling, no collective)'
                            % (d, n, rows), 'pairs_per_s': n * n / (ms * 1e-3), 'ms_per_step': ms, 'scaling': 'strong'}

    def extra_sphere_strong(self, n=32768, D=3, steps=5, gather=False):
        """ONE fixed n x n sphere Gram (fp64 out, the API's dtype) split in row blocks; `gather`: followed by the
        all-gather of the blocks so that every rank holds the whole matrix (report with and without)."""
        torch, lb, p = self.torch, self._lib, self.p
        from gabotorch_b200 import manifold_optimization as mo
        x = sphere_sample(np.random.default_rng(99 + D), n, D)
        lo, hi = mo.shard_range(n, self.rank, self.world)
        if gather and n % self.world:
            return {'workload': 'sphere Gram all-gather', 'skipped': 'n not divisible by the world size'}
        rows = hi - lo
        x1 = torch.from_numpy(np.ascontiguousarray(x[lo:hi])).to(self.dev)
        x2 = torch.from_numpy(x).to(self.dev)
        beta = 6.5 + math.log(2.0)
        full = torch.empty(n, n, dtype=torch.float64, device=self.dev) if gather and self.world > 1 else None
        out = full[lo:hi] if full is not None else torch.empty(rows, n, dtype=torch.float64, device=self.dev)

        def step():
            lb.check(self.lib.gabo_sphere_gram(p(x1), rows, p(x2), n, D, beta, lb.KIND_GAUSS, p(out), lb.GABO_F64, n,
                                               lb.stream_ptr()), 'gabo_sphere_gram')
            if full is not None:
                self.dist.all_gather_into_tensor(full, out)
        ms = self.time_steps(step, steps, 3) / steps
        nbytes = n * n * 8 + 2 * n * D * 8
        return {'workload': 'ONE sphere Gram S^%d N=%d fp64 out, row blocks of %d rows per GPU (strong scaling)%s'
                            % (D - 1, n, rows, ' + all_gather of the blocks (every rank ends with the full matrix)'
                               if full is not None else ', no collective'),
                'pairs_per_s': n * n / (ms * 1e-3), 'ms_per_step': ms, 'scaling': 'strong',
                'aggregate_write_GBps': nbytes / (ms * 1e-3) / 1e9}

    def extra_projection(self, n, D=20, d=5, steps=10):
        torch, ops = self.torch, self.ops
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(5 + self.rank)
        dvh, dvl = D * (D + 1) // 2, d * (d + 1) // 2
        x = torch.randn(n, dvh, dtype=torch.float32, device=self.dev, generator=gen)
        w, _ = np.linalg.qr(np.random.default_rng(3).standard_normal((D, d)))
        pack = ops.nested_projection_matrix(torch.from_numpy(w))
        y = torch.empty(n, dvl, dtype=torch.float32, device=self.dev)

        def step():
            self._lib.check(self.lib.gabo_nested_spd_project(self.p(x), n, D, d, self.p(pack), self.p(y),
                                                             self._lib.stream_ptr()), 'gabo_nested_spd_project')
        ms = self.time_steps(step, steps, 3) / steps
        nbytes = n * 4 * (dvh + dvl)
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {'workload': 'nested projection SPD(%d)->SPD(%d), N=%d Mandel vectors, 3xTF32' % (D, d, n),
                'matrices_per_s': self.world * n / (ms * 1e-3), 'ms_per_step': ms,
                'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': self.hbm_peak, 'unit': 'GB/s',
                             'frac': gbs / self.hbm_peak, 'algorithmic_bytes_per_launch': nbytes}}

    # -- SURVEY 8(f) "next" rows -----------------------------------------------------------------------------------
    def extra_acq_rtr(self, R=1024, D=6, n_train=32, noise=1e-2, steps=5):
        """The reference's default acquisition solver: TrustRegions (robust_trust_regions.py) with the finite-difference
        Hessian, its own stopping rules (mingradnorm 1e-6, maxiter 1000); R restarts per GPU on S^5."""
        import gabotorch_b200 as g
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(2024)
        xt = sphere_sample(rng, n_train, D)
        y = ackley_sphere(xt)
        model = g.ManifoldGP(torch.from_numpy(xt), torch.from_numpy(y),
                             g.ScaleKernel(g.SphereGaussianKernel(beta_min=1.0)), noise=noise)
        model.covar_module.outputscale = 1.0
        gp = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False).device_gp()
        x0 = torch.from_numpy(sphere_sample(np.random.default_rng(777 + self.rank), R, D)).to(self.dev)
        res = {}

        def step():
            res['out'] = ops.acq_rtr(gp, x0)
        ms = self.time_steps(step, steps, 3, flush=False) / steps
        iters = res['out'][2].double().mean().item()
        return {'workload': 'acq trust regions (tCG, FD Hessian) on EI, S^%d, %d restarts/GPU, TrustRegions() defaults, '
                            'n_train=%d, noise=%g' % (D - 1, R, n_train, noise),
                'solves_per_s': self.world * R / (ms * 1e-3), 'ms_per_step': ms, 'mean_outer_iters': iters}

    def extra_gp_fit(self, n_train=32, D=3, batch=4096, steps=5):
        """GP hyper-parameter fit (gabo_gp_mll): marginal-likelihood + gradient evaluations per second for a batch of
        hyper-parameter sets, and the wall time of one fit_gpytorch_model (scipy L-BFGS-B driving one launch + one
        48-byte read-back per objective evaluation) on the model of gabo_sphere.py:131-147."""
        import gabotorch_b200 as g
        from gabotorch_b200 import gp_fit
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(11)
        xt = sphere_sample(rng, n_train, D)
        y = ackley_sphere(xt)
        base = g.SphereGaussianKernel(beta_min=6.5)
        dmat, _ = gp_fit.kernel_distance_matrix(base, torch.from_numpy(xt))
        yd = torch.from_numpy(y).to(self.dev)
        th = np.column_stack([6.5 + rng.random(batch) * 3, 0.1 + rng.random(batch) * 5, 1e-3 + rng.random(batch),
                              rng.standard_normal(batch)])
        th = torch.from_numpy(th).to(self.dev)
        ms = self.time_steps(lambda: ops.gp_mll(dmat, yd, th), steps, 3, flush=False) / steps
        res = {}
        for opt in ('device', 'scipy'):
            t_fit = []
            for _ in range(3):
                cov = g.ScaleKernel(g.SphereGaussianKernel(beta_min=6.5), outputscale_prior=g.GammaPrior(2.0, 0.15))
                model = g.ManifoldGP(xt, y, cov, noise=2.0, mean=0.0, noise_prior=g.GammaPrior(1.1, 0.05))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g.fit_gpytorch_model(model, optimizer=opt)
                t_fit.append((time.perf_counter() - t0) * 1e3)
            res[opt] = (min(t_fit), model.fit_result)
        # the fit kernel alone (device time, 1 start and 64 starts side by side)
        obj = gp_fit.MarginalLogLikelihood(dmat, y, 6.5, outputscale_prior=(2.0, 0.15), noise_prior=(1.1, 0.05))
        raw0 = obj.inverse_transform((6.5 + math.log(2.0), math.log(2.0), 2.0, 0.0))
        pri = [0.0, 0.0, 2.0, 0.15, 1.1, 0.05]
        kern_ms = {}
        for nb in (1, 64):
            st = torch.from_numpy(np.repeat(raw0[None], nb, 0) + (np.random.default_rng(1).standard_normal((nb, 4))
                                                                 * np.array([1.0, 1.0, 2.0, 0.0]) if nb > 1 else 0.0))
            kern_ms[nb] = self.time_steps(lambda: ops.gp_fit(dmat, yd, st, 6.5, 1e-8, pri, [0, 0, 0, 0]), steps, 2,
                                          flush=False) / steps
        return {'workload': 'GP marginal likelihood + gradient, n_train=%d on S^%d, %d hyper-parameter sets per launch; '
                            'fit_gpytorch_model on the gabo_sphere.py model: one-launch device BFGS (gabo_gp_fit) and '
                            'scipy L-BFGS-B over one launch + read-back per evaluation' % (n_train, D - 1, batch),
                'mll_evaluations_per_s': batch / (ms * 1e-3), 'ms_per_step': ms,
                'fit_ms': res['device'][0], 'fit_objective': res['device'][1]['objective'],
                'fit_objective_evaluations': res['device'][1]['evaluations'],
                'fit_iterations': res['device'][1]['iterations'],
                'fit_kernel_ms_1_start': kern_ms[1], 'fit_kernel_ms_64_starts': kern_ms[64],
                'scipy_fit_ms': res['scipy'][0], 'scipy_fit_objective': res['scipy'][1]['objective'],
                'scipy_fit_objective_evaluations': res['scipy'][1]['evaluations']}

    def extra_reconstruct(self, n=1 << 16, D=20, d=5, steps=5):
        """projection_from_nested_spd_to_spd (nested_spd_utils.py:51-118): batched sqrtm + reconstruction, fp64."""
        from gabotorch_b200 import nested_mappings as nm
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(9)
        q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        a = rng.standard_normal((D - d, D - d))
        c = a @ a.T + np.eye(D - d)
        k = rng.standard_normal((d, D - d))
        k = 0.7 * k / np.linalg.norm(k, 2)
        rec = nm.NestedSpdReconstruction(torch.from_numpy(q[:, :d].copy()), torch.from_numpy(q[:, d:].copy()),
                                         torch.from_numpy(c), torch.from_numpy(k))
        b = torch.randn(n, d, d, dtype=torch.float64, device=self.dev)
        ylow = b @ b.transpose(-1, -2) + torch.eye(d, dtype=torch.float64, device=self.dev)
        ms = self.time_steps(lambda: rec(ylow), steps, 3) / steps
        ms_sqrt = self.time_steps(lambda: ops.spd_sqrtm(ylow), steps, 3) / steps
        nbytes = n * 8 * (2 * 2 * d * d + D * D)        # y read twice + sqrt written and read + x written
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {'workload': 'nested SPD reconstruction SPD(%d)->SPD(%d), N=%d, fp64 (sqrtm + rotation, 2 launches)'
                            % (d, D, n),
                'matrices_per_s': self.world * n / (ms * 1e-3), 'ms_per_step': ms, 'sqrtm_ms': ms_sqrt,
                'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': self.hbm_peak, 'unit': 'GB/s',
                             'frac': gbs / self.hbm_peak, 'algorithmic_bytes_per_launch': nbytes}}

    def extra_recon_fit(self, n=40, D=10, d=3, steps=5):
        """Reconstruction-parameter fit of the nested SPD mapping (nested_spd_optimization.py:95-186 as run by
        hd_gabo_spd.py:230-233: ALM around ConjugateGradient(maxiter=100), log-Euclidean cost) on data the mapping can
        reproduce; one cost + gradient evaluation (ONE gabo_sym_eig launch over the n reconstructed matrices), the
        eigensolver kernel alone on a screening-sized batch, and the reference's cost evaluation (oracle restatement: n
        eigendecompositions in a Python loop, host, 1 core) next to it."""
        import time
        import gabotorch_b200 as g
        from gabotorch_b200 import nested_optimization as nopt
        torch, ops = self.torch, self.ops
        rng = np.random.default_rng(21)
        np.random.seed(21)
        q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        w, v = q[:, :d].copy(), q[:, d:].copy()
        c = np.diag(np.linspace(1.0, 2.0, D - d))
        k = rng.standard_normal((d, D - d))
        k = 0.5 * k / np.linalg.norm(k)
        b = rng.standard_normal((n, d, d))
        y = torch.from_numpy(b @ b.transpose(0, 2, 1) + 0.5 * np.eye(d)).to(self.dev)
        tw, tv, tc, tk = (torch.from_numpy(t).to(self.dev) for t in (w, v, c, k))
        x = nopt._reconstruct_spd(y, nopt._SpectralFn.apply(y, 1), tw, tv, tc, tk)
        fn = nopt._SpdReconstructionCost(x, y, tw, 'log_euclidean')
        params = [t.clone().requires_grad_(True) for t in (tv, tc * 1.1, tk * 0.9)]

        def evaluate():
            for t in params:
                t.grad = None
            fn(*params).backward()
        ms_eval = self.time_steps(evaluate, steps, 3, flush=False) / steps
        batch = 100 * n
        mats = x.repeat(100, 1, 1)
        ms_eig = self.time_steps(lambda: ops.sym_eig(mats), steps, 3, flush=False) / steps
        t0 = time.perf_counter()
        g.optimize_reconstruction_parameters_nested_spd(
            x, y, tw, g.ConjugateGradient(maxiter=100), cost_function=g.min_log_euclidean_distance_reconstruction_cost,
            nb_init_candidates=100, maxiter=30)
        fit_s = time.perf_counter() - t0
        log = g.optimize_reconstruction_parameters_nested_spd.last_log
        out = {'workload': 'nested SPD reconstruction fit SPD(%d)->SPD(%d), N=%d data, ALM(30) around CG(100), log-Euclidean '
                           'cost, 100 candidates screened in one batch' % (d, D, n),
               'fit_s': fit_s, 'start_cost': log['start_cost'], 'final_cost': log['cost'], 'outer_iterations': log['iterations'],
               'inner_iterations': int(sum(i for i, _ in log['inner'])), 'cuda_graph': bool(log.get('cuda_graph')), 'cost_and_grad_ms_eager': ms_eval,
               'sym_eig_ms': ms_eig, 'sym_eig_matrices_per_s': batch / (ms_eig * 1e-3), 'sym_eig_batch': batch}
        if self.rank == 0 and self.world == 1:
            from oracle import nested as onest
            xs, ys = x.cpu().numpy(), y.cpu().numpy()
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                onest.min_log_euclidean_distance_reconstruction_cost(xs, ys, w, v, c * 1.1, k * 0.9)
            out['cpu_cost_only_ms'] = (time.perf_counter() - t0) / reps * 1e3
            out['cpu_kind'] = 'oracle port of the reference cost (value only, no gradient), 1 core'
        return out

    def extra_acq_tr_spd(self, constrained, R=256, d=3, n_train=32, noise=1e-2, steps=5, lockstep=True):
        """The solver configurations of gabo_spd.py on SPD(d): plain TrustRegions(maxiter=100) or
        ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100) with one max-eigenvalue inequality constraint, R restarts
        solved by ONE launch (gabo_acq_rtr / gabo_acq_ctr: one warp per restart, fp64, finite-difference Hessian).
        `lockstep`: the same solves through the host-driven lock-step driver of round 1 (still the route for generic
        constraint callables), wall clock, for comparison."""
        import functools
        from gabotorch_b200 import manifold_optimization as mo, riemannian_utils as ru
        torch, ops, _lib = self.torch, self.ops, self._lib
        rng = np.random.default_rng(31)
        xv = spd_sample_mandel(rng, n_train, d)
        y = np.array([float(np.sum(v * v)) for v in xv])
        y = (y - y.mean()) / (y.std() + 1e-12)
        import gabotorch_b200 as g
        model = g.ManifoldGP(torch.from_numpy(xv), torch.from_numpy(y),
                             g.ScaleKernel(g.SpdAffineInvariantGaussianKernel(beta_min=0.5)), noise=noise)
        model.covar_module.outputscale = 1.0
        gp = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False, compute='f64').device_gp()
        x0 = ops.mandel_unpack(torch.from_numpy(spd_sample_mandel(np.random.default_rng(32 + self.rank), R, d,
                                                                  min_eig=0.5, max_eig=2.5)))
        res = {}
        if constrained:
            def step():
                res['out'] = ops.acq_ctr(gp, x0, [('max', 3.0)], maxiter=100, mingradnorm=1e-4)
        else:
            def step():
                res['out'] = ops.acq_rtr(gp, x0, maxiter=100)
        ms = self.time_steps(step, steps, 3, flush=False) / steps
        it = res['out'][2].double()
        out = {'workload': 'acq %s on EI, SPD(%d), %d restarts/GPU, ONE launch (one warp per restart, fp64), n_train=%d'
                           % ('ConstrainedTrustRegions(mingradnorm=1e-4, maxiter=100) + max-eigenvalue constraint'
                              if constrained else 'TrustRegions(maxiter=100)', d, R, n_train),
               'solves_per_s': self.world * R / (ms * 1e-3), 'ms_per_step': ms, 'mean_outer_iters': it.mean().item(),
               'max_outer_iters': int(it.max().item())}
        if lockstep:
            kw = dict(maxiter=100)
            if constrained:
                cons = [functools.partial(ru.max_eigenvalue_constraint_torch, maximum_eigenvalue=3.0)]
                kw.update(mingradnorm=1e-4, ineq_constraints=mo.batched_constraints(cons, _lib.SPD))
            mo.batched_trust_regions(gp, x0, **kw)          # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref = mo.batched_trust_regions(gp, x0, **kw)
            torch.cuda.synchronize()
            lms = (time.perf_counter() - t0) * 1e3
            out['lockstep_driver_ms'] = lms
            out['lockstep_driver_solves_per_s'] = self.world * R / (lms * 1e-3)
            out['same_iteration_counts_as_lockstep'] = float((ref[2] == res['out'][2]).double().mean().item())
        return out

    def guarded(self, fn, *a, **kw):
        """Extras of the 'next' rows must never take the headline line down with them."""
        try:
            return fn(*a, **kw)
        except Exception as e:   # noqa: BLE001
            self.clocks.timed(False)            # never leave the clock sampler in a timed region
            return {'workload': fn.__name__, 'error': '%s: %s' % (type(e).__name__, e)}

    def cpu_baseline(self):
        torch = self.torch
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        rng = np.random.default_rng(1234)
        x = spd_sample_mandel(rng, N_POINTS, SPD_D)
        rows = self.args.cpu_rows
        cpu_reference_step(x[:4], x, BETA_SPD3)               # warm-up
        dt = cpu_reference_step(x[:rows], x, BETA_SPD3)
        dv = min(cpu_vectorised_step(x[:256], x, BETA_SPD3) for _ in range(2))
        return {'value': rows * N_POINTS / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                'sample': '%d rows x %d columns of the N=%d Gram (%d pairs, %.1f s): oracle port of the reference loop '
                          '(one symmetric eigen-solve per pair in Python, spd_utils_torch.py:109-110)'
                          % (rows, N_POINTS, N_POINTS, rows * N_POINTS, dt),
                'vectorised_value': 256 * N_POINTS / dv,
                'vectorised_note': 'same arithmetic with batched torch.linalg.cholesky/eigh on all cores (not what the '
                                   'reference does; reported so the speed-up is not credited to the Python loop)'}

    def run(self):
        args = self.args
        if args.only in ('acq', 'acq_spd'):      # developer switch: just one acquisition extra
            e = (self.extra_acq_sphere(R=args.acq_restarts, T=200) if args.only == 'acq'
                 else self.extra_acq_spd(args.acq_restarts, T=args.acq_steps, d=args.acq_dim))
            self.clocks.stop()
            if self.rank == 0:
                emit(e)
            return 0
        if args.only == 'next':                   # developer switch: the SURVEY 8(f) extras only
            out = [self.guarded(f) for f in (self.extra_acq_rtr, self.extra_gp_fit, self.extra_reconstruct)]
            out += [self.guarded(self.extra_acq_tr_spd, False), self.guarded(self.extra_acq_tr_spd, True),
                    self.guarded(self.extra_acq_tr_spd, True, R=4096, lockstep=False),
                    self.guarded(self.extra_acq_tr_spd, True, R=1024, d=8, lockstep=False)]
            self.clocks.stop()
            if self.rank == 0:
                emit({'extra': out})
            return 0
        if args.only == 'reconfit':               # developer switch: the nested SPD reconstruction fit
            out = [self.guarded(self.extra_recon_fit)]
            self.clocks.stop()
            if self.rank == 0:
                emit({'extra': out})
            return 0
        if args.only == 'trspd':                  # developer switch: the one-launch constrained trust regions on SPD(3)
            out = [self.guarded(self.extra_acq_tr_spd, True, R=4096, lockstep=False)]
            self.clocks.stop()
            if self.rank == 0:
                emit({'extra': out})
            return 0
        if args.only == 'scale':                  # developer switch: the sharded (multi-GPU) extras only
            out = [self.guarded(self.extra_acq_spd, 4096, 200), self.guarded(self.extra_spd_strong),
                   self.guarded(self.extra_sphere_strong), self.guarded(self.extra_sphere_strong, gather=True),
                   self.guarded(self.extra_projection, 1 << 20)]
            self.clocks.stop()
            if self.rank == 0:
                emit({'n_gpus': self.world, 'extra': out})
            return 0
        head = self.headline()
        extras = []
        if not args.no_extras:
            torch = self.torch
            # the sharded paths of SURVEY 8(e), at every N (so that the 1/2/4/8 runs line up):
            extras.append(self.extra_acq_spd(4096, 200))                      # BASELINE configs[3], winner asserted
            extras.append(self.extra_acq_sphere(R=1024, T=200))              # configs[2] per GPU + the record all-gather
            extras.append(self.guarded(self.extra_spd_strong))
            extras.append(self.guarded(self.extra_sphere_strong))
            if self.world > 1:
                extras.append(self.guarded(self.extra_sphere_strong, gather=True))
            extras.append(self.extra_projection(1 << 20))                     # configs[4]: N sharded, 2^20 per GPU
            if self.world == 1:
                extras.append(self.extra_acq_sphere(R=1024, T=200, forced=False))
                extras.append(self.extra_ei_screen())
                extras.append(self.extra_spd(N_POINTS, SPD_D, BETA_SPD3, symmetric=True))
                extras.append(self.extra_spd(2048, 8, 0.22 + math.log(2.0), symmetric=False, steps=5))
                extras.append(self.extra_sphere(256, 3, 6.5 + math.log(2.0), torch.float64, steps=20))
                extras.append(self.extra_sphere(32768, 3, 6.5 + math.log(2.0), torch.float32, steps=5))
                extras.append(self.extra_sphere(32768, 3, 6.5 + math.log(2.0), torch.float64, steps=5))   # the API's dtype
                extras.append(self.extra_sphere(32768, 9, 0.6 + math.log(2.0), torch.float32, steps=5))
                extras.append(self.extra_sphere(32768, 9, 0.6 + math.log(2.0), torch.float64, steps=5))
                extras.append(self.guarded(self.extra_acq_rtr))
                extras.append(self.guarded(self.extra_gp_fit))
                extras.append(self.guarded(self.extra_reconstruct))
                extras.append(self.guarded(self.extra_recon_fit))
                extras.append(self.guarded(self.extra_acq_tr_spd, False))
                extras.append(self.guarded(self.extra_acq_tr_spd, True))
                extras.append(self.guarded(self.extra_acq_tr_spd, True, R=4096, lockstep=False))
                extras.append(self.guarded(self.extra_acq_tr_spd, True, R=1024, d=8, lockstep=False))
        cpu, parity = None, None
        if self.rank == 0 and self.world == 1 and not args.no_cpu_baseline:
            cpu = self.cpu_baseline()
            # achieved accuracy of the timed configuration, every pair against the oracle (checker, not timed)
            xh = spd_sample_mandel(np.random.default_rng(1234), N_POINTS, SPD_D)
            parity = self.guarded(spd_parity_report, xh, BETA_SPD3, self._lib.GABO_F32)
        self.clocks.stop()
        clocks = self.clocks.summary()
        if self.rank == 0:
            line = {
                'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': self.world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'n_points': N_POINTS, 'spd_dim': SPD_D, 'beta': BETA_SPD3,
                           'output': 'float64 N x N', 'pairs_per_step_per_gpu': N_POINTS * N_POINTS,
                           'timing': 'CUDA events per step on the launching stream; L2 flushed (256 MB write) between '
                                     'steps; max over ranks',
                           'sharding': 'row blocks of ONE (%d x %d) Gram: rank g builds rows [%d g, %d (g+1)) of x1 against '
                                       'the replicated x2, no data-path collective (weak scaling: %d x %d pairs per GPU)'
                                       % (N_POINTS * self.world, N_POINTS, N_POINTS, N_POINTS, N_POINTS, N_POINTS),
                           'arithmetic': 'fp64 per-point Cholesky, fp64 triangular product per pair, fp32 closed-form eigenvalues (d = 3), 2^t on the MUFU',
                           'parity_vs_oracle_all_pairs': parity},
                'roofline': head['roofline'], 'compute_roofline': head['compute'], 'cpu_baseline': cpu,
                'e2e': head['e2e'], 'gpu_launches': head['launches'], 'clocks': clocks, 'extra': extras,
            }
            emit(line)
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """Everything libraries print on stdout while the bench runs (NCCL's version banner, torchrun notices) goes to
    stderr; the ONE JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + '\n')
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-extras', action='store_true')
    ap.add_argument('--only', default='', help='developer switch: run a single extra (acq)')
    ap.add_argument('--acq-restarts', type=int, default=1024)
    ap.add_argument('--acq-steps', type=int, default=200)
    ap.add_argument('--acq-dim', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-rows', type=int, default=640, help='rows of the N=2048 Gram in the cpu_baseline sample')
    ap.add_argument('--ref-rows', type=int, default=16, help='rows per step of the reference arm')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)          # both arms: at least 3 warm-up steps
    if args.impl == 'reference':
        return run_reference_arm(args)
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when called as `python bench.py --gpus N`
        import subprocess
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000)] + sys.argv
        return subprocess.call(cmd, stdout=_REAL_STDOUT)     # the children write their JSON line to the real stdout
    return Bench(args).run()


if __name__ == '__main__':
    sys.exit(main())
