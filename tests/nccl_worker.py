"""Worker of tests/test_multigpu_gpu.py (launched under torchrun, one rank per GPU, NCCL): the sharded paths of SURVEY
8(e) against their single-GPU results, bit for bit.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (input laws only)
import gabotorch_b200 as g  # noqa: E402
from gabotorch_b200 import _lib, ops  # noqa: E402
from gabotorch_b200 import manifold_optimization as mo  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    out = {}

    # --- A1-A4 sharded: S^5, 257 restarts (ragged shards), one all-gather, winner vs the single-GPU solve ------------
    rng = np.random.default_rng(2024)
    xt = bench.sphere_sample(rng, 32, 6)
    y = bench.ackley_sphere(xt)
    model = g.ManifoldGP(torch.from_numpy(xt), torch.from_numpy(y), g.ScaleKernel(g.SphereGaussianKernel(beta_min=1.0)),
                         noise=1e-2)
    model.covar_module.outputscale = 1.0
    gp = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False).device_gp()
    R = 257
    x0 = torch.from_numpy(bench.sphere_sample(np.random.default_rng(777), R, 6)).to(dev)
    lo, hi = mo.shard_range(R, rank, world)
    cand, val, _, _ = ops.acq_rcg(gp, x0[lo:hi], maxiter=60)
    gidx = torch.arange(lo, hi, device=dev)
    slot, best = ops.argmax_records(val, gidx)
    v, gi, c = mo.allgather_records(best.reshape(()), gidx[slot].reshape(()), cand[slot].reshape(-1))
    win, bv = ops.argmax_records(v, gi)
    ca, va, _, _ = ops.acq_rcg(gp, x0, maxiter=60)                      # every rank: the unsharded solve
    s1, b1 = ops.argmax_records(va, torch.arange(R, device=dev))
    out['acq_winner_gidx'] = int(gi[win])
    out['acq_ok'] = bool(int(gi[win]) == int(s1) and float(bv) == float(b1)
                         and torch.equal(c[win].reshape(-1), ca[int(s1)].reshape(-1))
                         and torch.equal(va[lo:hi], val))               # shard values = the same bits as unsharded

    # --- a 3-way tie across ranks and a NaN: lowest global index wins on every rank ----------------------------------
    vals = torch.tensor([0.3, 0.9, float('nan'), 0.9, 0.2, 0.9, 0.0], dtype=torch.float64, device=dev)
    lo2, hi2 = mo.shard_range(7, rank, world)
    g2 = torch.arange(lo2, hi2, device=dev)
    if hi2 > lo2:
        sl, bb = ops.argmax_records(vals[lo2:hi2], g2)
        rec = (bb.reshape(()), g2[sl].reshape(()), vals[lo2:hi2][sl].reshape(-1))
    else:
        rec = (torch.tensor(float('nan'), dtype=torch.float64, device=dev), torch.tensor(1 << 40, device=dev),
               torch.zeros(1, dtype=torch.float64, device=dev))
    v2, gi2, _ = mo.allgather_records(*rec)
    w2, _ = ops.argmax_records(v2, gi2)
    out['tie_ok'] = bool(int(gi2[w2]) == 1)

    # --- Gram row blocks: the gathered blocks equal the single-GPU Gram bit for bit ----------------------------------
    xm = torch.from_numpy(bench.spd_sample_mandel(np.random.default_rng(5), 515, 3)).to(dev)
    lo3, hi3 = mo.shard_range(515, rank, world)
    blk = ops.spd_ai_gram(xm[lo3:hi3], xm.clone(), 0.9, _lib.KIND_GAUSS)
    full = ops.spd_ai_gram(xm, xm.clone(), 0.9, _lib.KIND_GAUSS)
    out['gram_rows_ok'] = bool(torch.equal(blk, full[lo3:hi3]))
    xs = torch.from_numpy(bench.sphere_sample(np.random.default_rng(6), 1024, 3)).to(dev)
    lo4, hi4 = mo.shard_range(1024, rank, world)
    whole = torch.empty(1024, 1024, dtype=torch.float64, device=dev)
    whole[lo4:hi4] = ops.sphere_gram(xs[lo4:hi4], xs, 7.0, _lib.KIND_GAUSS)
    if 1024 % world == 0:
        dist.all_gather_into_tensor(whole, whole[lo4:hi4].clone())
        out['sphere_gather_ok'] = bool(torch.equal(whole, ops.sphere_gram(xs, xs, 7.0, _lib.KIND_GAUSS)))

    # --- the public entry with options={'distributed': True}: same candidate on every rank ---------------------------
    acq = g.ExpectedImprovement(model, best_f=float(y.min()), maximize=False)
    best = mo.joint_optimize_manifold(acq, g.Sphere(6), g.ConjugateGradient(maxiter=50), q=1, num_restarts=16,
                                      raw_samples=200, options={'distributed': True, 'seed': 11})
    ref = mo.joint_optimize_manifold(acq, g.Sphere(6), g.ConjugateGradient(maxiter=50), q=1, num_restarts=16,
                                     raw_samples=200, options={'seed': 11})
    allb = [torch.empty_like(best) for _ in range(world)]
    dist.all_gather(allb, best.contiguous())
    out['joint_ok'] = bool(all(torch.equal(t, allb[0]) for t in allb) and torch.equal(best, ref.to(best.device)))

    flags = torch.tensor([int(all(v for k, v in out.items() if k.endswith('_ok')))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out['all_ranks_ok'] = bool(int(flags) == 1)
    if rank == 0:
        print('NCCL_WORKER ' + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
