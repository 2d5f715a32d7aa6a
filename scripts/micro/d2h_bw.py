#!/usr/bin/env python
"""Host-write ceiling of the box: per-rank device -> pinned-host bandwidth with 1 / 2 / 4 / ... / WORLD ranks writing
at the same time (the e2e metric ships a 33.5 MB fp64 Gram block per rank per step over PCIe).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/micro/d2h_bw.py

Three ways of moving the same 2048 x 2048 fp64 block (33.5 MB) into pinned host memory:
  ce          one cudaMemcpyAsync (copy engine) from a device buffer
  ce_chunked  the same in 4 row chunks on a side stream (what a staged, pipelined delivery would do)
  zero_copy   a kernel that stores straight into the mapped host buffer: gabo_sphere_gram (fp64 out, N = 2048: 6 us of
              arithmetic, so the time is the PCIe write path), the mechanism SpdAffineInvariantGaussianKernel.forward uses
Ranks >= k idle while k ranks write; time = CUDA events per rank, max over the active ranks; aggregate = k * bytes / time.
Writes one JSON document to gpurun_out/d2h_bw.json (rank 0).
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gabotorch_b200 import _lib  # noqa: E402

N, REPS = 2048, 20


def main():
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    src = torch.randn(N, N, dtype=torch.float64, device=dev)
    host = torch.empty(N, N, dtype=torch.float64, pin_memory=True)
    rng = np.random.default_rng(rank)
    x = rng.standard_normal((N, 3))
    x = torch.from_numpy(x / np.linalg.norm(x, axis=1, keepdims=True)).to(dev)
    side = torch.cuda.Stream()
    nbytes = N * N * 8

    def ce():
        host.copy_(src, non_blocking=True)

    def ce_chunked():
        step = N // 4
        for lo in range(0, N, step):
            host[lo:lo + step].copy_(src[lo:lo + step], non_blocking=True)

    def zero_copy():
        _lib.check(lib.gabo_sphere_gram(ctypes.c_void_p(x.data_ptr()), N, ctypes.c_void_p(x.data_ptr()), N, 3, 7.19, 0,
                                        ctypes.c_void_p(host.data_ptr()), _lib.GABO_F64, N, _lib.stream_ptr()),
                   'gabo_sphere_gram')

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REPS):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / REPS

    levels = [k for k in (1, 2, 4, 8, 16) if k <= world]
    out = {'bytes_per_transfer': nbytes, 'world': world, 'levels': {}}
    for k in levels:
        row = {}
        for name, fn in (('ce', ce), ('ce_chunked', ce_chunked), ('zero_copy', zero_copy)):
            if world > 1:
                dist.barrier()
            ms = timed(fn) if rank < k else 0.0
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            row[name] = {'ms': ms, 'per_rank_GBps': nbytes / ms / 1e6, 'aggregate_GBps': k * nbytes / ms / 1e6}
        out['levels'][str(k)] = row
    if rank == 0:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'd2h_bw.json'), 'w') as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
