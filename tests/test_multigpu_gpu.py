"""NCCL-side correctness of the sharded paths ("argmax index bit-exact vs single GPU"):
needs >= 2 GPUs, launches tests/nccl_worker.py under torchrun.  (World-size-2 gloo tests of the same host logic run on
CPU in tests/test_host_logic.py.)"""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs (gpurun --gpus N)')
def test_sharded_paths_match_single_gpu_over_nccl():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
           '--master-addr', '127.0.0.1', '--master-port', str(29600 + os.getpid() % 300),
           os.path.join(ROOT, 'tests', 'nccl_worker.py')]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith('NCCL_WORKER ')][-1]
    out = json.loads(line[len('NCCL_WORKER '):])
    assert out['acq_ok'] and out['tie_ok'] and out['gram_rows_ok'] and out['joint_ok'] and out['all_ranks_ok'], out
    assert out.get('sphere_gather_ok', True), out
