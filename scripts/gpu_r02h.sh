#!/bin/bash
# 8-GPU visit: scaling bench at N=1,2,4,8 (both arms at N=8), D2H ceiling at 8 ranks, multi-GPU tests
mkdir -p gpurun_out/h
cd "$GRAFT_REPO_ROOT"
NG=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/h/bench_n1.json 2> gpurun_out/h/bench_n1.err
for n in 2 4 8; do
  [ $n -le $NG ] || continue
  timeout 400 $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/h/bench_n$n.json 2> gpurun_out/h/bench_n$n.err
done
timeout 300 $TR --nproc-per-node $NG --master-port 29600 scripts/micro/d2h_bw.py > gpurun_out/h/d2h.log 2>&1
cp gpurun_out/d2h_bw.json gpurun_out/h/ 2>/dev/null
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/h/multigpu_tests.log 2>&1
tail -3 gpurun_out/h/multigpu_tests.log
for n in 1 2 4 8; do tail -1 gpurun_out/h/bench_n$n.json | cut -c1-400; done
nvidia-smi topo -m > gpurun_out/h/topo.txt 2>&1
