#!/bin/bash
# Round 2, visit A (1 GPU): parity suite, smoke, bench, D2H micro-benchmark, the ncu captures VERDICT r01 asked for
# (sphere Gram at N = 32768, fp32 and fp64 out; launch list of one lock-step trust-region solve).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 1700 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1.json
( timeout 300 python bench.py --impl reference --steps 3 --warmup 3 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
timeout 120 python scripts/micro/d2h_bw.py > gpurun_out/d2h_n1.log 2>&1
timeout 300 python scripts/dev_sphere_big.py > gpurun_out/sphere_big.log 2>&1
# full captures: launches 1+3 (warm-ups are 2 per variant, timed 5): skip to a timed launch of D=3 fp32 and D=3 fp64
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_gram_kernel -s 3 -c 1 \
    -o gpurun_out/prof_sphere_gram_f32_n32768 -f python scripts/dev_sphere_big.py > gpurun_out/prof_sphere32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_gram_kernel -s 10 -c 1 \
    -o gpurun_out/prof_sphere_gram_f64_n32768 -f python scripts/dev_sphere_big.py > gpurun_out/prof_sphere64.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_lockstep_tr.csv \
    python scripts/dev_rtr_spd.py > gpurun_out/launches_lockstep.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 1500 gpurun_out/bench_n1.json; echo; tail -5 gpurun_out/bench.err
cat gpurun_out/sphere_big.log; cat gpurun_out/d2h_n1.log | tail -3
