#!/bin/bash
# Round 2, visit C (1 GPU): parity suite, smoke, both bench arms, D2H micro-benchmark, launch list of the bench command,
# full ncu captures: sphere Gram at N = 32768 (fp32 / fp64 out), SPD(3) Gram, the one-launch SPD trust-region kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 900 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1.json
( timeout 300 python bench.py --impl reference --steps 3 --warmup 3 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
timeout 120 python scripts/micro/d2h_bw.py > gpurun_out/d2h_n1.log 2>&1
timeout 300 python scripts/dev_sphere_big.py > gpurun_out/sphere_big.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sphere_gram_kernel -s 3 -c 1 \
    -o gpurun_out/prof_sphere_gram_f32_n32768 -f python scripts/dev_sphere_big.py > gpurun_out/prof_sphere32.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sphere_gram_kernel -s 10 -c 1 \
    -o gpurun_out/prof_sphere_gram_f64_n32768 -f python scripts/dev_sphere_big.py > gpurun_out/prof_sphere64.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spd_ai_gram_kernel -s 3 -c 1 \
    -o gpurun_out/prof_spd_gram -f python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/prof_spd.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spd_rtr_kernel -s 2 -c 1 \
    -o gpurun_out/prof_spd_rtr -f python bench.py --only trspd > gpurun_out/prof_spd_rtr.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 1500 gpurun_out/bench_n1.json; echo; tail -5 gpurun_out/bench.err
cat gpurun_out/sphere_big.log; tail -3 gpurun_out/d2h_n1.log; tail -3 gpurun_out/prof_spd_rtr.log
