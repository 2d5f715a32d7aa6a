#!/bin/bash
# One GPU visit: parity tests, bench, launch list, full ncu captures of the hot kernels.  Run under gpurun.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1.json
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
timeout 300 python scripts/dev_e2e.py > gpurun_out/dev_e2e.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spd_ai_gram_kernel -s 3 -c 1 \
    -o gpurun_out/prof_spd_gram -f python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/prof_spd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_gram_kernel -s 25 -c 1 \
    -o gpurun_out/prof_sphere_gram -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof_sphere.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_acq_kernel -s 2 -c 1 \
    -o gpurun_out/prof_sphere_acq -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof_acq.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench.err; cat gpurun_out/dev_e2e.log | tail -8
