#!/bin/bash
# ncu captures of the hot kernels (one launch each) + launch list.  Run under gpurun; reports land in gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spd_ai_gram_kernel -s 3 -c 1 -o gpurun_out/prof_spd_gram -f $B --no-extras > gpurun_out/prof_spd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_gram_kernel -s 25 -c 1 -o gpurun_out/prof_sphere_gram -f $B > gpurun_out/prof_sphere.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nested_project_kernel -s 4 -c 1 -o gpurun_out/prof_project -f $B > gpurun_out/prof_project.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_rcg_cta_kernel -s 2 -c 1 -o gpurun_out/prof_sphere_acq -f $B > gpurun_out/prof_acq.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spd_rcg_cta_kernel -s 1 -c 1 -o gpurun_out/prof_spd_acq -f $B > gpurun_out/prof_spd_acq.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1_full.json
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
ls -la gpurun_out | head -30
