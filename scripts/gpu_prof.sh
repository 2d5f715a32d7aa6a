#!/bin/bash
# ncu captures of the hot kernels (one launch each) + launch list + bench.  Run under gpurun in TWO calls
# (`gpu_prof.sh a`, `gpu_prof.sh b`): gpurun brings back at most 64 MiB of gpurun_out/ per call.
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
cap() { timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f $B $4 > gpurun_out/$3.log 2>&1; }
if [ "$1" = "a" ]; then
    cap spd_ai_gram_kernel 3 prof_spd_gram --no-extras
    cap sphere_gram_kernel 25 prof_sphere_gram
    cap nested_project_kernel 4 prof_project
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/launches.log 2>&1
else
    cap sphere_rcg_cta_kernel 2 prof_sphere_acq
    cap spd_rcg_cta_kernel 1 prof_spd_acq
    ( timeout 600 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1_full.json
    ( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
fi
du -sh gpurun_out
