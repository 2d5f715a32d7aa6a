#!/bin/bash
# usage: gpu_prof_one.sh <kernel-regex> <skip> <out-name> [bench args...]   -- one full ncu capture
mkdir -p gpurun_out
k=$1; s=$2; o=$3; shift 3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/$o -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/$o.log 2>&1
tail -2 gpurun_out/$o.log | cut -c1-300
