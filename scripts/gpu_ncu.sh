#!/bin/bash
# One full ncu capture: gpu_ncu.sh <kernel regex> <launches to skip> <report name> <command ...>
mkdir -p gpurun_out
k=$1; s=$2; o=$3; shift 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o gpurun_out/$o -f "$@" > gpurun_out/$o.log 2>&1
tail -5 gpurun_out/$o.log
