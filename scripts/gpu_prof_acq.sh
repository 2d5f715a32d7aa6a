#!/bin/bash
mkdir -p gpurun_out
GABO_ACQ_SPEC=${1:-1} timeout 900 ncu --set full --clock-control none --import-source on -k regex:sphere_rcg_cta_kernel -s 2 -c 1 -o gpurun_out/prof_sphere_acq -f \
    python bench.py --only acq --acq-restarts ${2:-1024} > gpurun_out/prof_sphere_acq.log 2>&1
tail -2 gpurun_out/prof_sphere_acq.log | cut -c1-200
