#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spd_rcg_cta_kernel -s 1 -c 1 -o gpurun_out/prof_spd_acq -f \
    python bench.py --only acq_spd --acq-restarts 512 --acq-steps ${1:-20} --acq-dim ${2:-8} > gpurun_out/prof_spd_acq.log 2>&1
tail -2 gpurun_out/prof_spd_acq.log | cut -c1-200
