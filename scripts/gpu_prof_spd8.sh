#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spd_ai_gram_kernel -s 3 -c 1 -o gpurun_out/prof_spd8_gram -f \
    python scripts/dev_spd8.py 8 > gpurun_out/prof_spd8.log 2>&1
tail -1 gpurun_out/prof_spd8.log | cut -c1-150
