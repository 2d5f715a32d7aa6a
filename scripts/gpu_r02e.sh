#!/bin/bash
# Round 2, visit E: SPD(2/3) closed form v2 (adjugate route) parity + timing; sphere Gram knob variants
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gram_gpu.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_e.log
timeout 300 python scripts/dev_spd3.py > gpurun_out/dev_spd3_e.log 2>&1
for v in def u1def u4def u1 u2 u4 u2m8 u2m9 u2m10 u2m12 u2t64 u4t64; do echo "== $v"; timeout 60 ./scripts/micro/sphere_$v; done > gpurun_out/sphere_variants.log 2>&1
cat gpurun_out/pytest_e.log gpurun_out/dev_spd3_e.log gpurun_out/sphere_variants.log
