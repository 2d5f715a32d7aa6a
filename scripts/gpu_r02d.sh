#!/bin/bash
# Round 2, visit D: Gram kernels after the closed-form SPD(2/3) eigenvalues and the ALU-pipe conversion in the sphere Gram
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gram_gpu.py tests/test_api_gpu.py tests/test_nested_gpu.py -m gpu -x -q -s 2>&1 | tail -30 ) > gpurun_out/pytest_d.log
timeout 300 python scripts/dev_spd3.py > gpurun_out/dev_spd3.log 2>&1
timeout 300 python scripts/dev_sphere_big.py > gpurun_out/sphere_big_d.log 2>&1
tail -30 gpurun_out/pytest_d.log; cat gpurun_out/dev_spd3.log gpurun_out/sphere_big_d.log
