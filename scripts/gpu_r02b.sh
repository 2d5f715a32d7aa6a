#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_acq_gpu.py tests/test_api_gpu.py tests/test_lockstep_gpu.py tests/test_gram_gpu.py -m gpu -q 2>&1 | tail -60 ) > gpurun_out/pytest_b.log
tail -60 gpurun_out/pytest_b.log
