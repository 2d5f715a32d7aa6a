#!/bin/bash
# compute-sanitizer passes over the parity tests: memcheck on the small cases of every kernel, racecheck and synccheck on
# the multi-warp acquisition kernels (shared solver state, barriers inside data-dependent loops).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_gram_gpu.py tests/test_nested_gpu.py tests/test_manifold_gpu.py tests/test_api_gpu.py -m gpu -x -q \
    -k "not full_size and not 30000" 2>&1 | tail -3 ) > gpurun_out/memcheck_pytest.log
tail -1 gpurun_out/memcheck_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/memcheck.log
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --log-file gpurun_out/racecheck.log \
    python -m pytest tests/test_acq_gpu.py -m gpu -x -q -k "speculative or improves or ei_and_gradient" 2>&1 | tail -3 ) > gpurun_out/racecheck_pytest.log
tail -1 gpurun_out/racecheck_pytest.log; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck.log
( timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/synccheck.log \
    python -m pytest tests/test_acq_gpu.py tests/test_gram_gpu.py -m gpu -x -q -k "speculative or improves or golden" 2>&1 | tail -3 ) > gpurun_out/synccheck_pytest.log
tail -1 gpurun_out/synccheck_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/synccheck.log
