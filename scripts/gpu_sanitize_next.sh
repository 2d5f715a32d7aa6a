#!/bin/bash
# compute-sanitizer over the kernels added with the SURVEY 8(f) rows: memcheck on their parity tests, racecheck on the
# shared-memory ones (gp_mll_kernel: in-place Cholesky / inverse behind block barriers; the reconstruction setup kernel:
# warp-synchronous Jacobi; the reconstruction batch kernel: staged tiles), synccheck on the same.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SEL='tests/test_gp_fit_gpu.py tests/test_nested_gpu.py tests/test_acq_gpu.py'
KEY='mll or gp_factor or objective_in_raw or reconstruction or sqrtm or round_trip or chain or rtr'
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_next.log \
    python -m pytest $SEL -m gpu -x -q -k "$KEY" 2>&1 | tail -3 ) > gpurun_out/memcheck_next_pytest.log
tail -1 gpurun_out/memcheck_next_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/memcheck_next.log
( timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --log-file gpurun_out/racecheck_next.log \
    python -m pytest tests/test_gp_fit_gpu.py tests/test_nested_gpu.py -m gpu -x -q \
    -k "mll_and_gradient or gp_factor or reconstruction_golden or right_inverse" 2>&1 | tail -3 ) > gpurun_out/racecheck_next_pytest.log
tail -1 gpurun_out/racecheck_next_pytest.log; grep -E "RACECHECK SUMMARY" gpurun_out/racecheck_next.log
( timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/synccheck_next.log \
    python -m pytest tests/test_gp_fit_gpu.py tests/test_nested_gpu.py tests/test_acq_gpu.py -m gpu -x -q \
    -k "mll_and_gradient or gp_factor or reconstruction_golden or rtr_f64" 2>&1 | tail -3 ) > gpurun_out/synccheck_next_pytest.log
tail -1 gpurun_out/synccheck_next_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/synccheck_next.log
