#!/bin/bash
# compute-sanitizer over the kernels added in round 2: memcheck on their parity tests, racecheck + synccheck on the shared-memory
# / warp-synchronous ones (sym_eig_kernel, gp_fit_kernel's register-resident evaluator, the reconstruction kernel with the
# operator in shared memory, the backward kernels, the one-launch SPD trust regions).
mkdir -p gpurun_out/san
export PYTHONUNBUFFERED=1
cd "$GRAFT_REPO_ROOT"
SEL='tests/test_nested_gpu.py tests/test_grad_gpu.py tests/test_gp_fit_gpu.py tests/test_gram_gpu.py'
KEY="sym_eig or reconstruction_costs or input_gradients or building_blocks or device_fit or fit_matches or closed_form or reconstruction_golden or golden"
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/san/memcheck.log \
    python -m pytest $SEL -m gpu -x -q -k "$KEY and not 30000 and not full" 2>&1 | tail -3 ) > gpurun_out/san/memcheck_pytest.log
tail -1 gpurun_out/san/memcheck_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/san/memcheck.log
( timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --log-file gpurun_out/san/racecheck.log \
    python -m pytest tests/test_nested_gpu.py tests/test_gp_fit_gpu.py tests/test_grad_gpu.py -m gpu -x -q \
    -k "sym_eig or device_fit or fit_matches or reconstruction_golden or building_blocks or nested_spd_kernels_input" 2>&1 | tail -3 ) > gpurun_out/san/racecheck_pytest.log
tail -1 gpurun_out/san/racecheck_pytest.log; grep -E "RACECHECK SUMMARY" gpurun_out/san/racecheck.log
( timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/san/synccheck.log \
    python -m pytest tests/test_nested_gpu.py tests/test_gp_fit_gpu.py tests/test_acq_gpu.py -m gpu -x -q \
    -k "sym_eig or device_fit or fit_matches or reconstruction_golden or rtr or ctr" 2>&1 | tail -3 ) > gpurun_out/san/synccheck_pytest.log
tail -1 gpurun_out/san/synccheck_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/san/synccheck.log
