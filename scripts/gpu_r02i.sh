#!/bin/bash
# visit I: new nested SPD reconstruction-fit tests (eigensolver kernel, costs, fit) on the device
mkdir -p gpurun_out/i
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_nested_gpu.py -m gpu -q -x -s -k "sym_eig or nested_spd_reconstruction or reconstruction_parameters_nested_spd" > gpurun_out/i/tests.log 2>&1
tail -25 gpurun_out/i/tests.log
