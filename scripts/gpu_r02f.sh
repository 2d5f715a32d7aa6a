#!/bin/bash
# Round 2, visit F: autograd surface + fit_gpytorch_manifold + full GPU suite
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_grad_gpu.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_f.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/pytest_f_all.log
cat gpurun_out/pytest_f.log; tail -15 gpurun_out/pytest_f_all.log
