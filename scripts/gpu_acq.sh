#!/bin/bash
# acquisition kernel: parity tests + speculation-width sweep
( timeout 900 python -m pytest tests/test_acq_gpu.py tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -4 )
for R in 1024 4096; do for w in 1 2 4; do echo "R=$R spec=$w"; GABO_ACQ_SPEC=$w python bench.py --only acq --acq-restarts $R 2>&1 | tail -1 | cut -c1-260; done; done
