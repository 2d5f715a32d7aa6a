#!/bin/bash
( timeout 1200 python -m pytest tests/test_acq_gpu.py tests/test_api_gpu.py -m gpu -x -q 2>&1 | tail -6 )
for w in 1 2 4 0; do echo "SPD(8) R=512 T=50 spec=$w"; GABO_ACQ_SPEC=$w timeout 300 python bench.py --only acq_spd --acq-restarts 512 --acq-steps 50 --acq-dim 8 2>&1 | tail -1 | cut -c1-200; done
echo "SPD(3) R=512 T=50"; timeout 300 python bench.py --only acq_spd --acq-restarts 512 --acq-steps 50 --acq-dim 3 2>&1 | tail -1 | cut -c1-200
