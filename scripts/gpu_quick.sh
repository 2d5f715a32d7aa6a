#!/bin/bash
# Parity tests + bench on one B200 (no profiler).  Run under gpurun.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
( timeout 600 python bench.py "$@" 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1.json
tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print('value %.3e pairs/s  %.1f us/step  e2e %.3e  roof %.3f  fp32 %.3f' % (d['value'], d['ms_per_step']*1e3, d['e2e']['value'], d['roofline']['frac'], d['compute_roofline']['frac']))
for e in d['extra']:
    print({k: (round(v, 4) if isinstance(v, float) and v < 1e6 else ('%.3e' % v if isinstance(v, float) else v)) for k, v in e.items() if k != 'roofline'}, 'frac=%.3f' % e['roofline']['frac'] if 'roofline' in e else '')
PY
