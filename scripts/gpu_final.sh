#!/bin/bash
# Round-end GPU visit: every parity test, smoke, both bench arms, the ncu launch list of the bench command and one full
# ncu capture of each kernel added with the SURVEY 8(f) rows (the captures of the older kernels are kept from the
# earlier visits: those kernels did not change).  Run under gpurun; summarise afterwards with scripts/make_profiles.py.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
( timeout 500 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1.json
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
for spec in "sphere_rtr_kernel prof_sphere_rtr" "gp_mll_kernel prof_gp_mll" "nested_spd_reconstruct_kernel prof_spd_reconstruct"; do
    set -- $spec
    timeout 240 ncu --set full --clock-control none --import-source on -k regex:$1 -s 4 -c 1 -o gpurun_out/$2 -f \
        python bench.py --only next > gpurun_out/$2.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; cut -c1-1500 gpurun_out/bench_n1.json; echo; cut -c1-600 gpurun_out/bench_ref.json; echo
tail -3 gpurun_out/bench.err; wc -l gpurun_out/launches.csv; ls -la gpurun_out/*.ncu-rep
