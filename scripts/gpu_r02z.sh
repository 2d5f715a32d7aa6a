#!/bin/bash
# Round 2, final visit (same plan as visit G, current kernels) (1 GPU), in two calls (gpurun brings back at most 64 MiB of gpurun_out/ per call):
#   gpu_r02g.sh a : full parity suite, smoke, both bench arms, launch list, ncu of the headline kernel + sphere Gram fp32
#   gpu_r02g.sh b : ncu of the other hot kernels
# Summarise with `python scripts/make_profiles.py r02c` after each call (reports accumulate under gpurun_out/).
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
cap() { timeout 400 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f "${@:4}" > gpurun_out/$3.log 2>&1; }
if [ "$1" = "a" ]; then
    nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
    ( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log
    ( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/smoke.log
    ( timeout 900 python bench.py 2> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_n1.json
    ( timeout 300 python bench.py --impl reference --steps 3 --warmup 3 2>> gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_ref.json
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
    cap spd_ai_gram_kernel 3 prof_spd_gram python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline
    cap sphere_gram_kernel 3 prof_sphere_gram_f32_n32768 python scripts/dev_sphere_big.py
    tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; head -c 600 gpurun_out/bench_n1.json; echo; tail -5 gpurun_out/bench.err
else
    cap spd_ai_gram_kernel 18 prof_spd_gram_n8192 python scripts/dev_spd3.py
    cap sphere_gram_kernel 10 prof_sphere_gram_f64_n32768 python scripts/dev_sphere_big.py
    cap nested_project_kernel 3 prof_project python bench.py --only scale
    cap nested_spd_reconstruct_dmma_kernel 2 prof_reconstruct python scripts/dev_recon.py
    cap sym_eig_kernel 2 prof_sym_eig python bench.py --only reconfit
    cap gp_fit_kernel 2 prof_gp_fit python scripts/dev_gpfit.py
    cap spd_rtr_kernel 2 prof_spd_rtr python bench.py --only trspd
fi
du -sh gpurun_out
