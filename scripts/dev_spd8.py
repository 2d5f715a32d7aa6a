"""Developer helper: only the SPD(8) N=2048 Gram (for ncu captures)."""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
a = argparse.Namespace(gpus=1, steps=5, warmup=3, impl='ours', no_extras=True, no_cpu_baseline=True, only='', acq_restarts=512,
                       acq_steps=50, acq_dim=8, cpu_rows=16, ref_rows=4)
b = bench.Bench(a)
print(b.extra_spd(2048, int(sys.argv[1]) if len(sys.argv) > 1 else 8, 0.22 + math.log(2.0), symmetric=False, steps=5))
b.clocks.stop()
