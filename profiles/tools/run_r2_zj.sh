#!/bin/bash
# GPU pass ZJ of round 2 (1 GPU): the reference's own benchmark water (24 k / 96 k atoms) and the reference's benchmark protocol through
# the compiled shim (BenchmarkSystem(32) = 96 k atoms) with the end-of-round kernel
mkdir -p gpurun_out/r2zj
O=gpurun_out/r2zj
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 32 pme 50 > $O/nbnxm_bench_gpu_96k_pme.json 2> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_96k_pme.json
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 32 rf 50 > $O/nbnxm_bench_gpu_96k_rf.json 2>> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_96k_rf.json
timeout 600 python bench.py --workload ref_water_24k --secondary ref_water_96k --no-sustained --no-cpu --steps 50 > $O/bench_refwater.json 2> $O/bench_refwater.err; echo "refwater rc=$?"
python - <<'E'
import json
d=json.load(open('gpurun_out/r2zj/bench_refwater.json'))
for m in (d, d['secondary']):
    print(m['config']['workload'], 'step',round(m['ms_per_step'],4),'kernel',round(m['roofline']['kernel_ms'],4),'frac',round(m['roofline']['frac'],4),'e2e',round(m['e2e']['ms_per_step'],4), 'useful', round(m['details']['useful_lane_fraction'],3))
E
