#!/bin/bash
# GPU pass O of round 2 (1 GPU): whole -m gpu suite (tabulated Ewald, C++ benchmark driver through the shim added), smoke, the
# C++ benchmark driver at 96 k atoms, the default bench line and the reference arm
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -12 $O/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 32 pme 50 > $O/nbnxm_bench_gpu_96k_pme.json 2> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_96k_pme.json
OMP_PROC_BIND=spread OMP_PLACES=cores timeout 600 ./shim/_build/nbnxm_bench_gpu 32 rf 50 > $O/nbnxm_bench_gpu_96k_rf.json 2>> $O/nbnxm_bench_gpu_96k.err; tail -1 $O/nbnxm_bench_gpu_96k_rf.json
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?"
timeout 600 python bench.py --workload ref_water_24k --secondary ref_water_96k --no-sustained --steps 50 > $O/bench_refwater.json 2> $O/bench_refwater.err; echo "refwater rc=$?"
python - <<'E'
import json
for f in ('bench_n1','bench_refwater'):
    d=json.load(open('gpurun_out/r2o/%s.json'%f))
    print(f, d['config']['workload'], 'step',round(d['ms_per_step'],4),'value %.4g'%d['value'],'kernel',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],4), 'useful', round(d['details']['useful_lane_fraction'],3))
    if 'search' in d: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
    if d.get('cpu_baseline'): print('cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('search_ms'))
    s=d['secondary']; print(' secondary', s['config']['workload'], round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4), round(s['roofline']['frac'],4), round(s['e2e']['ms_per_step'],4), round(s['details']['useful_lane_fraction'],3))
E
