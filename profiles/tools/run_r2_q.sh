#!/bin/bash
# GPU pass Q of round 2 (1 GPU): half-entry packing (4 i-atoms per packed j list, two half-entries per warp): parity, then the
# kernel sweep against the r2-o numbers, the j-force red as v2 + scalar (default) against one red.v4
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reflist.py tests/test_gpu_paths.py -q -x > $O/pytest_parity.txt 2>&1; tail -15 $O/pytest_parity.txt
for lib in default v4; do
  for wl in water_24k water_192k water_1M; do
    L=""; [ $lib = v4 ] && L=scratch/lib_v4.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
timeout 300 python profiles/tools/kernel_sweep.py water_192k 0 rf 2>&1 | grep -v Warning | tail -1 | sed "s/^/default rf /" | tee -a $O/sweep.txt
timeout 900 python -m pytest tests/test_gpu_domdec.py tests/test_gpu_domdec_ipc.py -q -x > $O/pytest_dd.txt 2>&1; tail -5 $O/pytest_dd.txt
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'E'
import json
d=json.load(open('gpurun_out/r2q/bench_n1.json'))
print('1M step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'useful',d['details']['useful_lane_fraction'])
print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
print('24k',d['secondary']['ms_per_step'],d['secondary']['roofline']['kernel_ms'],d['secondary']['roofline']['frac'])
E
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_force_192k python bench.py --workload water_192k --secondary "" --steps 3 --warmup 3 --no-cpu --no-search --no-sustained > $O/ncu_full_192k.log 2>&1
ncu -i $O/prof_force_192k.ncu-rep --page raw --csv > $O/prof_force_192k_raw.csv 2>/dev/null; timeout 300 python profiles/tools/ncu_summary.py $O/prof_force_192k_raw.csv > $O/ncu_k_force_water192k.txt 2>&1; cat $O/ncu_k_force_water192k.txt
ncu -i $O/prof_force_192k.ncu-rep --page source --csv > $O/src_sass.csv 2>/dev/null
rm -f $O/prof_force_192k.ncu-rep
