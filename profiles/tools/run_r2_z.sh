#!/bin/bash
# GPU pass Z of round 2 (1 GPU): device-side repartitioning (csrc/dd_partition.cu) against the host restatement, the cheaper rolling
# prune part, the slot-index stream marked evict-first in L2
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 1500 python -m pytest tests/test_gpu_domdec.py tests/test_gpu_domdec_ipc.py tests/test_gpu_paths.py -q -x > $O/pytest_dd.txt 2>&1; tail -15 $O/pytest_dd.txt
for lib in default ef; do
  for wl in water_192k water_1M; do
    L=""; [ $lib != default ] && L=scratch/lib_$lib.so
    B200NB_LIBRARY=$L timeout 300 python profiles/tools/kernel_sweep.py $wl 0 ewald 2>&1 | grep -v Warning | tail -1 | sed "s/^/$lib /" | tee -a $O/sweep.txt
  done
done
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'E'
import json
d=json.load(open('gpurun_out/r2z/bench_n1.json'))
print('1M step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],'useful',d['details']['useful_lane_fraction'])
print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
E
