#!/bin/bash
# GPU pass ZG of round 2 (1 GPU), for the record at the end of the round: whole -m gpu suite, smoke, ncu --set full of the force
# kernel at 24 k / 192 k / 1 M atoms (traffic.json), launch list, memcheck of the device-side repartitioning, default bench line
# and the reference arm
mkdir -p gpurun_out/r2zg
O=gpurun_out/r2zg
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt
for wl in water_24k water_192k water_1M; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_force -s 3 -c 1 -o $O/prof_$wl python bench.py --workload $wl --secondary "" --steps 3 --warmup 3 --no-cpu --no-search --no-sustained > $O/ncu_$wl.log 2>&1
  ncu -i $O/prof_$wl.ncu-rep --page raw --csv > $O/raw_$wl.csv 2>/dev/null
  python profiles/tools/ncu_summary.py $O/raw_$wl.csv > $O/ncu_k_force_$wl.txt 2>&1
  rm -f $O/prof_$wl.ncu-rep
  grep "gpu__time_duration\|fma_cycles_active\|issue_active\|dram__bytes_read" $O/ncu_k_force_$wl.txt
done
python profiles/tools/ncu_traffic.py profiles/r2/traffic.json water_24k=$O/raw_water_24k.csv=profiles/r2/zg_ncu_k_force_water_24k.txt water_192k=$O/raw_water_192k.csv=profiles/r2/zg_ncu_k_force_water_192k.txt water_1M=$O/raw_water_1M.csv=profiles/r2/zg_ncu_k_force_water_1M.txt > /dev/null; cp profiles/r2/traffic.json $O/traffic.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_water1M.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sustained --secondary "" > $O/launches_bench.log 2>&1
T1='tests/test_gpu_domdec.py::test_device_repartition_equals_host_repartition[2]'
timeout 1200 compute-sanitizer --tool memcheck --log-file $O/memcheck_device_repartition.log python -m pytest "$T1" -q -x > $O/memcheck_device_repartition.pytest.txt 2>&1
echo "memcheck repartition: $(tail -1 $O/memcheck_device_repartition.pytest.txt) | $(grep 'ERROR SUMMARY' $O/memcheck_device_repartition.log | tail -1)"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; echo "ref rc=$?"
python - <<'E'
import json
d=json.load(open('gpurun_out/r2zg/bench_n1.json'))
print(d['config']['workload'], 'step',round(d['ms_per_step'],4),'value %.4g'%d['value'],'kernel',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['ms_per_step'],4), 'useful', round(d['details']['useful_lane_fraction'],3))
if 'search' in d: print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
if d.get('cpu_baseline'): print('cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('search_ms'))
s=d['secondary']; print(' secondary', s['config']['workload'], round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4), round(s['roofline']['frac'],4), round(s['e2e']['ms_per_step'],4), round(s['details']['useful_lane_fraction'],3))
print(d.get('sustained'))
E
