#!/bin/bash
# GPU pass J of round 2 (1 GPU): search kernels after the latency work (warp-aggregated column counting, 4-way candidate
# prefetch in k_search, software-pipelined k_pack), N-D decomposition through the peer-memory windows: whole GPU suite, bench
# search block, launch lists
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 2400 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; tail -15 $O/pytest_gpu.txt
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu --no-sustained > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
python - <<'E'
import json
d=json.load(open('gpurun_out/r2j/bench_n1.json'))
print('1M step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
print({k:v for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
print('24k',d['secondary']['ms_per_step'],d['secondary']['roofline']['kernel_ms'],d['secondary']['roofline']['frac'])
E
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_24k.csv python bench.py --workload water_24k --secondary "" --steps 3 --warmup 3 --no-cpu --no-sustained > $O/ncu_launches_24k.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_1M.csv python bench.py --workload water_1M --secondary "" --steps 3 --warmup 3 --no-cpu --no-sustained > $O/ncu_launches_1M.log 2>&1
ls -la $O
