#!/bin/bash
# GPU pass I of round 2 (1 GPU): prune-while-packing (k_pack on the outer list, one pass), single-pass search, no per-search
# host syncs: whole -m gpu suite, bench line (search block), launch lists of a re-search at 24 k and 1 M
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 2400 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.txt 2>&1; tail -15 $O/pytest_gpu.txt
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
python - <<'E'
import json
d=json.load(open('gpurun_out/r2i/bench_n1.json'))
print('1M step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms'],'frac',d['roofline']['frac'])
print(json.dumps(d['search'],indent=1))
print('24k',d['secondary']['ms_per_step'],d['secondary']['roofline']['kernel_ms'],d['secondary']['roofline']['frac'])
E
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_24k.csv python bench.py --workload water_24k --secondary "" --steps 3 --warmup 3 --no-cpu --no-sustained > $O/ncu_launches_24k.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_1M.csv python bench.py --workload water_1M --secondary "" --steps 3 --warmup 3 --no-cpu --no-sustained > $O/ncu_launches_1M.log 2>&1
ls -la $O
