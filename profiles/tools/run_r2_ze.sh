#!/bin/bash
# GPU pass ZE of round 2 (1 GPU): k_search with the column windows flattened (all columns of an i-cluster at once): parity, timing
mkdir -p gpurun_out/r2ze
O=gpurun_out/r2ze
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reflist.py tests/test_gpu_paths.py -q > $O/pytest_parity.txt 2>&1; tail -4 $O/pytest_parity.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-sustained > $O/bench.json 2> $O/bench.err
python - <<'E'
import json
d=json.load(open('gpurun_out/r2ze/bench.json'))
print('step', round(d['ms_per_step'],4), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
print(d['details'].get('setup'))
E
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-sustained --secondary "" > $O/launches_bench.log 2>&1
python - <<'E'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2ze/launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
tot=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<15: continue
    name=r[4].split("(")[0][:50]; ns=float(r[14].replace(",",""))
    tot.setdefault(name,[0,0]); tot[name][0]+=1; tot[name][1]+=ns
for k,(c,ns) in tot.items():
    if k.startswith("k_"): print("%-52s %4d  %9.1f us avg"%(k,c,ns/1e3/c))
E
