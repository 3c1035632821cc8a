#!/bin/bash
# GPU pass ZM of round 2 (1 GPU): execution order = stable counting sort (packing = grid order within a step count): warps that run side by
# side belong to neighbouring i-clusters; parity, kernel sweep against ZL's default column, search-step time
mkdir -p gpurun_out/r2zm
O=gpurun_out/r2zm
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reflist.py tests/test_gpu_paths.py -q > $O/pytest_parity.txt 2>&1; tail -3 $O/pytest_parity.txt
for cfg in "water_1M ewald" "water_1M rf" "water_192k ewald" "water_24k ewald"; do
  set -- $cfg
  timeout 300 python profiles/tools/kernel_sweep.py $1 0 $2 2>&1 | grep -v Warning | tail -1 | sed "s/^/stable $2 /" | tee -a $O/sweep.txt
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-sustained > $O/bench.json 2> $O/bench.err
python - <<'E'
import json
d=json.load(open('gpurun_out/r2zm/bench.json'))
print('step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
s=d['secondary']; print('24k', round(s['ms_per_step'],4), round(s['roofline']['kernel_ms'],4))
E
