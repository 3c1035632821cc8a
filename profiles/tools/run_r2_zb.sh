#!/bin/bash
# GPU pass ZB of round 2 (1 GPU): pair-search step with 6 / 8 / 10 resident search CTAs per SM; k_order_assign without the contended atomics
mkdir -p gpurun_out/r2zb
O=gpurun_out/r2zb
for lib in default sb8 sb10; do
  L=""; [ $lib != default ] && L=scratch/lib_$lib.so
  B200NB_LIBRARY=$L timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-sustained --secondary "" > $O/bench_$lib.json 2> $O/bench_$lib.err
  python - <<E
import json
d=json.load(open('$O/bench_$lib.json'))
print('$lib', 'step', round(d['ms_per_step'],4), {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['search'].items() if not k.endswith('note') and k!='scenario'})
E
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tile_list or dynamic or pairs_forces" > $O/pytest.txt 2>&1; tail -3 $O/pytest.txt
