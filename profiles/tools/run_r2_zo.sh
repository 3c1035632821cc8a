#!/bin/bash
# GPU pass ZO of round 2 (1 GPU): the whole -m gpu suite and smoke on the final tree
mkdir -p gpurun_out/r2zo
O=gpurun_out/r2zo
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -5 $O/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -3 $O/smoke.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --no-sustained > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'search', d['search']['search_ms'])"
