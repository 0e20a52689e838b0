#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02j}
timeout 600 python -m pytest tests/test_gpu_fit_frames.py -m gpu -q -s -k float64 > gpurun_out/${T}_f64.log 2>&1; echo "pytest rc=$?"
grep -n "passed\|failed\|frame [01]: vertex\|^E " gpurun_out/${T}_f64.log | tail -20
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'exact',d.get('value_exact'),'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'evals',d['evals_per_frame'],d['roofline']['evals_max_frame'])
print(d.get('parity')); print(d.get('cpu_baseline'))"
