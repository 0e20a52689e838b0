#!/bin/bash
# verification of the head after the container re-creation: GPU suite, default bench line, cycle profile
mkdir -p gpurun_out
T=${1:-r02p}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${T}_gpu_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'exact',d.get('value_exact'),'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'evals',d['evals_per_frame'],d['roofline']['evals_max_frame'])
print(d.get('parity')); print(d.get('cpu_baseline'))"
timeout 300 python profiles/prof_cycles.py > gpurun_out/${T}_prof_cycles_default.txt 2>&1
head -20 gpurun_out/${T}_prof_cycles_default.txt
