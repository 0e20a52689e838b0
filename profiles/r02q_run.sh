#!/bin/bash
# steps in flight: default bench line (depth 2), depth 1 and 3 beside it, GPU suite
mkdir -p gpurun_out
T=${1:-r02q}
echo skip tests

for d in 4 6; do
timeout 900 python bench.py --steps 10 --warmup 3 --depth $d $( [ $d != 4 ] && echo --no-cpu-baseline ) > gpurun_out/${T}_bench_depth$d.json 2> gpurun_out/${T}_bench_depth$d.err; echo "bench depth $d rc=$?"
tail -2 gpurun_out/${T}_bench_depth$d.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_depth$d.json').read().strip().splitlines()[-1])
print('value',d['value'],'serial',d['value_one_step_at_a_time'],'exact',d.get('value_exact'),'e2e',d['e2e']['value'],d['e2e']['value_one_step_at_a_time'],'ms',d['ms_per_step'],'busy',d['roofline']['sm_time_busy_frac'],d['roofline']['sm_time_busy_frac_in_flight'])
print(d.get('parity')); print(d.get('cpu_baseline'))"
done
