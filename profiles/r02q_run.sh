#!/bin/bash
# steps in flight: bench line at two depths (r02q: 2 and 3 with the GPU suite first; r02s: 4 and 6)
#   bash profiles/r02q_run.sh r02s "4 6"
mkdir -p gpurun_out
T=${1:-r02q}
DEPTHS=${2:-"2 3"}
first=1
for d in $DEPTHS; do
timeout 900 python bench.py --steps 10 --warmup 3 --depth $d $( [ $first != 1 ] && echo --no-cpu-baseline ) > gpurun_out/${T}_bench_depth$d.json 2> gpurun_out/${T}_bench_depth$d.err; echo "bench depth $d rc=$?"
first=0
tail -2 gpurun_out/${T}_bench_depth$d.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_depth$d.json').read().strip().splitlines()[-1])
print('value',d['value'],'serial',d['value_one_step_at_a_time'],'exact',d.get('value_exact'),'e2e',d['e2e']['value'],d['e2e']['value_one_step_at_a_time'],'ms',d['ms_per_step'])"
done
