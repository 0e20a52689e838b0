#!/bin/bash
# round-2 evidence for the bench line with steps in flight: default line (K = 20), launch list and
# full ncu capture of the one-block pipeline launch, configs 3 and 4
mkdir -p gpurun_out /tmp/ncu
T=${1:-r02u}
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'serial',d['value_one_step_at_a_time'],'exact',d.get('value_exact'),'e2e',d['e2e']['value'],d['e2e']['value_one_step_at_a_time'],'ms',d['ms_per_step'],'busy',d['roofline']['sm_time_busy_frac'],d['roofline']['sm_time_busy_frac_in_flight'])
print(d.get('parity')); print(d.get('cpu_baseline')); print(d['roofline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --single-mode > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fit_pipeline' -c 1 -o /tmp/ncu/pipeline python bench.py --steps 1 --warmup 1 --no-cpu-baseline --single-mode > gpurun_out/${T}_pipeline_ncu.log 2>&1; echo "pipeline ncu rc=$?"
ncu -i /tmp/ncu/pipeline.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_pipeline_kernel.csv 2>/dev/null
timeout 900 python bench.py --steps 5 --warmup 3 --vposer > gpurun_out/${T}_bench_vposer.json 2> gpurun_out/${T}_bench_vposer.err; echo "vposer rc=$?"
timeout 1200 python bench.py --steps 3 --warmup 1 --interpenetration > gpurun_out/${T}_bench_coll.json 2> gpurun_out/${T}_bench_coll.err; echo "coll rc=$?"
for f in vposer coll; do python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_$f.json').read().strip().splitlines()[-1])
print('$f value',d['value'],'serial',d.get('value_one_step_at_a_time'),'e2e',d['e2e']['value'],'ms',d['ms_per_step']); print(d.get('cpu_baseline'))"; done
ls -la gpurun_out | tail -12
