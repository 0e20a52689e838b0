#!/bin/bash
# ncu --set full of the pipeline kernel with the interpenetration term on (config 4's launch, 16
# frames so that the ~40 replays of a seconds-long launch stay within minutes); raw page as CSV
mkdir -p gpurun_out /tmp/ncu
T=${1:-r02ab}
timeout 1100 ncu --set full --clock-control none --import-source on -k regex:'fit_pipeline' -c 1 -o /tmp/ncu/coll \
    python bench.py --interpenetration --frames 16 --steps 1 --warmup 1 --depth 1 --no-cpu-baseline --single-mode \
    > gpurun_out/${T}_coll_ncu.log 2>&1; echo "coll ncu rc=$?"
ncu -i /tmp/ncu/coll.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_pipeline_kernel_interpenetration.csv 2>/dev/null
tail -3 gpurun_out/${T}_coll_ncu.log | cut -c1-400
ls -la /tmp/ncu
