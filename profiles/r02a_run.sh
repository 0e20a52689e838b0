#!/bin/bash
# round 2, first GPU call: existing GPU tests, the new default bench line (config 2 as 8d),
# the regression-prior line, cycle profiles of both
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02a_gpu_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench_default.json 2> gpurun_out/r02a_bench_default.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --regression-prior --no-cpu-baseline > gpurun_out/r02a_bench_reg.json 2> gpurun_out/r02a_bench_reg.err; echo "bench reg rc=$?"
python profiles/prof_cycles.py > gpurun_out/r02a_prof_cycles_default.txt 2>&1
python profiles/prof_cycles.py --regression-prior > gpurun_out/r02a_prof_cycles_reg.txt 2>&1
cat gpurun_out/r02a_prof_cycles_default.txt
