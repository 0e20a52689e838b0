#!/bin/bash
# development builds of the library (float32 kernels only: half the compile time), side by side:
#   libsfx_dev.so       product flags
#   libsfx_dev_prof.so  + -DSFX_CYCLE_PROF (clock64 lap timers, profiles/prof_cycles.py)
# Neither is the product build (__graft_entry__.build()).
cd "$(dirname "$0")/../smplify-x-partial_b200/csrc" || exit 1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -shared -DSFX_DEV_F32_ONLY"
nvcc $F -Xptxas -v -o libsfx_dev.so sfx_lib.cu > /tmp/dev_build.log 2>&1 &
if [ "$1" != "noprof" ]; then nvcc $F -DSFX_CYCLE_PROF -o libsfx_dev_prof.so sfx_lib.cu > /tmp/dev_prof_build.log 2>&1 & fi
wait
grep -i " error" /tmp/dev_build.log /tmp/dev_prof_build.log | head
grep -A2 "fit_pipeline_kernelIf" /tmp/dev_build.log | grep -E "spill|registers"
