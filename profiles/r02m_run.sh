#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
T=${1:-r02m}
timeout 600 ncu --set full --clock-control none --import-source on -c 40 -o /tmp/ncu/mesh python profiles/ncu_mesh_modes.py > gpurun_out/${T}_mesh_ncu.log 2>&1; echo "mesh ncu rc=$?"
ncu -i /tmp/ncu/mesh.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_mesh_and_eval_kernels.csv 2>/dev/null
ls -la gpurun_out/ | tail -4
