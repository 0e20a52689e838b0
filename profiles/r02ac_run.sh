#!/bin/bash
# ncu --set full of the mesh / evaluation kernels at the head (K2 of the fused kernel on float16
# split operands), raw page as CSV; then the GPU suite
mkdir -p gpurun_out /tmp/ncu
T=${1:-r02ac}
timeout 400 ncu --set full --clock-control none --import-source on -c 12 -o /tmp/ncu/mesh python profiles/ncu_mesh_modes.py > gpurun_out/${T}_mesh_ncu.log 2>&1; echo "mesh ncu rc=$?"
ncu -i /tmp/ncu/mesh.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_raw_mesh_kernels.csv 2>/dev/null
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_gpu_tests.log
