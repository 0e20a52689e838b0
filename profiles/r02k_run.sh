#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02k}
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mesh" > gpurun_out/${T}_mesh_tests.log 2>&1; echo "pytest rc=$?"
grep -n "passed\|failed\|fused tensor\|^E " gpurun_out/${T}_mesh_tests.log | tail -12
timeout 120 python profiles/bench_mesh.py 128 > gpurun_out/${T}_bench_mesh.txt 2>&1; echo "rc=$?"; cat gpurun_out/${T}_bench_mesh.txt | tail -5
