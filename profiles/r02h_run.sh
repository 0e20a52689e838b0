#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02h}
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_gpu_tests.log 2>&1; echo "pytest rc=$?"
grep -n "passed\|failed\|FAILED\|frame [01]: vertex\|loss rel median\|reprojection distance\|wide vs" gpurun_out/${T}_gpu_tests.log | tail -40
