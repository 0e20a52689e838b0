#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02o}
timeout 300 python -m pytest tests/test_gpu_ingest.py -m gpu -q -s > gpurun_out/${T}_ingest.log 2>&1; echo "pytest rc=$?"
grep -n "passed\|failed\|^E " gpurun_out/${T}_ingest.log | tail -12
