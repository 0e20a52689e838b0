#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02d}
timeout 300 python -m pytest tests/test_gpu_fit_frames.py -m gpu -x -q -s -k "wide" > gpurun_out/${T}_wide_tests.log 2>&1; echo "wide pytest rc=$?"
tail -15 gpurun_out/${T}_wide_tests.log
