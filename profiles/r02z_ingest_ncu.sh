#!/bin/bash
# ncu --set full of the small kernels: device ingestion / blending (tests/test_gpu_ingest.py) and guess_init
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none -k regex:'pack_keypoints|keypoint_masks|blend_keypoints|guess_init' -c 8 -o /tmp/ncu/ingest python -m pytest tests/test_gpu_ingest.py tests/test_gpu_fit_frames.py -m gpu -q -k "ingest or blend or guess_init or pack or mask" > gpurun_out/r02z_ingest_ncu.log 2>&1; echo "rc=$?"
ncu -i /tmp/ncu/ingest.ncu-rep --page raw --csv > gpurun_out/r02z_ncu_raw_ingest_kernels.csv 2>/dev/null
tail -3 gpurun_out/r02z_ingest_ncu.log; ls -la gpurun_out/r02z_ncu_raw_ingest_kernels.csv
