#!/bin/bash
# round 2, call W: host profile of the inference pipeline
mkdir -p gpurun_out
timeout 600 python tools/gpu_infer_hostprof.py > gpurun_out/r2w_infer_hostprof.log 2>&1; echo "hostprof rc=$?"; head -90 gpurun_out/r2w_infer_hostprof.log | cut -c1-180
