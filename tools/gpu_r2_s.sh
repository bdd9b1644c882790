#!/bin/bash
# round 2, call S: host-side profile of the training step
mkdir -p gpurun_out
timeout 600 python tools/gpu_train_hostprof.py > gpurun_out/r2s_hostprof.log 2>&1; echo "hostprof rc=$?"; head -75 gpurun_out/r2s_hostprof.log | cut -c1-200
