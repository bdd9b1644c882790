#!/bin/bash
# round 2, call AI: gather split sweep with the out-edge lists, pipeline depth sweep at the final state
mkdir -p gpurun_out
timeout 600 python tools/gpu_gather_modes.py > gpurun_out/r2ai_gather_modes.log 2>&1; echo "gather modes rc=$?"; cat gpurun_out/r2ai_gather_modes.log | cut -c1-300
timeout 600 python tools/gpu_infer_hostprof.py > gpurun_out/r2ai_infer_hostprof.log 2>&1; echo "hostprof rc=$?"; head -9 gpurun_out/r2ai_infer_hostprof.log | cut -c1-180
