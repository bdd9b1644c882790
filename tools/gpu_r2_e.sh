#!/bin/bash
# round 2, call E: 5-stage ring, coalesced tile stores, startup probes; wider test sweep
mkdir -p gpurun_out
timeout 300 python tools/gpu_h16_timeline.py > gpurun_out/r2e_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/r2e_timeline.log | tail -50
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2e_h16_check.log 2>&1; echo "h16_check rc=$?"; tail -5 gpurun_out/r2e_h16_check.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2e_pytest.log
