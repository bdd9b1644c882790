#!/bin/bash
# round 2, call A: first run of the two-tile fp16 chain
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2a_h16_check.log 2>&1; echo "h16_check rc=$?"
tail -30 gpurun_out/r2a_h16_check.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "two_tile or headline or out_of_range or tensor_core or packed" > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2a_pytest.log
