#!/bin/bash
# round 2, call A: first run of the two-tile fp16 chain
mkdir -p gpurun_out
timeout 300 python tools/gpu_h16_small.py > gpurun_out/r2a_small.log 2>&1; echo "small rc=$?"; tail -5 gpurun_out/r2a_small.log
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2a_h16_check.log 2>&1; echo "h16_check rc=$?"
tail -30 gpurun_out/r2a_h16_check.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -k "two_tile or headline or out_of_range or tensor_core or packed or comenet or indices" > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2a_pytest.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_h16_small.py > gpurun_out/r2a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 gpurun_out/r2a_memcheck.log
