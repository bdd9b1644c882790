#!/bin/bash
# round 2, call B: where does the two-tile chain spend its time?  timeline + ncu source-level capture
mkdir -p gpurun_out
timeout 300 python tools/gpu_h16_timeline.py > gpurun_out/r2b_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/r2b_timeline.log | tail -50
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2b_h16_check.log 2>&1; echo "h16_check rc=$?"; tail -4 gpurun_out/r2b_h16_check.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_h16 -s 2 -c 1 -o gpurun_out/r02_b_h16 python tools/gpu_one_forward.py 2 > gpurun_out/r2b_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triplet_gather_node -s 2 -c 1 -o gpurun_out/r02_gather_node python tools/gpu_one_forward.py 2 > gpurun_out/r2b_ncu_g.log 2>&1; echo "ncu g rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -k "two_tile or headline or out_of_range or comenet or indices" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
ls -la gpurun_out | tail
