#!/bin/bash
# round 2, call D: warpgroup register re-allocation (setmaxnreg), update_v on the tensor engine
mkdir -p gpurun_out
timeout 300 python tools/gpu_h16_small.py > gpurun_out/r2d_small.log 2>&1; echo "small rc=$?"; tail -3 gpurun_out/r2d_small.log
timeout 300 python tools/gpu_h16_timeline.py > gpurun_out/r2d_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/r2d_timeline.log | tail -48
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2d_h16_check.log 2>&1; echo "h16_check rc=$?"; tail -22 gpurun_out/r2d_h16_check.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_train.py -x -q -k "two_tile or headline or out_of_range or update_v or flat_adam or node_centred or headline_size or config5" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2d_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_h16 -s 2 -c 1 -o gpurun_out/r02_b_h16_v3 python tools/gpu_one_forward.py 2 > gpurun_out/r2d_ncu_b.log 2>&1; echo "ncu b rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_update_v_h16 -c 1 -o gpurun_out/r02_update_v_h16 python tools/gpu_one_forward.py 2 > gpurun_out/r2d_ncu_v.log 2>&1; echo "ncu v rc=$?"
