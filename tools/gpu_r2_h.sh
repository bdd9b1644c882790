#!/bin/bash
# round 2, call H: node-centred projection, full GPU suite, training step both dense modes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "projection or two_tile or headline" > gpurun_out/r2h_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -6 gpurun_out/r2h_pytest_a.log
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2h_h16_check.log 2>&1; echo "h16_check rc=$?"; tail -4 gpurun_out/r2h_h16_check.log
for mode in simt h16; do
DIG3D_TRAIN_DENSE=$mode timeout 600 python tools/gpu_train_step.py 12 > gpurun_out/r2h_train_$mode.log 2>&1; echo "train $mode rc=$?"; grep "^mode" gpurun_out/r2h_train_$mode.log
done
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2h_pytest.log
