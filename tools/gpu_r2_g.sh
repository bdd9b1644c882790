#!/bin/bash
# round 2, call G: training linears on the two-tile engine
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2g_pytest.log
for mode in simt h16; do
DIG3D_TRAIN_DENSE=$mode timeout 600 python tools/gpu_train_step.py 12 > gpurun_out/r2g_train_$mode.log 2>&1; echo "train $mode rc=$?"; tail -3 gpurun_out/r2g_train_$mode.log
done
