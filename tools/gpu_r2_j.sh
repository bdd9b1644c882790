#!/bin/bash
# round 2, call J: tensor-core triplet gather (parity + time), mixed-mode training step, full GPU suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -x -q -k "triplet_gather" > gpurun_out/r2j_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -8 gpurun_out/r2j_pytest_a.log
timeout 600 python tools/gpu_gather_modes.py > gpurun_out/r2j_gather.log 2>&1; echo "gather rc=$?"; tail -5 gpurun_out/r2j_gather.log
for mode in mixed h16; do
DIG3D_TRAIN_DENSE=$mode timeout 600 python tools/gpu_train_step.py 12 > gpurun_out/r2j_train_$mode.log 2>&1; echo "train $mode rc=$?"; grep "^mode" gpurun_out/r2j_train_$mode.log
done
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2j_pytest.log
