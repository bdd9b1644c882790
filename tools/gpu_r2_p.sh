#!/bin/bash
# round 2, call P: ComENet on the tensor engine
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "comenet or baseline_configs" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2p_pytest.log
timeout 600 python tools/gpu_comenet.py > gpurun_out/r2p_comenet.log 2>&1; echo "comenet rc=$?"; tail -12 gpurun_out/r2p_comenet.log | cut -c1-1500
