#!/bin/bash
# round 2, call K: force training (tangent network) for DimeNet++ / SphereNet
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -k "force" > gpurun_out/r2k_pytest_force.log 2>&1; echo "pytest force rc=$?"; tail -40 gpurun_out/r2k_pytest_force.log
