#!/bin/bash
# round 2, call L: tensor-core weight gradient (parity, timing), training step, two-batch pipeline, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "weight_gradient" > gpurun_out/r2l_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -15 gpurun_out/r2l_pytest_a.log
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q > gpurun_out/r2l_pytest_b.log 2>&1; echo "pytest b rc=$?"; tail -15 gpurun_out/r2l_pytest_b.log
for mode in mixed tc; do
DIG3D_TRAIN_DENSE=$mode timeout 600 python tools/gpu_train_step.py 12 > gpurun_out/r2l_train_$mode.log 2>&1; echo "train $mode rc=$?"; grep "^mode" -A1 gpurun_out/r2l_train_$mode.log | cut -c1-1800
done
timeout 900 python -m pytest tests/test_gpu_train.py -x -q > gpurun_out/r2l_pytest_train.log 2>&1; echo "pytest train rc=$?"; tail -8 gpurun_out/r2l_pytest_train.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2l_bench_quick.json 2> gpurun_out/r2l_bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r2l_bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r2l_bench_quick.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','e2e'): print(k, d.get(k))
"
