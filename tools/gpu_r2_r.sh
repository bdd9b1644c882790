#!/bin/bash
# round 2, call R (evidence): launch lists of the inference bench and of a training step, ncu --set full of the tensor-core wgrad
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2r_pytest.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2r_ncu_bench.log 2>&1; echo "ncu bench rc=$?"; tail -2 gpurun_out/r2r_ncu_bench.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 700 --csv --log-file gpurun_out/r02_train_launches.csv python tools/gpu_train_step.py 2 > gpurun_out/r2r_ncu_train.log 2>&1; echo "ncu train rc=$?"; tail -2 gpurun_out/r2r_ncu_train.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 70 -c 1 -o gpurun_out/r02_wgrad_tc python tools/gpu_train_step.py 1 > gpurun_out/r2r_ncu_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:comenet_filter_sum -s 2 -c 1 -o gpurun_out/r02_comenet_filter_sum python tools/gpu_comenet.py > gpurun_out/r2r_ncu_filter.log 2>&1; echo "ncu filter rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*launches.csv 2>/dev/null
