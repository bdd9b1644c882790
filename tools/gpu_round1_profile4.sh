# round-1 training-path evidence: launch list of one training step + full captures of the backward kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv --log-file gpurun_out/r01_train_launches.csv python tools/gpu_train_step.py 2 > gpurun_out/ncu_t1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 40 -c 1 -o gpurun_out/r01_wgrad python tools/gpu_train_step.py 1 > gpurun_out/ncu_t2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_triplet_gather_bwd -s 0 -c 1 -o gpurun_out/r01_gather_bwd python tools/gpu_train_step.py 1 > gpurun_out/ncu_t3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:triplet_basis_project_bwd_kernel -s 0 -c 1 -o gpurun_out/r01_project_bwd python tools/gpu_train_step.py 1 > gpurun_out/ncu_t4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:linear_tiled_kernel -s 30 -c 1 -o gpurun_out/r01_linear_tiled python tools/gpu_train_step.py 1 > gpurun_out/ncu_t5.log 2>&1
tail -2 gpurun_out/ncu_t1.log gpurun_out/ncu_t2.log gpurun_out/ncu_t5.log
ls -la gpurun_out/*.ncu-rep | tail -6
