#!/bin/bash
# round 2, call AJ (evidence, final state): ncu --set full of the wide-epilogue fused chain, launch list of the bench command
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_x16_kernel -s 4 -c 1 -o gpurun_out/r02_ba_x16 python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2aj_ncu_ba.log 2>&1; echo "ncu ba rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_q.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2aj_ncu_bench.log 2>&1; echo "ncu bench rc=$?"
ls -la gpurun_out/r02_ba_x16.ncu-rep gpurun_out/r02_launches_q.csv
