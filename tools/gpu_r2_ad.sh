#!/bin/bash
# round 2, call AD (evidence): ncu --set full of the fused B+A chain, launch list of the bench command
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_h16_kernel -s 4 -c 2 -o gpurun_out/r02_ba_h16 python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2ad_ncu_ba.log 2>&1; echo "ncu ba rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_p.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r2ad_ncu_bench.log 2>&1; echo "ncu bench rc=$?"; tail -1 gpurun_out/r2ad_ncu_bench.log | cut -c1-200
ls -la gpurun_out/r02_ba_h16.ncu-rep gpurun_out/r02_launches_p.csv
