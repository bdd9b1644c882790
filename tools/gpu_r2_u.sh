#!/bin/bash
# round 2, call U: clean FFMA / FFMA2 rate, ncu --set full of the packed gather and the packed projection
mkdir -p gpurun_out
timeout 120 ./tools/ffma_rate > gpurun_out/r2u_ffma_rate.log 2>&1; echo "ffma_rate rc=$?"; cat gpurun_out/r2u_ffma_rate.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_triplet_gather_node -s 6 -c 1 -o gpurun_out/r02_gather_node_packed python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2u_ncu_gather.log 2>&1; echo "ncu gather rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triplet_basis_project_packed -s 2 -c 1 -o gpurun_out/r02_project_packed python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2u_ncu_project.log 2>&1; echo "ncu project rc=$?"
ls -la gpurun_out/*.ncu-rep
