#!/bin/bash
# round 2, call Y: ncu --set full of the warp-per-node gather and the triplet geometry kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sphere_triplet_gather_warp -s 6 -c 1 -o gpurun_out/r02_gather_warp python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2y_ncu_gather.log 2>&1; echo "ncu gather rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triplet_geometry -s 2 -c 1 -o gpurun_out/r02_triplet_geometry python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r2y_ncu_geom.log 2>&1; echo "ncu geometry rc=$?"
ls -la gpurun_out/*.ncu-rep
