timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_a_tc -s 2 -c 1 -o gpurun_out/r01_a_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_tc -s 2 -c 1 -o gpurun_out/r01_b_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out | tail -5
