set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r01_b.json; cat gpurun_out/bench_r01_b.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms'], d['cpu_baseline'])"
# launch list (cold, serialised; shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log | cut -c1-300
# full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b -s 2 -c 2 -o gpurun_out/r01_update_e_b python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:triplet_basis_project -s 1 -c 1 -o gpurun_out/r01_project python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/
