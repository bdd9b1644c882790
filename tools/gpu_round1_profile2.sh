set -x
python bench.py --steps 30 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r01_e.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_e.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['kernel_ms']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_launch','share_of_step_kernel_time')}); print(d['cpu_baseline']); print(d['gpu_comparator'])"
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 700 --csv --log-file gpurun_out/r01_launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_tc -s 2 -c 1 -o gpurun_out/r01_b_tc2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_triplet_gather -s 2 -c 1 -o gpurun_out/r01_gather python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_i.log 2>&1
