python bench.py --steps 300 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_r01_g.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r01_g.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['kernel_ms']); print({k:d['roofline'][k] for k in ('achieved','frac','ms_per_launch','share_of_step_kernel_time')}); print(d['cpu_baseline']); print(d['gpu_comparator']); print(d['clocks']); print(d['scatter_roofline'])"
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 900 --csv --log-file gpurun_out/r01_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_update_e_b_tc -s 2 -c 1 -o gpurun_out/r01_b_tc_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_update_v_batched -s 0 -c 1 -o gpurun_out/r01_update_v_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
