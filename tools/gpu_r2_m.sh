#!/bin/bash
# round 2, call M: batches in flight (1..4 streams), prefetching tensor-core wgrad, training step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "weight_gradient" > gpurun_out/r2m_pytest_a.log 2>&1; echo "pytest a rc=$?"; tail -5 gpurun_out/r2m_pytest_a.log
for d in 1 2 3 4; do
DIG3D_BENCH_STREAMS=$d timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2m_bench_s$d.json 2> gpurun_out/r2m_bench_s$d.err; echo "bench streams=$d rc=$?"; tail -2 gpurun_out/r2m_bench_s$d.err
python -c "
import json
d=json.loads(open('gpurun_out/r2m_bench_s$d.json').read().strip().splitlines()[-1])
print('value', d.get('value'), 'ms', d.get('ms_per_step'), 'serial', d.get('serial',{}).get('value'), 'e2e', d['e2e']['value'], 'e2e serial', d['e2e']['serial']['value'], 'parity', d.get('parity',{}).get('rel_err_max'))
"
done
DIG3D_TRAIN_DENSE=mixed timeout 600 python tools/gpu_train_step.py 12 > gpurun_out/r2m_train_mixed.log 2>&1; echo "train rc=$?"; grep "^mode" -A1 gpurun_out/r2m_train_mixed.log | cut -c1-700
