#!/bin/bash
# round 2, call Q: split edge basis, leaner wgrad producer, ComENet invalidation + fused residual epilogues; quick bench + train step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -x -q -k "edge_basis or basis_bit_exact or comenet or weight_gradient or energy_parity or two_tile_engine or baseline_configs" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2q_pytest.log
timeout 600 python tools/gpu_comenet.py > gpurun_out/r2q_comenet.log 2>&1; echo "comenet rc=$?"; tail -8 gpurun_out/r2q_comenet.log | cut -c1-900
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2q_bench_quick.json 2> gpurun_out/r2q_bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r2q_bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r2q_bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial'], 'e2e', d['e2e']['value'], 'in flight', d.get('batches_in_flight'))
print(d['roofline']['kernel_ms'])
"
DIG3D_TRAIN_DENSE=mixed timeout 600 python tools/gpu_train_step.py 12 > gpurun_out/r2q_train_mixed.log 2>&1; echo "train rc=$?"; grep "^mode" -A1 gpurun_out/r2q_train_mixed.log | cut -c1-500
