#!/bin/bash
# round 2, call V: warp-per-node gather (splits), warp-per-node triplet_count / edge_fill; parity + timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_train.py -x -q -k "not force_training and not headline_size" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2v_pytest.log
timeout 600 python tools/gpu_gather_modes.py > gpurun_out/r2v_gather_modes.log 2>&1; echo "gather modes rc=$?"; cat gpurun_out/r2v_gather_modes.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2v_bench_quick.json 2> gpurun_out/r2v_bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r2v_bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r2v_bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial'], 'e2e', d['e2e']['value'], 'in flight', d.get('batches_in_flight'))
print(d['roofline']['kernel_ms'])
"
