#!/bin/bash
# round 2, call Z: out-edge lists of the graph build (projection / gather without the search), pipeline depth sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py tests/test_gpu_train.py -x -q -k "not force_training and not headline_size" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2z_pytest.log
timeout 600 python tools/gpu_infer_hostprof.py > gpurun_out/r2z_infer_hostprof.log 2>&1; echo "hostprof rc=$?"; head -9 gpurun_out/r2z_infer_hostprof.log | cut -c1-180
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2z_bench_quick.json 2> gpurun_out/r2z_bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'], 'in flight', d.get('batches_in_flight'))
print(d['roofline']['kernel_ms'])
"
