#!/bin/bash
# round 2, call X: lean inference path (cached plan): bit-identity test, host profile, quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q -k "lean or deterministic or energy_parity or headline or pipeline or baseline_configs" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2x_pytest.log
timeout 600 python tools/gpu_infer_hostprof.py > gpurun_out/r2x_infer_hostprof.log 2>&1; echo "hostprof rc=$?"; head -45 gpurun_out/r2x_infer_hostprof.log | cut -c1-180
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2x_bench_quick.json 2> gpurun_out/r2x_bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r2x_bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r2x_bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial'], 'e2e', d['e2e'], 'in flight', d.get('batches_in_flight'))
print(d['roofline']['kernel_ms'])
"
