#!/bin/bash
# round 2, call C: leaner epilogue (ftz math, prefetched skip rows), MMA issue-rate microbenchmark, new harness
mkdir -p gpurun_out
timeout 120 ./tools/mma_rate > gpurun_out/r2c_mma_rate.log 2>&1; echo "mma_rate rc=$?"; cat gpurun_out/r2c_mma_rate.log
timeout 300 python tools/gpu_h16_timeline.py > gpurun_out/r2c_timeline.log 2>&1; echo "timeline rc=$?"; cat gpurun_out/r2c_timeline.log | tail -48
timeout 600 python tools/gpu_h16_check.py > gpurun_out/r2c_h16_check.log 2>&1; echo "h16_check rc=$?"; tail -22 gpurun_out/r2c_h16_check.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_train.py -x -q -k "two_tile or headline or out_of_range or comenet or indices or flat_adam or node_centred" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2c_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2c_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','windows','e2e','gpu_launches','clocks','parity','cpu_baseline','gpu_comparator','train'): print(k, d.get(k))
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('kernel_ms','per_step_ms')})
print(d['roofline']['per_step_ms'])
"
