#!/bin/bash
# round 2, call T: FP32 issue-rate microbenchmark (FFMA vs FFMA2), packed triplet gather / packed projection with the
# recurrence harmonics: parity tests + quick bench
mkdir -p gpurun_out
timeout 120 ./tools/ffma_rate > gpurun_out/r2t_ffma_rate.log 2>&1; echo "ffma_rate rc=$?"; cat gpurun_out/r2t_ffma_rate.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2t_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2t_bench_quick.json 2> gpurun_out/r2t_bench_quick.err; echo "bench rc=$?"; tail -3 gpurun_out/r2t_bench_quick.err
python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench_quick.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial'], 'e2e', d['e2e']['value'], 'in flight', d.get('batches_in_flight'))
print(d['roofline']['kernel_ms'])
print('parity', d.get('parity'))
"
