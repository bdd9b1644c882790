#!/bin/bash
# round 2, call AE: init_e with the embedding panels folded into tables (one job per tile), fewer memsets in the graph build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2ae_pytest.log
for f in 1 0; do
DIG3D_INIT_TABLES=$f timeout 600 python bench.py --steps 100 --warmup 5 --quick > gpurun_out/r2ae_bench_tab$f.json 2> gpurun_out/r2ae_bench_tab$f.err; echo "bench tables=$f rc=$?"; tail -2 gpurun_out/r2ae_bench_tab$f.err
python -c "
import json
d=json.loads(open('gpurun_out/r2ae_bench_tab$f.json').read().strip().splitlines()[-1])
print('tables=$f value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'])
print(d['roofline']['per_step_ms'])
"
done
