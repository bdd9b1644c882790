#!/bin/bash
# round 2, call N: full GPU suite + smoke + full bench (final numbers of the round)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2n_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2n_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','serial','windows','e2e','gpu_launches','parity','train','roofline'): print(k, d.get(k))
for o in d.get('other_configs', []): print(o)
"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2n_bench_ref.json
