#!/bin/bash
# round 2, call F: startup overlap in chain B, new coverage tests (ComENet-OCP, G-SphereNet geometry, node features), bench
mkdir -p gpurun_out
timeout 300 python tools/gpu_h16_timeline.py > gpurun_out/r2f_timeline.log 2>&1; echo "timeline rc=$?"; head -4 gpurun_out/r2f_timeline.log
timeout 1500 python -m pytest tests -x -q -m gpu -k "ocp or gspherenet or node_feature or two_tile or headline or flat_adam" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2f_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','windows','e2e','gpu_launches','parity','train'): print(k, d.get(k))
print(d['roofline']['per_step_ms'])
for o in d['other_configs']: print(o)
"
