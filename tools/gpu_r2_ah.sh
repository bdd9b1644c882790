#!/bin/bash
# round 2, call AH: wide epilogue (sixteen warps on the ready tile) for update_e part B (+ A): parity + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide_epilogue or lean or headline or energy_parity or fp16_chain" > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2ah_pytest.log
for f in 1 0; do
DIG3D_H16_WIDE=$f timeout 600 python bench.py --steps 100 --warmup 5 --quick > gpurun_out/r2ah_bench_wide$f.json 2> gpurun_out/r2ah_bench_wide$f.err; echo "bench wide=$f rc=$?"; tail -2 gpurun_out/r2ah_bench_wide$f.err
python -c "
import json
d=json.loads(open('gpurun_out/r2ah_bench_wide$f.json').read().strip().splitlines()[-1])
print('wide=$f value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'])
print(d['roofline']['kernel_ms'])
"
done
