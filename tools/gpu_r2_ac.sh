#!/bin/bash
# round 2, call AC: part A of block l+1 fused into the chain of part B of block l (two launches per block): bit-identity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "lean or headline or energy_parity or baseline_configs or fp16_chain" > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2ac_pytest.log
for f in 1 0; do
DIG3D_FUSE_BA=$f timeout 600 python bench.py --steps 100 --warmup 5 --quick > gpurun_out/r2ac_bench_fuse$f.json 2> gpurun_out/r2ac_bench_fuse$f.err; echo "bench fuse=$f rc=$?"; tail -2 gpurun_out/r2ac_bench_fuse$f.err
python -c "
import json
d=json.loads(open('gpurun_out/r2ac_bench_fuse$f.json').read().strip().splitlines()[-1])
print('fuse=$f value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity'))
print(d['roofline']['per_step_ms'])
"
done
