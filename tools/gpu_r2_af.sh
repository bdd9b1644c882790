#!/bin/bash
# round 2, call AF: full bench run P (final state of the round)
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2af_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2af_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'serial', d['serial']['ms_per_step'], 'e2e', d['e2e']['value'], 'in flight', d.get('batches_in_flight'), 'parity', d['parity']['rel_err_max'])
print('roofline', d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['ms_per_launch'])
print('train', d['train']['ms_per_step'], 'gpu_comp', d['gpu_comparator']['value'], 'cpu', d['cpu_baseline']['value'])
for c in d['other_configs']:
    print(c['config'][:40], {k: (v.get('ms_per_step'), v.get('in_flight', {}).get('ms_per_step') if isinstance(v.get('in_flight'), dict) else None) for k, v in c.items() if isinstance(v, dict)})
"
