#!/bin/bash
# round 2, call I (2 GPUs): weak scaling of inference and of the data-parallel training step (overlapped flat-gradient all-reduce)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err; echo "bench n2 rc=$?"; tail -c 600 gpurun_out/r2i_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','windows','e2e','train'): print(k, d.get(k))
"
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench n1 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2i_bench_n1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','windows','e2e'): print(k, d.get(k))
"
