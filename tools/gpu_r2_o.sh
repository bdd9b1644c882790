#!/bin/bash
# round 2, call O (8 GPUs): weak scaling of inference (batches in flight) and of the data-parallel training step
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -10
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2o_bench_n$N.json 2> gpurun_out/r2o_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 800 gpurun_out/r2o_bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2o_bench_n$N.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','n_gpus','serial','windows','e2e','train'): print(k, d.get(k))
"
