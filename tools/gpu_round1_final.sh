# round-1 closing run on 2 GPUs: the N=2 bench line (inference weak scaling + data-parallel training step with the NCCL
# gradient all-reduce).  Run with: gpurun --gpus 2 -- bash tools/gpu_round1_final.sh
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_k_n2.json 2> gpurun_out/bench_k_n2.err
echo rc=$?
tail -2 gpurun_out/bench_k_n2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_k_n2.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value']); print(d['train'])"
