"""Where the HOST time of the inference loop goes (cProfile over the batches-in-flight pipeline and over the plain
loop; SphereNet, 128 QM9-shape molecules) -- test infrastructure.  python tools/gpu_infer_hostprof.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dig_b200.data import synthetic_batch  # noqa: E402
from dig_b200.pipeline import InferencePipeline  # noqa: E402
from dig_b200.threedgraph.method import SphereNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SphereNet().to(dev).eval()
host = [synthetic_batch(128, "qm9", seed=s).pin_memory() for s in range(8)]
pipe = InferencePipeline(model, dev)


def run(n):
    for out in pipe.map(host[i % 8] for i in range(n)):
        pass


run(24)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    run(200)
    torch.cuda.synchronize()
    print(f"pipeline (depth {pipe.depth}): {1e3 * (time.perf_counter() - t0) / 200:.4f} ms/batch wall", flush=True)
for depth in (1, 2, 3, 4, 6):
    pd = InferencePipeline(model, dev, depth=depth)
    for out in pd.map(host[i % 8] for i in range(24)):
        pass
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for out in pd.map(host[i % 8] for i in range(200)):
            pass
        torch.cuda.synchronize()
        best = min(best, 1e3 * (time.perf_counter() - t0) / 200)
    print(f"depth {depth}: {best:.4f} ms/batch wall = {128e3 / best:.0f} molecules/s", flush=True)
pr = cProfile.Profile()
pr.enable()
run(200)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
print("---- pipeline, 200 batches, by internal time")
st.sort_stats("tottime").print_stats(30)
print("---- pipeline, by cumulative time")
st.sort_stats("cumulative").print_stats(30)
