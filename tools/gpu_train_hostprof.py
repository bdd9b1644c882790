"""Where the HOST time of a SphereNet training step goes (cProfile over 6 steps, backward forced onto the calling thread so
that its Python callbacks are seen) -- test infrastructure.  python tools/gpu_train_hostprof.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dig_b200 import _lib, parallel  # noqa: E402
from dig_b200.data import synthetic_batch  # noqa: E402
from dig_b200.threedgraph.method import SphereNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SphereNet().to(dev)
flat = parallel.FlatParameters(model)
opt = parallel.FlatAdam(flat, lr=5e-4)
b = synthetic_batch(128, "qm9", seed=1).to(dev)
y = torch.randn(128, 1, device=dev)


def step():
    opt.zero_grad()
    loss = torch.nn.functional.l1_loss(model(b), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
# host time per step WITHOUT waiting for the GPU (launch-issue time): the queue never fills in 6 steps
t0 = time.perf_counter()
for _ in range(6):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time {1e3 * (t1 - t0) / 6:.3f} ms/step; with final sync {1e3 * (t2 - t0) / 6:.3f} ms/step")
with torch.autograd.set_multithreading_enabled(False):
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(6):
        step()
    pr.disable()
    torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(22)
