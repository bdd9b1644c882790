"""Times the 128 -> 128 training linear ([34520, 128] rows) for the three tile configurations and the whole training step
under each: python tools/gpu_linear_cfg.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dig_b200 import ops  # noqa: E402
from dig_b200.data import synthetic_batch  # noqa: E402
from dig_b200.threedgraph.method import SphereNet  # noqa: E402

dev = torch.device("cuda:0")
x = torch.randn(34520, 128, device=dev)
w = torch.randn(128, 128, device=dev) / 11.3
b = torch.randn(128, device=dev)
ref = torch.nn.functional.linear(x.double(), w.double(), b.double())


def t_us(fn, n=50):
    for _ in range(5):
        fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / n * 1e3


torch.manual_seed(0)
model = SphereNet().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4)
bt = synthetic_batch(128, "qm9", seed=1).to(dev)
yt = torch.randn(128, 1, device=dev)


def step():
    opt.zero_grad()
    torch.nn.functional.l1_loss(model(bt), yt).backward()
    opt.step()


for cfg in (0, 1, 2):
    ops.linear_set_config(cfg)
    y = ops.linear(x, w, b)
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    plain = t_us(lambda: ops.linear(x, w, b))
    fused = t_us(lambda: ops.linear(x, w, b, want_act=True))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    print(f"cfg {cfg}: linear {plain:.1f} us, linear+swish {fused:.1f} us, rel err {err:.1e}, "
          f"training step {(time.perf_counter() - t0) * 100:.2f} ms")
ops.linear_set_config(1)
