"""Per-kernel time table of one SphereNet TRAINING step (fwd + bwd), bucketed by entry point and shape.
    python tools/gpu_train_profile.py [nmol]"""
import collections
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dig_b200 import _lib, ops  # noqa: E402
from dig_b200.data import synthetic_batch  # noqa: E402
from dig_b200.threedgraph.method import SphereNet  # noqa: E402

nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SphereNet().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4)
b = synthetic_batch(nmol, "qm9", seed=1).to(dev)
y = torch.randn(nmol, 1, device=dev)
events = []
orig = _lib.call


def tagged(name, *args):
    tag = name.replace("dig3d_", "")
    v = [a for a in args]
    if name == "dig3d_linear":
        tag += f"[{v[1]}x{v[2]}->{v[3]}]"
    elif name == "dig3d_wgrad":
        tag += f"[{v[2]}:{v[3]}x{v[4]}]"
    elif name in ("dig3d_act", "dig3d_act_bwd"):
        tag += f"[{v[1] if name == 'dig3d_act' else v[2]}]"
    elif name == "dig3d_ewise":
        tag += f"[{v[2]},op{v[3]}]"
    elif name in ("dig3d_gather_rows", "dig3d_scatter_add_rows"):
        tag += f"[{v[3]}x{v[4]}]"
    elif name == "dig3d_segment_sum":
        tag += f"[{v[2]}segs x{v[3]}]"
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    orig(name, *args)
    e.record()
    events.append((tag, a, e))


def step():
    opt.zero_grad()
    out = model(b)
    loss = torch.nn.functional.l1_loss(out, y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"untimed-hook step: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms  ({nmol} molecules)")
ops.call = tagged
step()
torch.cuda.synchronize()
ops.call = orig
tot = collections.defaultdict(lambda: [0.0, 0])
for tag, a, e in events:
    tot[tag][0] += a.elapsed_time(e)
    tot[tag][1] += 1
total = sum(v[0] for v in tot.values())
print(f"sum of kernel times: {total:.3f} ms over {len(events)} launches")
for tag, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{ms:9.3f} ms  {n:4d}x  {tag}")
