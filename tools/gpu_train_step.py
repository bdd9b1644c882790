"""N SphereNet training steps (128 QM9-shape molecules): flat parameters + fused Adam, timed with CUDA events and with
the per-entry-point breakdown.  python tools/gpu_train_step.py [steps]   (DIG3D_TRAIN_DENSE=h16|tc|simt)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dig_b200 import _lib, parallel  # noqa: E402
from dig_b200.data import synthetic_batch  # noqa: E402
from dig_b200.threedgraph.method import SphereNet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
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
    return loss


for _ in range(3):
    step()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
l0 = _lib.launch_count
a.record()
for _ in range(steps):
    loss = step()
e.record()
torch.cuda.synchronize()
print(f"mode {os.environ.get('DIG3D_TRAIN_DENSE', 'h16')}: {a.elapsed_time(e) / steps:.3f} ms/step, "
      f"{(_lib.launch_count - l0) // steps} launches/step, loss {float(loss):.5f}")
_lib.start_timing()
step()
per = _lib.stop_timing()
tot = {k.replace("dig3d_", ""): (round(sum(v), 3), len(v)) for k, v in per.items()}
print(dict(sorted(tot.items(), key=lambda kv: -kv[1][0])))
