"""N plain SphereNet training steps (128 QM9-shape molecules) for ncu: python tools/gpu_train_step.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dig_b200.data import synthetic_batch  # noqa: E402
from dig_b200.threedgraph.method import SphereNet  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SphereNet().to(dev)
opt = torch.optim.Adam(model.parameters(), lr=5e-4)
b = synthetic_batch(128, "qm9", seed=1).to(dev)
y = torch.randn(128, 1, device=dev)
for _ in range(steps):
    opt.zero_grad()
    torch.nn.functional.l1_loss(model(b), y).backward()
    opt.step()
torch.cuda.synchronize()
print("done")
