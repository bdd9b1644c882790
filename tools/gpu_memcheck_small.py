"""Small end-to-end pass over the training / force / ProNet kernels for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/gpu_memcheck_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dig_b200.data import synthetic_batch, synthetic_proteins  # noqa: E402
from dig_b200.threedgraph.method import ComENet, DimeNetPP, ProNet, SchNet, SphereNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
b = synthetic_batch(3, "qm9", seed=3, variable=True).to(dev)
for make in (lambda: SchNet(energy_and_force=True, num_layers=2, hidden_channels=32, num_filters=32, cutoff=5.0),
             lambda: DimeNetPP(energy_and_force=True, num_layers=2), lambda: SphereNet(energy_and_force=True, num_layers=2),
             lambda: ComENet(cutoff=5.0, num_layers=2)):
    m = make().to(dev)
    b.pos = b.pos.detach().clone()
    out = m(b)
    if getattr(m, "energy_and_force", False):
        f = torch.autograd.grad(out, b.pos, grad_outputs=torch.ones_like(out), create_graph=True)[0]
        (out.abs().mean() + (f.abs().mean() if f.requires_grad else 0.0)).backward()
    else:
        out.abs().mean().backward()
    with torch.no_grad():
        m(b)
    print(type(m).__name__, "ok", float(out.sum()))
p = synthetic_proteins(2, length=16, seed=1).to(dev)
m = ProNet(level="allatom", num_blocks=1).to(dev)
m(p).abs().mean().backward()
torch.cuda.synchronize()
print("ProNet ok")
