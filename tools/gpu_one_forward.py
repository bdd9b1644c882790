"""A few SphereNet headline-size forwards (for ncu captures) -- test infrastructure."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
model = SphereNet(); model.load_state_dict(formula_state_dict(model.state_dict(), seed=2)); model = model.to(dev)
b = synthetic_batch(128, "qm9", seed=2).to(dev)
with torch.no_grad():
    for _ in range(n):
        u = model(b)
torch.cuda.synchronize()
print("done", float(u.sum()))
