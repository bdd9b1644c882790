"""Tiny run of the two-tile fp16 chain (for compute-sanitizer): 3 molecules, every h16 kernel once."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict, rel_err
from oracle import restated
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet
dev = torch.device("cuda:0")
m = SphereNet(); sd = formula_state_dict(m.state_dict(), seed=2); m.load_state_dict(sd); m = m.to(dev)
b = synthetic_batch(3, "qm9", seed=2).to(dev)
with torch.no_grad():
    u = m(b)
    ref = restated.spherenet_forward({k: v.to(dev) for k, v in sd.items()}, b.z, b.pos, b.batch)
torch.cuda.synchronize()
print("h16 small: rel", rel_err(u.cpu().numpy(), ref.cpu().numpy()))
