"""ComENet stage-by-stage diagnostic on a GPU box (test infrastructure)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import case_inputs, formula_state_dict, rel_err
from oracle import restated
from dig_b200 import ops
from dig_b200.threedgraph.method import ComENet

dev = torch.device("cuda:0")
g, z, pos, batch = case_inputs("comenet_oc20", dev)
model = ComENet(cutoff=6.0)
sd = formula_state_dict(model.state_dict(), seed=4)
model.load_state_dict(sd); model = model.to(dev)
sdd = {k: v.to(dev) for k, v in sd.items()}
u_ref, it = restated.comenet_forward(sdd, z, pos, batch, cutoff=6.0, return_intermediates=True)
gr = ops.build_graph(pos, batch, 6.0)
f1, f2, ang = ops.comenet_geometry(gr, pos, 6.0, want_angles=True)
print("E", gr.n_edges, "edge_index", torch.equal(gr.edge_index, it["edge_index"]))
for name, a, b in (("dist", gr.dist, it["dist"]), ("theta", ang[:, 0], it["theta"]), ("phi", ang[:, 1], it["phi"]),
                   ("tau", ang[:, 2], it["tau"]), ("f1", f1, it["feature1"]), ("f2", f2, it["feature2"])):
    print(name, "bit-equal", torch.equal(a, b), "mismatch", int((a != b).sum()), "/", a.numel(),
          "maxabs", float((a - b).abs().max()))
class B: pass
b = B(); b.z, b.pos, b.batch = z, pos, batch
with torch.no_grad():
    u = model(b)
print("mine", u.flatten().tolist(), "oracle(cuda)", u_ref.flatten().tolist(), "golden f32", g["energy_f32"].ravel().tolist(),
      "golden f64", g["energy_f64"].ravel().tolist())
print("rel(mine, oracle cuda)", rel_err(u.cpu().numpy(), u_ref.cpu().numpy()))
