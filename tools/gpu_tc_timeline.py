"""Timeline of CTA 0 of sphere_update_e_a_tc_kernel (clock64 probes) -- test infrastructure."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict
from dig_b200 import ops, _lib
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet
dev = torch.device("cuda:0")
model = SphereNet(); model.load_state_dict(formula_state_dict(model.state_dict(), seed=2)); model = model.to(dev)
b = synthetic_batch(128, "qm9", seed=2).to(dev)
g = ops.build_graph(b.pos, b.batch, 5.0, num_graphs=128)
ops.triplet_geometry(g, b.pos, use_torsion=True, want_idx=False)
rbf0, bess = ops.edge_basis(g.dist, 5.0, 5, model.emb.dist_emb.freq, 0, False, 6, 42)
e1, _ = ops.sphere_init_e(b.z, g, rbf0, ops.pack_init_e(model.init_e), 128)
_cache = {}
wt = ops.tc_pack_update_e(model.update_es[0], True, _cache)
x_ji = torch.empty(g.n_edges, 128, device=dev); x_down = torch.empty(g.n_edges, 64, device=dev)
lib = _lib.load()
for _ in range(3):
    _lib.call("dig3d_sphere_update_e_a_tc", ops._p(e1), ops._p(rbf0), g.n_edges, ctypes.byref(wt), ops._p(x_ji), ops._p(x_down), ops._stream())
torch.cuda.synchronize()
lib.dig3d_tc_trace(1, None)
_lib.call("dig3d_sphere_update_e_a_tc", ops._p(e1), ops._p(rbf0), g.n_edges, ctypes.byref(wt), ops._p(x_ji), ops._p(x_down), ops._stream())
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
lib.dig3d_tc_trace(0, buf)
t = list(buf); t0 = t[0]
names = {0: "setup done (before sync)", 1: "after sync", 2: "A0 written", 3: "D0 ready (epi)", 4: "epi0 done", 5: "D1 ready", 6: "epi1 done",
         7: "D2 ready", 8: "mma: A ready g0", 9: "mma: W chunk0 g0", 10: "mma: issued g0", 12: "mma: A ready g1", 13: "mma: W0 g1",
         14: "mma: issued g1", 16: "mma: A ready g2", 17: "mma: W0 g2", 18: "mma: issued g2"}
for k in sorted(names):
    print(f"{names[k]:28s} +{t[k] - t0:8d} cycles")
print("producer chunk issue times:", [t[40 + i] - t0 for i in range(10)])
