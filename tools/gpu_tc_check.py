"""tcgen05 update_e vs the exact-fp32 SIMT update_e on a GPU box (test infrastructure)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict
from dig_b200 import ops
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet, DimeNetPP

dev = torch.device("cuda:0")
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

for cls, tors in ((SphereNet, True), (DimeNetPP, False)):
    model = cls()
    model.load_state_dict(formula_state_dict(model.state_dict(), seed=2))
    model = model.to(dev)
    for nmol in (6, 128):
        b = synthetic_batch(nmol, "qm9", seed=2, variable=(nmol == 6)).to(dev)
        g = ops.build_graph(b.pos, b.batch, 5.0, num_graphs=nmol)
        ops.triplet_geometry(g, b.pos, use_torsion=tors, want_idx=False)
        rbf0, bess = ops.edge_basis(g.dist, 5.0, 5, model.emb.dist_emb.freq, 0, not tors, 6, 42)
        w_s, w_t = model._projection_rows(0, 4)
        sbf_p, t_p = ops.triplet_basis_project(g, bess, 0, w_s, w_t)
        e1, v0 = ops.sphere_init_e(b.z, g, rbf0, ops.pack_init_e(model.init_e), 128)
        ue = model.update_es[0]
        ws = ops.pack_update_e(ue, tors)
        cache = {}
        wt = ops.tc_pack_update_e(ue, tors, cache)
        torch.cuda.synchronize()
        # SIMT reference
        e_ref, v_ref = ops.sphere_update_e(e1, g, rbf0, sbf_p, t_p, 0, ws, 128, 64)
        torch.cuda.synchronize()
        e_tc, v_tc, xji, xd = ops.sphere_update_e_tc(e1, g, rbf0, sbf_p, t_p, 0, wt, 128, 64)
        torch.cuda.synchronize()
        print(f"{cls.__name__} nmol={nmol} E={g.n_edges} T={g.n_triplets} timeouts={ops.tc_timeouts()} "
              f"rel(e1_out)={rel(e_tc, e_ref):.3e} rel(v_in)={rel(v_tc, v_ref):.3e} finite={bool(torch.isfinite(e_tc).all())}",
              flush=True)
        if nmol == 128:
            for name, fn in (("simt", lambda: ops.sphere_update_e(e1, g, rbf0, sbf_p, t_p, 0, ws, 128, 64)),
                             ("tc", lambda: ops.sphere_update_e_tc(e1, g, rbf0, sbf_p, t_p, 0, wt, 128, 64))):
                for _ in range(3): fn()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); a.record()
                for _ in range(20): fn()
                c.record(); torch.cuda.synchronize()
                print(f"   {name}: {a.elapsed_time(c) / 20:.4f} ms per update_e", flush=True)
