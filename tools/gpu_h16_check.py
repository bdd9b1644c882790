"""Second-generation dense chain (two tiles in flight, 3xFP16) on a GPU box: per-kernel agreement with the 3xTF32
chain and the exact-fp32 twin, full-model energy error vs the oracle, per-kernel timings (test infrastructure)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict, rel_err
from oracle import restated
from dig_b200 import ops, _lib
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet, DimeNetPP

dev = torch.device("cuda:0")
r = lambda a, b: rel_err(a.cpu().numpy(), b.cpu().numpy())

# ---- kernel level, ragged + odd tile count
model = SphereNet(); model.load_state_dict(formula_state_dict(model.state_dict(), seed=2)); model = model.to(dev)
for nmol in (1, 11, 24, 128):
    b = synthetic_batch(nmol, "qm9", seed=2, variable=(nmol != 128)).to(dev)
    g = ops.build_graph(b.pos, b.batch, 5.0, num_graphs=nmol)
    ops.triplet_geometry(g, b.pos, use_torsion=True, want_idx=False)
    rbf0, bess = ops.edge_basis(g.dist, 5.0, 5, model.emb.dist_emb.freq, 0, False, 6, 42)
    w_s, w_t = model._projection_rows(0, 4)
    sbf_p, t_p = ops.triplet_basis_project(g, bess, 0, w_s, w_t)
    e1_s, v_s = ops.sphere_init_e(b.z, g, rbf0, ops.pack_init_e(model.init_e), 128)
    cache = {}
    pk = ops.tc_pack_matrix(model.init_e.lin.weight, cache, "k", kind="h16")
    e1_h, v_h = ops.sphere_init_e_h16(b.z, g, rbf0, ops.pack_init_e(model.init_e), pk, 128)
    torch.cuda.synchronize()
    print(f"nmol {nmol:4d} E {g.n_edges:6d} tiles {(g.n_edges + 127) // 128:4d} | init_e: e1 {r(e1_h, e1_s):.2e} v {r(v_h, v_s):.2e}", end=" | ", flush=True)
    ue = model.update_es[1]
    e_ref, v_ref = ops.sphere_update_e(e1_s, g, rbf0, sbf_p, t_p, 8, ops.pack_update_e(ue, True), 128, 64)
    e_t, v_t, xji_t, xd_t = ops.sphere_update_e_tc(e1_s, g, rbf0, sbf_p, t_p, 8, ops.tc_pack_update_e(ue, True, cache), 128, 64)
    e_h, v_hh, xji_h, xd_h = ops.sphere_update_e_h16(e1_s, g, rbf0, sbf_p, t_p, 8, ops.tc_pack_update_e(ue, True, cache, kind="h16"), 128, 64)
    torch.cuda.synchronize()
    print(f"A: x_ji {r(xji_h, xji_t):.2e} x_down {r(xd_h, xd_t):.2e} | B vs simt: e1 {r(e_h, e_ref):.2e} v {r(v_hh, v_ref):.2e}"
          f" (tc: e1 {r(e_t, e_ref):.2e} v {r(v_t, v_ref):.2e})", flush=True)
print("timeouts", ops.tc_timeouts(), "overflow", ops.h16_overflow())

# ---- full-model energy error vs the oracle on the same GPU
for name, cls, tors, nmol in (("spherenet", SphereNet, True, 128), ("dimenetpp", DimeNetPP, False, 64)):
    m = cls(); sd = formula_state_dict(m.state_dict(), seed=2); m.load_state_dict(sd); m = m.to(dev)
    b = synthetic_batch(nmol, "qm9", seed=2).to(dev)
    with torch.no_grad():
        ref = restated.dimenet_family_forward({k: v.to(dev) for k, v in sd.items()}, b.z, b.pos, b.batch, torsion=tors)
    for dense in ("simt", "tc", "h16"):
        os.environ["DIG3D_DENSE"] = dense
        for fast in (1, 0):
            ops.h16_set_fast_swish(fast); ops.tc_set_fast_swish(fast)
            with torch.no_grad():
                u = m(b)
            print(f"{name} b={nmol} {dense:5s} fast_swish={fast} rel(energy, oracle cuda) = {r(u, ref):.3e}", flush=True)
ops.h16_set_fast_swish(1); ops.tc_set_fast_swish(1)

# ---- per-kernel timing at the headline size
model = SphereNet(); model.load_state_dict(formula_state_dict(model.state_dict(), seed=2)); model = model.to(dev)
b = synthetic_batch(128, "qm9", seed=2).to(dev)
for dense in ("tc", "h16"):
    os.environ["DIG3D_DENSE"] = dense
    with torch.no_grad():
        for _ in range(3): model(b)
        _lib.start_timing()
        for _ in range(10): model(b)
        per = _lib.stop_timing()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): model(b)
        e.record(); torch.cuda.synchronize()
    print(dense, "step ms", round(a.elapsed_time(e) / 20, 4), {k.replace("dig3d_", ""): round(sum(v) / len(v), 4) for k, v in sorted(per.items())}, flush=True)
print("timeouts", ops.tc_timeouts(), "overflow", ops.h16_overflow())
