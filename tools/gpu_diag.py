"""Diagnostic dump for a GPU box: compares every stage of the CUDA path against the travelling
oracle executed on the same GPU (ATen CUDA kernels) and against the CPU golden fixtures.
Writes gpurun_out/diag.txt.  Test infrastructure (imports oracle/)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CASES, case_inputs, formula_state_dict, rel_err  # noqa: E402
from oracle import restated  # noqa: E402
from dig_b200 import ops  # noqa: E402
from dig_b200.threedgraph.method import SphereNet, DimeNetPP  # noqa: E402

out_lines = []


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out_lines.append(s)


def cmp(name, mine, ref):
    mine = mine.detach().cpu().numpy()
    ref = ref.detach().cpu().numpy() if isinstance(ref, torch.Tensor) else ref
    if mine.shape != ref.shape:
        P(f"  {name}: SHAPE {mine.shape} vs {ref.shape}")
        return
    eq = np.array_equal(mine, ref)
    if mine.dtype.kind == "f":
        nbad = int((mine != ref).sum())
        P(f"  {name}: bit-equal={eq} mismatches={nbad}/{mine.size} max_abs={np.abs(mine.astype(np.float64) - ref).max():.3e} rel={rel_err(mine, ref):.3e}")
    else:
        P(f"  {name}: equal={eq}")


def main():
    dev = torch.device("cuda:0")
    P(torch.cuda.get_device_name(0), torch.__version__)
    for name, cls, tors in (("spherenet_qm9", SphereNet, True), ("dimenetpp_md17", DimeNetPP, False),
                            ("spherenet_ns3", SphereNet, True)):
        P("==", name)
        g, z, pos, batch = case_inputs(name, dev)
        kw = CASES[name][1]
        ns = kw.get("num_spherical", 7)
        model = cls(**kw)
        sd = formula_state_dict(model.state_dict(), seed=CASES[name][3])
        model.load_state_dict(sd)
        model = model.to(dev)
        sd_dev = {k: v.to(dev) for k, v in sd.items()}
        u_ref, it = restated.dimenet_family_forward(sd_dev, z, pos, batch, torsion=tors, cutoff=kw["cutoff"],
                                                    num_spherical=ns, return_intermediates=True)
        gr = ops.build_graph(pos, batch, kw["cutoff"])
        ops.triplet_geometry(gr, pos, use_torsion=tors, want_idx=True, want_idx64=True)
        P(f"  E={gr.n_edges} T={gr.n_triplets}")
        cmp("edge_index vs oracle(cuda)", gr.edge_index, it["edge_index"])
        cmp("edge_index vs golden", gr.edge_index, g["edge_index"])
        cmp("idx_kj", gr.idx_kj64, it["idx_kj"])
        cmp("idx_ji", gr.idx_ji64, it["idx_ji"])
        cmp("dist vs oracle(cuda)", gr.dist, it["dist"])
        cmp("dist vs golden(cpu)", gr.dist, g["dist"])
        cmp("angle vs oracle(cuda)", gr.angle, it["angle"])
        cmp("angle vs golden(cpu)", gr.angle, g["angle"])
        if tors:
            cmp("torsion vs oracle(cuda)", gr.torsion, it["torsion"])
            cmp("torsion vs golden(cpu)", gr.torsion, g["torsion"])
            a = gr.torsion.cpu().numpy()
            b = it["torsion"].cpu().numpy()
            P(f"  torsion ~0 fraction: mine {(a < 1e-3).mean():.4f} oracle(cuda) {(b < 1e-3).mean():.4f} "
              f"golden(cpu) {(g['torsion'] < 1e-3).mean():.4f}; coin-flip disagreements vs oracle(cuda): "
              f"{int(((a < 1e-3) != (b < 1e-3)).sum())}")
        bid = ops.BASIS_IDS[("dimenet", ns, 6)]
        rbf0, bess = ops.edge_basis(gr.dist, kw["cutoff"], 5, sd_dev["emb.dist_emb.freq"], bid, not tors, 6, ns * 6)
        cmp("rbf0 vs oracle(cuda)", rbf0, it["rbf0"])
        cmp("rbf0 vs golden(cpu)", rbf0, g["rbf0"])
        sbf, tbf = ops.triplet_basis(bess, gr.angle, gr.torsion, gr.idx_kj, bid, ns, 6, tors)
        cmp("sbf vs oracle(cuda)", sbf, it["sbf"])
        cmp("sbf[:256] vs golden(cpu)", sbf[:256], g["sbf_head"])
        if tors:
            cmp("tbf vs oracle(cuda)", tbf, it["tbf"])
            cmp("tbf[:64] vs golden(cpu)", tbf[:64], g["tbf_head"])
        # fused projection vs explicit
        L = 4
        w_s, w_t = model._projection_rows(0, L)
        sbf_p, t_p = ops.triplet_basis_project(gr, bess, bid, w_s, w_t)
        sbf_p = sbf_p.permute(1, 0, 2).reshape(-1, 32)
        t_p = t_p.permute(1, 0, 2).reshape(-1, 32) if t_p is not None else None
        cmp("sbf_p vs sbf@W (fp32 matmul)", sbf_p, it["sbf"] @ w_s.t())
        if tors:
            cmp("t_p vs tbf@W (fp32 matmul)", t_p, it["tbf"] @ w_t.t())
        b = type("B", (), {})()
        b.z, b.pos, b.batch = z, pos, batch
        with torch.no_grad():
            u = model(b)
        P("  energy mine        ", u.flatten().tolist())
        P("  energy oracle(cuda)", u_ref.flatten().tolist())
        P("  energy golden f32  ", g["energy_f32"].ravel().tolist())
        P("  energy golden f64  ", g["energy_f64"].ravel().tolist())
        P(f"  rel(mine, oracle cuda) = {rel_err(u.cpu().numpy(), u_ref.cpu().numpy()):.3e}")
        P(f"  rel(mine, golden f32)  = {rel_err(u.cpu().numpy(), g['energy_f32']):.3e}")
        P(f"  rel(oracle cuda, golden f32) = {rel_err(u_ref.cpu().numpy(), g['energy_f32']):.3e}")
        P(f"  rel(golden f32, golden f64)  = {rel_err(g['energy_f32'], g['energy_f64']):.3e}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w") as fh:
        fh.write("\n".join(out_lines) + "\n")


if __name__ == "__main__":
    main()
