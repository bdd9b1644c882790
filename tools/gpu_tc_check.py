"""tcgen05 update_e vs the exact-fp32 SIMT update_e on a GPU box (test infrastructure)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict, case_inputs, CASES, rel_err
from oracle import restated
from dig_b200 import ops, _lib
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet, DimeNetPP

dev = torch.device("cuda:0")
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))

# ---- full-model energy error vs the oracle on the same GPU, three dense variants
for name, cls, tors in (("spherenet_qm9", SphereNet, True), ("dimenetpp_md17", DimeNetPP, False)):
    g, z, pos, batch = case_inputs(name, dev)
    kw = CASES[name][1]
    model = cls(**kw)
    sd = formula_state_dict(model.state_dict(), seed=CASES[name][3])
    model.load_state_dict(sd); model = model.to(dev)
    u_ref = restated.dimenet_family_forward({k: v.to(dev) for k, v in sd.items()}, z, pos, batch, torsion=tors, cutoff=5.0)
    class B: pass
    b = B(); b.z, b.pos, b.batch = z, pos, batch
    for variant in ("simt", "tc", "tc-fast"):
        os.environ["DIG3D_DENSE"] = "simt" if variant == "simt" else "tc"
        ops.tc_set_fast_swish(variant == "tc-fast")
        with torch.no_grad():
            u = model(b)
        print(f"{name} {variant:8s} rel(energy, oracle cuda) = {rel_err(u.cpu().numpy(), u_ref.cpu().numpy()):.3e}", flush=True)
ops.tc_set_fast_swish(False)
os.environ["DIG3D_DENSE"] = "tc"

# ---- per-kernel timing at the headline size
model = SphereNet(); model.load_state_dict(formula_state_dict(model.state_dict(), seed=2)); model = model.to(dev)
b = synthetic_batch(128, "qm9", seed=2).to(dev)
for fast in (0, 1):
    ops.tc_set_fast_swish(fast)
    with torch.no_grad():
        for _ in range(3): model(b)
        _lib.start_timing()
        for _ in range(10): model(b)
        per = _lib.stop_timing()
    print("fast_swish", fast, {k.replace("dig3d_", ""): round(sum(v) / len(v), 4) for k, v in sorted(per.items())}, flush=True)
print("timeouts", ops.tc_timeouts())
