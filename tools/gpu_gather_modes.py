"""Triplet gather organisations at the headline size: energy error vs the oracle and per-kernel time for
ops.GATHER_MODE in (node, warp x split) -- test infrastructure."""
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
for name, cls, tors in (("spherenet", SphereNet, True), ("dimenetpp", DimeNetPP, False)):
    model = cls(); sd = formula_state_dict(model.state_dict(), seed=2); model.load_state_dict(sd); model = model.to(dev)
    b = synthetic_batch(128, "qm9", seed=2).to(dev)
    with torch.no_grad():
        ref = restated.dimenet_family_forward({k: v.to(dev) for k, v in sd.items()}, b.z, b.pos, b.batch, torsion=tors)
    for mode, split in (("node", None), ("warp", 1), ("warp", 2), ("warp", 4)):
        ops.GATHER_MODE[0], ops.GATHER_SPLIT[0] = mode, split
        with torch.no_grad():
            for _ in range(3): u = model(b)
            _lib.start_timing()
            for _ in range(10): model(b)
            per = _lib.stop_timing()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20): model(b)
            e.record(); torch.cuda.synchronize()
        gk = {k.replace("dig3d_", ""): round(sum(v) / len(v), 4) for k, v in per.items() if "gather" in k}
        print(f"{name} gather={mode} split={split}: step {a.elapsed_time(e) / 20:.4f} ms, {gk}, rel(energy, oracle) "
              f"{rel_err(u.cpu().numpy(), ref.cpu().numpy()):.3e}, timeouts {ops.tc_timeouts()}", flush=True)
ops.GATHER_MODE[0], ops.GATHER_SPLIT[0] = "warp", None
