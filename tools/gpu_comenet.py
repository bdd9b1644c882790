"""ComENet inference at the BASELINE configs[3] size (64 OC20-IS2RE-shape structures): tensor-engine forward vs round 1's
fused FFMA block kernel -- time per forward, per-entry-point breakdown, energy error vs the oracle (test infrastructure)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import formula_state_dict, rel_err
from oracle import restated
from dig_b200 import ops, _lib
from dig_b200.data import synthetic_batch
from dig_b200.pipeline import InferencePipeline
from dig_b200.threedgraph.method import ComENet

dev = torch.device("cuda:0")
model = ComENet(cutoff=6.0); sd = formula_state_dict(model.state_dict(), seed=9); model.load_state_dict(sd); model = model.to(dev).eval()
b = synthetic_batch(64, "oc20-is2re", seed=4).to(dev)
with torch.no_grad():
    ref = restated.comenet_forward({k: v.to(dev) for k, v in sd.items()}, b.z, b.pos, b.batch, cutoff=6.0)
for mode in ("simt", "h16"):
    os.environ["DIG3D_COMENET_DENSE"] = mode
    with torch.no_grad():
        for _ in range(3): u = model(b)
        _lib.start_timing()
        for _ in range(5): model(b)
        per = _lib.stop_timing()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); l0 = _lib.launch_count
        a.record()
        for _ in range(20): model(b)
        e.record(); torch.cuda.synchronize()
    tot = {k.replace("dig3d_", ""): (round(sum(v) / 5, 4), len(v) // 5) for k, v in per.items()}
    print(f"comenet {mode}: {a.elapsed_time(e) / 20:.4f} ms/forward, {(_lib.launch_count - l0) // 20} launches, "
          f"rel(energy, oracle) {rel_err(u.cpu().numpy(), ref.cpu().numpy()):.3e}", flush=True)
    print("   ", dict(sorted(tot.items(), key=lambda kv: -kv[1][0])), flush=True)
    hb = [b.pin_memory() if False else b for _ in range(4)]
    pipe = InferencePipeline(model, dev)
    for _ in pipe.map(hb * 2): pass
    torch.cuda.synchronize(); a.record()
    n = 0
    for _ in pipe.map(hb * 6): n += 1
    cur = torch.cuda.current_stream()
    for st in pipe.streams: cur.wait_stream(st)
    e.record(); torch.cuda.synchronize()
    print(f"    {pipe.depth} batches in flight: {a.elapsed_time(e) / n:.4f} ms/forward", flush=True)
print("timeouts", ops.tc_timeouts(), "overflow", ops.h16_overflow())
