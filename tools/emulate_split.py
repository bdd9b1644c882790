"""CPU emulation of operand splits for the tensor-core dense chain (test infrastructure).

Replaces the wide linears of the oracle's SphereNet forward by an emulated split product
(partial products in fp64, rounded once to fp32) and reports the energy error against the
plain fp32 oracle -- isolates the error of the SPLIT itself (not of the TMEM accumulation).

    python tools/emulate_split.py [n_molecules]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import restated
from oracle.weights import formula_state_dict
from dig_b200.data import synthetic_batch
from dig_b200.threedgraph.method import SphereNet


def rna_tf32(x):
    i = x.view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)


def split_tf32(x):
    hi = rna_tf32(x)
    lo = rna_tf32(x - hi)
    return hi.double(), lo.double()


def split_f16(x, scale):
    xs = x * scale
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi.double() / scale, lo.double() / scale


def split_f16_bf16(x):
    hi = x.half()
    lo = (x - hi.float()).bfloat16()
    return hi.double(), lo.double()


MODE = "fp32"
SA, SW = 16.0, 64.0
WIDE = ("lin_ji", "lin_kj", "lin_down", "lin_up", ".lin1", ".lin2", ".lin")
STATS = {}


def emu_lin(sd, name, x):
    w, b = sd[name + ".weight"], sd.get(name + ".bias")
    wide = w.size(1) >= 64 and name.startswith(("update_es", "init_e.lin")) and not name.endswith("lin_rbf_0")
    if MODE == "fp32" or not wide:
        return F.linear(x, w, b)
    if MODE == "tf32x3":
        ah, al = split_tf32(x); wh, wl = split_tf32(w)
    elif MODE == "f16x3s":
        ah, al = split_f16(x, SA); wh, wl = split_f16(w, SW)
        STATS["amax"] = max(STATS.get("amax", 0.0), float(x.abs().max()))
    elif MODE == "f16bf16":
        ah, al = split_f16_bf16(x); wh, wl = split_f16_bf16(w)
    elif MODE == "tf32x1":
        ah, al = split_tf32(x); wh, wl = split_tf32(w); al = al * 0; wl = wl * 0
    y = (ah @ wh.t() + al @ wh.t() + ah @ wl.t()).float()
    return y + b if b is not None else y


def main():
    global MODE
    nmol = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    torch.manual_seed(0)
    model = SphereNet()
    sd = formula_state_dict(model.state_dict(), seed=7)
    batch = synthetic_batch(nmol, "qm9", seed=2)
    restated._lin = emu_lin
    ref64 = restated.spherenet_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()},
                                       batch.z, batch.pos.double(), batch.batch)
    out = {}
    for MODE in ("fp32", "tf32x3", "f16x3s", "f16bf16"):
        out[MODE] = restated.spherenet_forward(sd, batch.z, batch.pos, batch.batch)
    base = out["fp32"]
    den = base.abs().max()
    print("fp32 vs fp64 (geometry noise incl.):", float((base - ref64.float()).abs().max() / den))
    for k, v in out.items():
        if k != "fp32":
            print(f"{k:8s} vs fp32 oracle: max rel {float((v - base).abs().max() / den):.3e}")
    print("max |activation| into a wide linear:", STATS.get("amax"))


if __name__ == "__main__":
    main()
