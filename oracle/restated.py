"""TEST INFRASTRUCTURE ONLY -- the travelling oracle.

A functional, device-agnostic (CPU or CUDA, fp32 or fp64) restatement in plain torch ops of the
reference algorithms on the 3D-graph hot path, driven by a state_dict with the reference's key
names.  Unlike oracle/ref_loader.py it does not need /root/reference, so it runs on the GPU box:
  * on CPU it is the `cpu_baseline` / `--impl reference` leg of bench.py ("port");
  * on CUDA it executes the same op sequence through ATen's CUDA kernels, i.e. "the reference's
    own PyG/torch path run on the same GPU" -- the thing the kernels must match to 1e-5.

Pinned against the real reference (tests/test_oracle.py, run where /root/reference exists) and
against tests/golden/*.npz everywhere.  The closed-form basis strings come from
tests/golden/basis_formulas.json, written by the reference's own sympy code.

Each function cites the reference lines it restates.
"""
import json
import math
import os

import torch
import torch.nn.functional as F

from . import shim

_GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_FORMULAS = None


def formulas():
    global _FORMULAS
    if _FORMULAS is None:
        with open(os.path.join(_GOLD, "basis_formulas.json")) as fh:
            _FORMULAS = json.load(fh)
    return _FORMULAS


def _compile(src, args):
    """What sympy.lambdify produces for the reference: a python function over torch tensors."""
    ns = {"sin": torch.sin, "cos": torch.cos, "sqrt": torch.sqrt, "pi": math.pi}
    return eval(f"lambda {', '.join(args)}: {src}", ns)


def swish(x):                                   # spherenet.py:14
    return x * torch.sigmoid(x)


# ----------------------------------------------------------------------------- geometry
def radius_graph(pos, r, batch, max_num_neighbors=32):
    return shim.radius_graph(pos, r, batch, max_num_neighbors=max_num_neighbors)


def _csr(index, n):
    cnt = torch.bincount(index, minlength=n)
    ptr = torch.zeros(n + 1, dtype=torch.long, device=index.device)
    ptr[1:] = torch.cumsum(cnt, 0)
    return ptr, cnt


def _expand(ptr_start, counts):
    """For rows r with counts[r] items starting at ptr_start[r]: (row id, item index) pairs."""
    total = int(counts.sum())
    rows = torch.repeat_interleave(torch.arange(counts.numel(), device=counts.device), counts)
    first = torch.zeros(counts.numel() + 1, dtype=torch.long, device=counts.device)
    first[1:] = torch.cumsum(counts, 0)
    within = torch.arange(total, device=counts.device) - first[:-1][rows]
    return rows, ptr_start[rows] + within


def xyz_to_dat(pos, edge_index, num_nodes, use_torsion=False):
    """utils/geometric_computing.py:12-80 for an edge_index sorted by (target, source) -- which is
    what radius_graph yields, so SparseTensor's sort is the identity and `value` == edge id."""
    j, i = edge_index
    dist = (pos[i] - pos[j]).pow(2).sum(dim=-1).sqrt()                    # :25
    ptr, cnt = _csr(i, num_nodes)                                         # rows of adj_t
    e_of_t, kj = _expand(ptr[j], cnt[j])                                  # adj_t[j]       :29
    idx_i, idx_j, idx_k = i[e_of_t], j[e_of_t], j[kj]
    keep = idx_i != idx_k                                                 # :36
    idx_i, idx_j, idx_k = idx_i[keep], idx_j[keep], idx_k[keep]
    idx_kj, idx_ji = kj[keep], e_of_t[keep]                               # :40-41
    pos_ji = pos[idx_i] - pos[idx_j]
    pos_jk = pos[idx_k] - pos[idx_j]
    a = (pos_ji * pos_jk).sum(dim=-1)
    b = torch.linalg.cross(pos_ji, pos_jk, dim=-1).norm(dim=-1)
    angle = torch.atan2(b, a)                                             # :44-48
    if not use_torsion:
        return dist, angle, i, j, idx_kj, idx_ji
    # every triplet (k, j, i) against all in-neighbours k_n of j, k_n != i        :53-62
    t_of_q, kn_edge = _expand(ptr[idx_j], cnt[idx_j])
    k_n = j[kn_edge]
    keep = idx_i[t_of_q] != k_n
    t_of_q, k_n = t_of_q[keep], k_n[keep]
    pos_j0 = pos[idx_k[t_of_q]] - pos[idx_j[t_of_q]]
    pos_ji = pos[idx_i[t_of_q]] - pos[idx_j[t_of_q]]
    pos_jk = pos[k_n] - pos[idx_j[t_of_q]]
    dist_ji = pos_ji.pow(2).sum(dim=-1).sqrt()
    plane1 = torch.linalg.cross(pos_ji, pos_j0, dim=-1)
    plane2 = torch.linalg.cross(pos_ji, pos_jk, dim=-1)
    a = (plane1 * plane2).sum(dim=-1)
    b = (torch.linalg.cross(plane1, plane2, dim=-1) * pos_ji).sum(dim=-1) / dist_ji
    torsion1 = torch.atan2(b, a)                                          # :73
    torsion1[torsion1 <= 0] += 2 * math.pi                                # :74
    torsion = shim.scatter(torsion1, t_of_q, dim=0, dim_size=idx_i.numel(), reduce="min")
    return dist, angle, torsion, i, j, idx_kj, idx_ji


# ----------------------------------------------------------------------------- basis
def envelope(x, exponent):                       # spherenet/features.py:151-164
    p = exponent + 1
    a, b, c = -(p + 1) * (p + 2) / 2, p * (p + 2), -p * (p + 1) / 2
    x0 = x.pow(p - 1)
    x1 = x0 * x
    x2 = x1 * x
    return 1. / x + a * x0 + b * x1 + c * x2


def dist_emb(dist, freq, cutoff, exponent):      # spherenet/features.py:180-182
    d = dist.unsqueeze(-1) / cutoff
    return envelope(d, exponent) * (freq * d).sin()


class Basis:
    """angle_emb / torsion_emb closures (spherenet/features.py:185-263, dimenetpp/features.py:183-220)."""

    def __init__(self, tag, ns, nr):
        f = formulas()[tag]
        self.ns, self.nr = ns, nr
        self.bessel = [_compile(s, ["x"]) for s in f["bessel"]]
        self.y0_const = float(f["yl0"][0])
        self.yl0 = [_compile(s, ["theta"]) for s in f["yl0"][1:]]
        if "ylm" in f:
            self.ylm_const = float(f["ylm"][0])
            self.ylm = [_compile(s, ["theta", "phi"]) for s in f["ylm"][1:]]

    def rbf(self, dist, cutoff):
        d = dist / cutoff
        return torch.stack([fn(d) for fn in self.bessel], dim=1)

    def angle_emb(self, dist, angle, idx_kj, cutoff, envelope_exponent=None):
        rbf = self.rbf(dist, cutoff)
        if envelope_exponent is not None:                                 # dimenetpp/features.py:214
            rbf = envelope(dist / cutoff, envelope_exponent).unsqueeze(-1) * rbf
        cbf = torch.stack([torch.zeros_like(angle) + self.y0_const] + [fn(angle) for fn in self.yl0], dim=1)
        n, k = self.ns, self.nr
        return (rbf[idx_kj].view(-1, n, k) * cbf.view(-1, n, 1)).view(-1, n * k)

    def torsion_emb(self, dist, angle, phi, idx_kj, cutoff):
        rbf = self.rbf(dist, cutoff)
        cbf = torch.stack([torch.zeros_like(angle) + torch.zeros_like(phi) + self.ylm_const]
                          + [fn(angle, phi) for fn in self.ylm], dim=1)
        n, k = self.ns, self.nr
        return (rbf[idx_kj].view(-1, 1, n, k) * cbf.view(-1, n, n, 1)).view(-1, n * n * k)


_BASIS_CACHE = {}


def basis(tag, ns, nr):
    key = (tag, ns, nr)
    if key not in _BASIS_CACHE:
        _BASIS_CACHE[key] = Basis(tag, ns, nr)
    return _BASIS_CACHE[key]


# ----------------------------------------------------------------------------- models
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _residual(sd, name, x):                      # spherenet.py:49-50
    return x + swish(_lin(sd, name + ".lin2", swish(_lin(sd, name + ".lin1", x))))


def _update_v(sd, name, e2, i, num_nodes, n_lins):                        # spherenet.py:209-216
    v = shim.scatter(e2, i, dim=0, dim_size=num_nodes)
    v = _lin(sd, name + ".lin_up", v)
    for l in range(n_lins):
        v = swish(_lin(sd, f"{name}.lins.{l}", v))
    return _lin(sd, name + ".lin", v)


def dimenet_family_forward(sd, z, pos, batch, *, torsion, cutoff=5.0, num_layers=4, num_spherical=7,
                           num_radial=6, envelope_exponent=5, num_before_skip=1, num_after_skip=2,
                           num_output_layers=3, num_graphs=None, return_intermediates=False):
    """SphereNet.forward (spherenet.py:296-320, torsion=True) / DimeNetPP.forward
    (dimenetpp.py:273-293, torsion=False) as one function over a state_dict."""
    n = z.size(0)
    if num_graphs is None:
        num_graphs = int(batch.max()) + 1
    edge_index = radius_graph(pos, cutoff, batch)
    tag = "dimenetpp_7_6" if (not torsion and (num_spherical, num_radial) == (7, 6)) \
        else f"spherenet_{num_spherical}_{num_radial}"
    bs = basis(tag, num_spherical, num_radial)
    if torsion:
        dist, angle, tors, i, j, idx_kj, idx_ji = xyz_to_dat(pos, edge_index, n, use_torsion=True)
    else:
        dist, angle, i, j, idx_kj, idx_ji = xyz_to_dat(pos, edge_index, n, use_torsion=False)
    rbf0 = dist_emb(dist, sd["emb.dist_emb.freq"], cutoff, envelope_exponent)
    if torsion:
        sbf = bs.angle_emb(dist, angle, idx_kj, cutoff)
        tbf = bs.torsion_emb(dist, angle, tors, idx_kj, cutoff)
    else:
        sbf = bs.angle_emb(dist, angle, idx_kj, cutoff, envelope_exponent)
        tbf = None
    # init_e                                                               spherenet.py:79-91
    x = F.embedding(z, sd["init_e.emb.weight"])
    r0 = swish(_lin(sd, "init_e.lin_rbf_0", rbf0))
    e1 = swish(_lin(sd, "init_e.lin", torch.cat([x[i], x[j], r0], dim=-1)))
    e2 = _lin(sd, "init_e.lin_rbf_1", rbf0) * e1
    v = _update_v(sd, "init_v", e2, i, n, num_output_layers)
    u = shim.scatter(v, batch, dim=0, dim_size=num_graphs)                 # spherenet.py:313 (0 + scatter)
    inter = dict(edge_index=edge_index, dist=dist, angle=angle, torsion=tors if torsion else None,
                 idx_kj=idx_kj, idx_ji=idx_ji, rbf0=rbf0, sbf=sbf, tbf=tbf, e1_init=e1, v_init=v)
    for l in range(num_layers):                                            # spherenet.py:150-182
        p = f"update_es.{l}"
        x_ji = swish(_lin(sd, p + ".lin_ji", e1))
        x_kj = swish(_lin(sd, p + ".lin_kj", e1))
        x_kj = x_kj * _lin(sd, p + ".lin_rbf2", _lin(sd, p + ".lin_rbf1", rbf0))
        x_kj = swish(_lin(sd, p + ".lin_down", x_kj))
        x_kj = x_kj[idx_kj] * _lin(sd, p + ".lin_sbf2", _lin(sd, p + ".lin_sbf1", sbf))
        if torsion:
            x_kj = x_kj * _lin(sd, p + ".lin_t2", _lin(sd, p + ".lin_t1", tbf))
        x_kj = shim.scatter(x_kj, idx_ji, dim=0, dim_size=e1.size(0))
        x_kj = swish(_lin(sd, p + ".lin_up", x_kj))
        h = x_ji + x_kj
        for r in range(num_before_skip):
            h = _residual(sd, f"{p}.layers_before_skip.{r}", h)
        h = swish(_lin(sd, p + ".lin", h)) + e1
        for r in range(num_after_skip):
            h = _residual(sd, f"{p}.layers_after_skip.{r}", h)
        e1 = h
        e2 = _lin(sd, p + ".lin_rbf", rbf0) * e1
        v = _update_v(sd, f"update_vs.{l}", e2, i, n, num_output_layers)
        u = u + shim.scatter(v, batch, dim=0, dim_size=num_graphs)         # spherenet.py:224
    if return_intermediates:
        inter["e1_final"] = e1
        return u, inter
    return u


def spherenet_forward(sd, z, pos, batch, **kw):
    return dimenet_family_forward(sd, z, pos, batch, torsion=True, **kw)


def dimenetpp_forward(sd, z, pos, batch, **kw):
    return dimenet_family_forward(sd, z, pos, batch, torsion=False, **kw)


def schnet_forward(sd, z, pos, batch, *, cutoff=10.0, num_layers=6, num_gaussians=50, num_graphs=None):
    """SchNet.forward (schnet.py:151-168)."""
    if num_graphs is None:
        num_graphs = int(batch.max()) + 1
    edge_index = radius_graph(pos, cutoff, batch)
    row, col = edge_index
    dist = (pos[row] - pos[col]).norm(dim=-1)                              # :158
    offset = sd["dist_emb.offset"]
    coeff = -0.5 / (offset[1] - offset[0]).item() ** 2                     # :89
    demb = torch.exp(coeff * torch.pow(dist.view(-1, 1) - offset.view(1, -1), 2))   # :92-94
    shift = torch.log(torch.tensor(2.0)).item()
    ssp = lambda t: F.softplus(t) - shift                                  # :97-103
    v = F.embedding(z, sd["init_v.weight"])
    for l in range(num_layers):
        p = f"update_es.{l}"
        C = 0.5 * (torch.cos(dist * math.pi / cutoff) + 1.0)              # :31
        W = _lin(sd, p + ".mlp.2", ssp(_lin(sd, p + ".mlp.0", demb))) * C.view(-1, 1)
        e = _lin(sd, p + ".lin", v)[row] * W                               # :33-34
        q = f"update_vs.{l}"
        out = shim.scatter(e, col, dim=0, dim_size=v.size(0))              # :55
        v = v + _lin(sd, q + ".lin2", ssp(_lin(sd, q + ".lin1", out)))     # :56-59
    h = _lin(sd, "update_u.lin2", ssp(_lin(sd, "update_u.lin1", v)))       # :78-80
    return shim.scatter(h, batch, dim=0, dim_size=num_graphs)
