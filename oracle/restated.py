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
def xyztodat_knn(pos, edge_index, num_nodes, batch):
    """G-SphereNet's private geometry (ggraph3D/method/G_SphereNet/model/geometric_computing.py:54-104): same distances,
    triplets and angles as xyz_to_dat; the torsion of a triplet (k -> j -> i) uses ONE reference atom, the nearest
    neighbour of j in its graph (the second nearest when the nearest is i), and is mapped to (0, 2 pi]."""
    dist, angle, i, j, idx_kj, idx_ji = xyz_to_dat(pos, edge_index, num_nodes, use_torsion=False)
    n = pos.size(0)
    nbr = shim.knn_graph(pos, 2, batch)                                     # grouped by query, nearest first  :14-19
    cnt = torch.bincount(nbr[1], minlength=n)
    assert int(cnt.min()) == 2, "every graph needs at least three atoms (the reference errors otherwise)"
    near = nbr[0].view(n, 2)
    idx_i, idx_j, idx_k = i[idx_ji], j[idx_ji], j[idx_kj]
    k_n = near[idx_j, 0].clone()                                            # :87-93
    mask = k_n == idx_i
    k_n[mask] = near[idx_j, 1][mask]
    pos_j0, pos_ji, pos_jk = pos[idx_k] - pos[idx_j], pos[idx_i] - pos[idx_j], pos[k_n] - pos[idx_j]   # :96-98
    dist_ji = pos_ji.pow(2).sum(dim=-1).sqrt()
    plane1 = torch.linalg.cross(pos_ji, pos_j0, dim=-1)
    plane2 = torch.linalg.cross(pos_ji, pos_jk, dim=-1)
    a = (plane1 * plane2).sum(dim=-1)
    b = (torch.linalg.cross(plane1, plane2, dim=-1) * pos_ji).sum(dim=-1) / dist_ji
    torsion = torch.atan2(b, a)
    torsion = torch.where(torsion <= 0, torsion + 2 * math.pi, torsion)    # :102-103
    return dist, angle, torsion, i, j, idx_kj, idx_ji


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
                           num_output_layers=3, num_graphs=None, return_intermediates=False, node_feature=None):
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
    if "init_e.emb.weight" in sd:
        x = F.embedding(z, sd["init_e.emb.weight"])
    else:                                                                  # use_node_features=False   spherenet.py:83-84
        x = sd["init_e.node_embedding"][None, :].expand(z.shape[0], -1)
    if node_feature is not None and "extra_emb.weight" in sd:              # use_extra_node_feature    spherenet.py:85-86,298-299
        x = torch.cat((x, _lin(sd, "extra_emb", node_feature)), 1)
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


# ----------------------------------------------------------------------------- ComENet
def comenet_geometry(pos, edge_index, num_nodes, cutoff, vecs=None):
    """comenet.py:295-385: nearest / second-nearest reference atoms and the angles theta, phi, tau."""
    j, i = edge_index
    if vecs is None:
        vecs = pos[j] - pos[i]                                             # :297
    dist = vecs.norm(dim=-1)
    n_e = i.numel()

    def two_nearest(index):                                                # :304-327
        _, a0 = shim.scatter_min(dist, index, dim_size=num_nodes)
        a0[a0 >= n_e] = 0
        add = torch.zeros_like(dist)
        add[a0] = cutoff
        _, a1 = shim.scatter_min(dist + add, index, dim_size=num_nodes)
        a1[a1 >= n_e] = 0
        return a0, a1

    argmin0, argmin1 = two_nearest(i)
    argmin0_j, argmin1_j = two_nearest(j)
    n0, n1 = j[argmin0][i], j[argmin1][i]
    n0_j, n1_j = i[argmin0_j][j], i[argmin1_j][j]
    mask_iref = n0 == j                                                    # :344-354
    idx_iref = argmin0[i].clone()
    idx_iref[mask_iref] = argmin1[i][mask_iref]
    mask_jref = n0_j == i
    idx_jref = argmin0_j[j].clone()
    idx_jref[mask_jref] = argmin1_j[j][mask_jref]
    pos_ji, pos_in0, pos_in1 = vecs, vecs[argmin0][i], vecs[argmin1][i]
    pos_iref, pos_jref_j = vecs[idx_iref], vecs[idx_jref]
    cross = lambda a, b: torch.linalg.cross(a, b, dim=-1)

    def fold(t):                                                           # x[x<0] += pi   :368,377,385
        t = t.clone()
        t[t < 0] = t[t < 0] + math.pi
        return t

    a = ((-pos_ji) * pos_in0).sum(dim=-1)
    b = cross(-pos_ji, pos_in0).norm(dim=-1)
    theta = fold(torch.atan2(b, a))
    dist_ji = pos_ji.pow(2).sum(dim=-1).sqrt()
    plane1, plane2 = cross(-pos_ji, pos_in0), cross(-pos_ji, pos_in1)
    a = (plane1 * plane2).sum(dim=-1)
    b = (cross(plane1, plane2) * pos_ji).sum(dim=-1) / dist_ji
    phi = fold(torch.atan2(b, a))
    plane1, plane2 = cross(pos_ji, pos_jref_j), cross(pos_ji, pos_iref)
    a = (plane1 * plane2).sum(dim=-1)
    b = (cross(plane1, plane2) * pos_ji).sum(dim=-1) / dist_ji
    tau = fold(torch.atan2(b, a))
    return dist, theta, phi, tau


def comenet_features(dist, theta, phi, tau, cutoff, ns=2, nr=3):
    """comenet/features.py:284-295 (angle_emb) and :340-348 (torsion_emb)."""
    bs = basis(f"comenet_{ns}_{nr}", ns, nr)
    rbf = bs.rbf(dist, cutoff)                                             # [E, ns*nr]
    y0 = torch.stack([torch.zeros_like(tau) + bs.y0_const] + [fn(tau) for fn in bs.yl0], dim=1)
    feature2 = (rbf.view(-1, ns, nr) * y0.view(-1, ns, 1)).view(-1, ns * nr)
    ylm = torch.stack([torch.zeros_like(theta) + bs.ylm_const] + [fn(theta, phi) for fn in bs.ylm], dim=1)
    degree = torch.arange(ns, device=dist.device) * 2 + 1
    r = rbf.view(-1, ns, nr).repeat_interleave(degree, dim=1).view(-1, ns ** 2 * nr)
    feature1 = r * ylm.repeat_interleave(nr, dim=1)
    return feature1, feature2


def comenet_forward(sd, z, pos, batch, *, cutoff=8.0, num_layers=4, num_radial=3, num_spherical=2,
                    num_output_layers=3, num_graphs=None, return_intermediates=False):
    """ComENet._forward (comenet.py:288-399)."""
    n = z.size(0)
    if num_graphs is None:
        num_graphs = int(batch.max()) + 1
    edge_index = radius_graph(pos, cutoff, batch)
    j, i = edge_index
    dist, theta, phi, tau = comenet_geometry(pos, edge_index, n, cutoff)
    f1, f2 = comenet_features(dist, theta, phi, tau, cutoff, num_spherical, num_radial)
    x = swish(F.embedding(z.long(), sd["emb.emb.weight"]))                  # comenet.py:125-127
    for b in range(num_layers):                                            # comenet.py:195-215
        p = f"interaction_blocks.{b}"
        x = swish(_lin(sd, p + ".lin", x))
        hs = []
        for c, feat in ((1, f1), (2, f2)):
            w = _lin(sd, f"{p}.lin_feature{c}.lin2", _lin(sd, f"{p}.lin_feature{c}.lin1", feat))
            agg = torch.zeros_like(x).index_add_(0, i, w * x[j])           # EdgeGraphConv.message + add aggr
            h = _lin(sd, f"{p}.conv{c}.lin_rel", agg) + _lin(sd, f"{p}.conv{c}.lin_root", x)
            hs.append(swish(_lin(sd, f"{p}.lin{c}", h)))
        h = _lin(sd, p + ".lin_cat", torch.cat(hs, 1)) + x
        for l in range(num_output_layers):
            h = swish(_lin(sd, f"{p}.lins.{l}", h)) + h
        # GraphNorm (torch_geometric 2.1.0)
        mean = shim.scatter(h, batch, dim=0, dim_size=num_graphs, reduce="mean")
        out = h - mean.index_select(0, batch) * sd[p + ".norm.mean_scale"]
        var = shim.scatter(out.pow(2), batch, dim=0, dim_size=num_graphs, reduce="mean")
        std = (var + 1e-5).sqrt().index_select(0, batch)
        h = sd[p + ".norm.weight"] * out / std + sd[p + ".norm.bias"]
        x = _lin(sd, p + ".final", h)
    for l in range(num_output_layers):
        x = swish(_lin(sd, f"lins.{l}", x))
    x = _lin(sd, "lin_out", x)
    energy = shim.scatter(x, batch, dim=0, dim_size=num_graphs)
    if return_intermediates:
        return energy, dict(edge_index=edge_index, dist=dist, theta=theta, phi=phi, tau=tau,
                            feature1=f1, feature2=f2)
    return energy


def comenet_ocp_forward(sd, data, *, cutoff=6.0, num_blocks=4, num_radial=3, num_spherical=2, num_output_layers=3):
    """ComENet._forward of the OCP variant (comenet/ocp/comenet-ocp.py:335-470) with use_pbc=True, otf_graph=False,
    hetero=False: get_pbc_distances (oracle/ocp_stub.py restates the third-party function), the same reference-atom /
    angle code as comenet.py on the periodic distance vectors, blocks with middle = hidden width."""
    from .ocp_stub import get_pbc_distances
    z, batch = data.atomic_numbers.long(), data.batch
    n = z.size(0)
    num_graphs = int(batch.max()) + 1
    out = get_pbc_distances(data.pos, data.edge_index, data.cell, data.cell_offsets, data.neighbors,
                            return_distance_vec=True)                     # :352-365
    edge_index, vecs = out["edge_index"], out["distance_vec"]
    j, i = edge_index
    dist, theta, phi, tau = comenet_geometry(None, edge_index, n, cutoff, vecs=vecs)
    f1, f2 = comenet_features(dist, theta, phi, tau, cutoff, num_spherical, num_radial)
    x = swish(F.embedding(z, sd["emb.emb.weight"]))
    for b in range(num_blocks):                                            # :241-266
        p = f"interaction_blocks.{b}"
        x = swish(_lin(sd, p + ".lin", x))
        hs = []
        for c, feat in ((1, f1), (2, f2)):
            w = _lin(sd, f"{p}.lin_feature{c}.lin2", _lin(sd, f"{p}.lin_feature{c}.lin1", feat))
            agg = torch.zeros_like(x).index_add_(0, i, w * x[j])
            h = _lin(sd, f"{p}.conv{c}.lin_rel", agg) + _lin(sd, f"{p}.conv{c}.lin_root", x)
            hs.append(swish(_lin(sd, f"{p}.lin{c}", h)))
        h = _lin(sd, p + ".lin_cat", torch.cat(hs, 1)) + x
        for l in range(num_output_layers):
            h = swish(_lin(sd, f"{p}.lins.{l}", h)) + h
        mean = shim.scatter(h, batch, dim=0, dim_size=num_graphs, reduce="mean")
        o = h - mean.index_select(0, batch) * sd[p + ".norm.mean_scale"]
        var = shim.scatter(o.pow(2), batch, dim=0, dim_size=num_graphs, reduce="mean")
        std = (var + 1e-5).sqrt().index_select(0, batch)
        h = sd[p + ".norm.weight"] * o / std + sd[p + ".norm.bias"]
        x = _lin(sd, p + ".final", h)
    for l in range(num_output_layers):
        x = swish(_lin(sd, f"lins.{l}", x))
    x = _lin(sd, "lin_out", x)
    return shim.scatter(x, batch, dim=0, dim_size=num_graphs)               # :469


# ----------------------------------------------------------------------------- ProNet
def pronet_geometry(pos, pos_n, pos_c, edge_index, level):
    """dist, theta, phi and tau (aminoacid) or the three Euler angles (backbone / allatom); pronet.py:383-449."""
    j, i = edge_index
    n = pos.size(0)
    dist = (pos[i] - pos[j]).norm(dim=1)
    refi0, refi1 = (i - 1) % n, (i + 1) % n
    a = ((pos[j] - pos[i]) * (pos[refi0] - pos[i])).sum(dim=-1)
    b = torch.cross(pos[j] - pos[i], pos[refi0] - pos[i], dim=-1).norm(dim=-1)
    theta = torch.atan2(b, a)
    plane1 = torch.cross(pos[refi0] - pos[i], pos[refi1] - pos[i], dim=-1)
    plane2 = torch.cross(pos[refi0] - pos[i], pos[j] - pos[i], dim=-1)
    a = (plane1 * plane2).sum(dim=-1)
    b = (torch.cross(plane1, plane2, dim=-1) * (pos[refi0] - pos[i])).sum(dim=-1) / ((pos[refi0] - pos[i]).norm(dim=-1))
    phi = torch.atan2(b, a)
    if level == "aminoacid":
        refi = (i - 1) % n
        refj0 = (j - 1) % n
        refj = (j - 1) % n
        refj1 = (j + 1) % n
        mask = refi0 == j
        refi[mask] = refi1[mask]
        mask = refj0 == i
        refj[mask] = refj1[mask]
        plane1 = torch.cross(pos[j] - pos[i], pos[refi] - pos[i], dim=-1)
        plane2 = torch.cross(pos[j] - pos[i], pos[refj] - pos[j], dim=-1)
        a = (plane1 * plane2).sum(dim=-1)
        b = (torch.cross(plane1, plane2, dim=-1) * (pos[j] - pos[i])).sum(dim=-1) / dist
        return dist, theta, phi, [torch.atan2(b, a)]
    or1_x = pos_n[i] - pos[i]
    or1_z = torch.cross(or1_x, torch.cross(or1_x, pos_c[i] - pos[i], dim=-1), dim=-1)
    or1_len = or1_z.norm(dim=1) + 1e-7
    or2_x = pos_n[j] - pos[j]
    or2_z = torch.cross(or2_x, torch.cross(or2_x, pos_c[j] - pos[j], dim=-1), dim=-1)
    or2_len = or2_z.norm(dim=1) + 1e-7
    nn_ = torch.cross(or1_z, or2_z, dim=-1)
    angle1 = torch.atan2((torch.cross(or1_x, nn_, dim=-1) * or1_z).sum(dim=-1) / or1_len, (or1_x * nn_).sum(dim=-1))
    angle2 = torch.atan2(torch.cross(or1_z, or2_z, dim=-1).norm(dim=-1), (or1_z * or2_z).sum(dim=-1))
    angle3 = torch.atan2((torch.cross(nn_, or2_x, dim=-1) * or2_z).sum(dim=-1) / or2_len, (nn_ * or2_x).sum(dim=-1))
    return dist, theta, phi, [angle1, angle2, angle3]


def pronet_pos_emb(edge_index, num_pos_emb=16):                              # pronet.py:352-362
    import numpy as np
    d = edge_index[0] - edge_index[1]
    frequency = torch.exp(torch.arange(0, num_pos_emb, 2, dtype=torch.float32, device=edge_index.device)
                          * -(np.log(10000.0) / num_pos_emb))
    angles = d.unsqueeze(-1) * frequency
    return torch.cat((torch.cos(angles), torch.sin(angles)), -1)


def pronet_features(dist, theta, phi, angles, cutoff, ns=2, nr=6):
    """feature0 = d_theta_phi_emb(dist, theta, phi) [E, nr*ns^2]; feature1 = cat of d_angle_emb(dist, a) [E, nr*ns] per
    angle (pronet/features.py:253-344: the ComENet closed forms with num_radial = 6)."""
    bs = basis(f"pronet_{ns}_{nr}", ns, nr)
    rbf = bs.rbf(dist, cutoff)
    ylm = torch.stack([torch.zeros_like(theta) + bs.ylm_const] + [fn(theta, phi) for fn in bs.ylm], dim=1)
    degree = torch.arange(ns, device=dist.device) * 2 + 1
    r = rbf.view(-1, ns, nr).repeat_interleave(degree, dim=1).view(-1, ns ** 2 * nr)
    f0 = r * ylm.repeat_interleave(nr, dim=1)
    f1 = []
    for ang in angles:
        y0 = torch.stack([torch.zeros_like(ang) + bs.y0_const] + [fn(ang) for fn in bs.yl0], dim=1)
        f1.append((rbf.view(-1, ns, nr) * y0.view(-1, ns, 1)).view(-1, ns * nr))
    return f0, torch.cat(f1, 1)


def pronet_forward(sd, data, *, level="aminoacid", cutoff=10.0, num_blocks=4, int_emb_layers=3, out_layers=2,
                   num_pos_emb=16, max_num_neighbors=32, num_radial=6, num_spherical=2, return_intermediates=False):
    """ProNet.forward (pronet.py:364-469), dropout 0, no noise."""
    z = torch.squeeze(data.x.long())
    pos, batch = data.coords_ca, data.batch
    if level == "aminoacid":
        x = F.embedding(z, sd["embedding.weight"])
    else:
        feats = [F.one_hot(z, num_classes=26).float(), data.bb_embs]
        if level == "allatom":
            feats.append(data.side_chain_embs)
        x = _lin(sd, "embedding", torch.cat(feats, dim=1))
    edge_index = radius_graph(pos, cutoff, batch, max_num_neighbors=max_num_neighbors)
    pe = pronet_pos_emb(edge_index, num_pos_emb)
    j, i = edge_index
    dist, theta, phi, angles = pronet_geometry(pos, data.coords_n, data.coords_c, edge_index, level)
    f0, f1 = pronet_features(dist, theta, phi, angles, cutoff, num_spherical, num_radial)
    n = z.size(0)
    for b in range(num_blocks):                                              # InteractionBlock.forward, pronet.py:222-253
        p = f"interaction_blocks.{b}"
        x1 = swish(_lin(sd, p + ".lin_1", x))
        x2 = swish(_lin(sd, p + ".lin_2", x))
        hs = []
        for c, feat in ((0, f0), (1, f1), (2, pe)):
            w = _lin(sd, f"{p}.lin_feature{c}.lin2", _lin(sd, f"{p}.lin_feature{c}.lin1", feat))
            agg = torch.zeros_like(x1).index_add_(0, i, w * x1[j])
            h = _lin(sd, f"{p}.conv{c}.lin_l", agg) + _lin(sd, f"{p}.conv{c}.lin_r", x1)
            hs.append(swish(_lin(sd, f"{p}.lin{c}", h)))
        h = torch.cat(hs, 1)
        for l in range(int_emb_layers):
            h = swish(_lin(sd, f"{p}.lins_cat.{l}", h))
        h = h + x2
        for l in range(int_emb_layers - 1):
            h = swish(_lin(sd, f"{p}.lins.{l}", h))
        x = _lin(sd, p + ".final", h)
    num_graphs = int(batch.max()) + 1
    y = shim.scatter(x, batch, dim=0, dim_size=num_graphs)
    for l in range(out_layers - 1):
        y = torch.relu(_lin(sd, f"lins_out.{l}", y))
    y = _lin(sd, "lin_out", y)
    if return_intermediates:
        return y, dict(edge_index=edge_index, dist=dist, theta=theta, phi=phi, angles=angles, feature0=f0, feature1=f1,
                       pos_emb=pe)
    return y
