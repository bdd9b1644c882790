"""Host-side batch plumbing for the 3D-graph path.

* `Batch`            -- the attribute bag the model `forward(batch_data)` reads
                        (`.z .pos .batch [.y .force .num_graphs .ptr]`), standing in for
                        torch_geometric.data.Batch (reference run.py:121-124).
* `collate` / `DataLoader` -- minimal replacement for torch_geometric.data.DataLoader
                        (reference run.py:6,53-55): concatenates z/pos/force, stacks y, builds
                        the sorted `batch` vector and `ptr`.
* `synthetic_molecules` -- seeded synthetic QM9 / MD17 / OC20-shaped molecules;
                        there is no network, so benchmarks and tests use these shapes.
"""
import math

import torch


class Batch:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device, non_blocking=False):
        out = Batch()
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                v = v.to(device, non_blocking=non_blocking)
            setattr(out, k, v)
        return out

    def pin_memory(self):
        out = Batch()
        for k, v in self.__dict__.items():
            if isinstance(v, torch.Tensor):
                v = v.pin_memory()
            setattr(out, k, v)
        return out

    def keys(self):
        return [k for k, v in self.__dict__.items() if isinstance(v, torch.Tensor)]

    def __repr__(self):
        parts = [f"{k}={tuple(v.shape)}" if isinstance(v, torch.Tensor) else f"{k}={v}"
                 for k, v in self.__dict__.items()]
        return "Batch(" + ", ".join(parts) + ")"


def collate(items, pin_memory=False, sort_by_size=False):
    """List of per-molecule objects (attributes z, pos, optional y/force/...) -> Batch.

    pin_memory:   the concatenated tensors are written straight into page-locked host buffers (one copy, no
                  pageable intermediate), so `batch.to(device, non_blocking=True)` is a true asynchronous DMA and
                  the H2D copy of step n+1 overlaps the kernels of step n.
    sort_by_size: molecules are ordered by atom count (largest first, stable) inside the batch, which packs the
                  128-edge tiles of the interaction kernels with edges of similar molecules; every per-molecule
                  tensor (y, force, ...) moves with its molecule and `perm` records the original positions
                  (`energies[perm.argsort()]` restores the input order)."""
    out = Batch()
    if sort_by_size:
        order = sorted(range(len(items)), key=lambda i: -int(items[i].z.size(0)))
        items = [items[i] for i in order]
        out.perm = torch.tensor(order, dtype=torch.long)
    first = items[0]
    keys = [k for k, v in vars(first).items() if isinstance(v, torch.Tensor)]
    sizes = [int(it.z.size(0)) for it in items]

    def cat(vals):
        if not pin_memory:
            return torch.cat(vals, dim=0)
        rows = sum(v.size(0) for v in vals)
        buf = torch.empty((rows,) + tuple(vals[0].shape[1:]), dtype=vals[0].dtype, pin_memory=True)
        return torch.cat(vals, dim=0, out=buf)
    for k in keys:
        vals = [getattr(it, k) for it in items]
        if k == "y":
            out.y = cat([v.reshape(-1) for v in vals])
        else:
            setattr(out, k, cat(vals))
    sz = torch.tensor(sizes, dtype=torch.long)
    batch = torch.repeat_interleave(torch.arange(len(items)), sz)
    out.batch = batch.pin_memory() if pin_memory else batch
    ptr = torch.zeros(len(items) + 1, dtype=torch.long)
    ptr[1:] = torch.cumsum(sz, 0)
    out.ptr = ptr
    out.num_graphs = len(items)
    return out


class _Collate:
    def __init__(self, pin_memory, sort_by_size):
        self.pin_memory, self.sort_by_size = pin_memory, sort_by_size

    def __call__(self, items):
        return collate(items, self.pin_memory, self.sort_by_size)


class DataLoader(torch.utils.data.DataLoader):
    """`DataLoader(dataset, batch_size, shuffle)` as built at reference run.py:53-55 (PyG's loader: concatenating
    collate).  Two extra keyword switches feed the kernels without host stalls: `pin_memory` (default: on when CUDA is
    present; the collate itself writes into page-locked buffers) and `sort_by_size` (off by default -- it changes
    the order of the molecules inside a batch, which the reference's loader never does)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, pin_memory=None, sort_by_size=False, **kw):
        kw.pop("collate_fn", None)
        if pin_memory is None:
            pin_memory = torch.cuda.is_available()
        super().__init__(dataset, batch_size, shuffle, collate_fn=_Collate(bool(pin_memory), bool(sort_by_size)), **kw)


class Molecule:
    def __init__(self, z, pos, y=None, force=None):
        self.z, self.pos = z, pos
        if y is not None:
            self.y = y
        if force is not None:
            self.force = force


def _sample_points(n, box, min_dist, gen):
    """Rejection-sample n points uniform in `box` with pairwise distance >= min_dist."""
    box_t = torch.tensor(box, dtype=torch.float32)
    pts = torch.empty(0, 3)
    tries = 0
    while pts.size(0) < n:
        cand = torch.rand(4 * n, 3, generator=gen) * box_t
        for c in cand:
            if pts.size(0) == 0 or float((pts - c).norm(dim=1).min()) >= min_dist:
                pts = torch.cat([pts, c[None]], 0)
                if pts.size(0) == n:
                    break
        tries += 1
        if tries > 200:  # box too small for n points: relax by growing it
            box_t = box_t * 1.1
            tries = 0
    return pts


SHAPES = {
    # name: (atoms, box in Angstrom, cutoff)   SURVEY.md §8d
    "schnet-plumbing": (12, (4.0, 4.0, 4.0), 10.0),
    "qm9": (18, (6.0, 5.0, 4.0), 5.0),
    "md17-aspirin": (21, (7.0, 5.0, 3.5), 5.0),
    "oc20-is2re": (73, (12.0, 12.0, 7.0), 6.0),
}


def synthetic_molecules(nmol, shape="qm9", seed=0, natoms=None, variable=False, min_dist=0.95):
    """Seeded list of `Molecule`s of one of the SHAPES.  `variable=True` draws the atom count
    per molecule around the nominal one (QM9: 9..29) and scales the box with it."""
    n0, box, _ = SHAPES[shape]
    if natoms is not None:
        n0 = natoms
    gen = torch.Generator().manual_seed(seed)
    mols = []
    for _ in range(nmol):
        n = n0
        b = box
        if variable:
            n = int(torch.clamp(torch.round(n0 + 0.25 * n0 * torch.randn(1, generator=gen)),
                                max(2, n0 // 2), int(1.6 * n0) + 1).item())
            s = (n / n0) ** (1.0 / 3.0)
            b = tuple(x * s for x in box)
        pos = _sample_points(n, b, min_dist, gen)
        if shape == "md17-aspirin" and n == 21:
            z = torch.tensor([6] * 9 + [8] * 4 + [1] * 8, dtype=torch.long)
        else:
            z = torch.randint(1, 10, (n,), generator=gen)
        y = torch.randn(1, generator=gen)
        force = torch.randn(n, 3, generator=gen)
        mols.append(Molecule(z, pos, y, force))
    return mols


def synthetic_pbc_batch(nsys, natoms=40, cell=(9.0, 9.0, 14.0), cutoff=6.0, max_neighbors=50, seed=0, shuffle_edges=True):
    """Seeded OC20-IS2RE-like periodic batch for the ComENet-OCP variant (reference comenet-ocp.py:335-365 reads
    `atomic_numbers, pos, batch, tags, cell, edge_index, cell_offsets, neighbors`): `nsys` slabs of `natoms` atoms in a
    sheared cell that is periodic in x and y; the edge list holds, per target atom, its `max_neighbors` nearest periodic
    images within `cutoff` (the role of OCP's preprocessed graphs), in a deliberately shuffled order."""
    gen = torch.Generator().manual_seed(seed)
    out = {k: [] for k in ("atomic_numbers", "pos", "tags", "cell", "edge_index", "cell_offsets", "neighbors")}
    sizes, base = [], 0
    for _ in range(nsys):
        n = natoms
        a, b, c = cell
        box = torch.tensor([[a, 0.0, 0.0], [0.3 * b, b, 0.0], [0.0, 0.0, c]]) * (0.9 + 0.2 * torch.rand(1, generator=gen))
        frac = torch.rand(n, 3, generator=gen) * torch.tensor([1.0, 1.0, 0.6])
        pos = frac @ box
        shifts = torch.tensor([[i, j, 0] for i in (-1, 0, 1) for j in (-1, 0, 1)], dtype=torch.float32)
        # image (s, source) seen from target: pos[src] + shifts . box - pos[tgt]
        d = (pos[None, :, None, :] + (shifts @ box)[None, None, :, :] - pos[:, None, None, :]).norm(dim=-1)   # [tgt, src, 9]
        d = d.reshape(n, n * 9)
        ok = (d < cutoff) & (d > 1e-6)
        dd = torch.where(ok, d, torch.full_like(d, float("inf")))
        order = torch.argsort(dd, dim=1)[:, :max_neighbors]
        valid = torch.gather(dd, 1, order) < float("inf")
        tgt = torch.arange(n)[:, None].expand_as(order)[valid]
        flat = order[valid]
        src, sh = flat // 9, flat % 9
        if shuffle_edges:
            p = torch.randperm(tgt.numel(), generator=gen)
            tgt, src, sh = tgt[p], src[p], sh[p]
        out["edge_index"].append(torch.stack([src, tgt]) + base)
        out["cell_offsets"].append(shifts[sh])
        out["neighbors"].append(torch.tensor([tgt.numel()]))
        out["atomic_numbers"].append(torch.randint(1, 84, (n,), generator=gen))
        out["tags"].append((frac[:, 2] * 5).long().clamp(max=2))
        out["pos"].append(pos)
        out["cell"].append(box[None])
        sizes.append(n)
        base += n
    b = Batch(atomic_numbers=torch.cat(out["atomic_numbers"]), pos=torch.cat(out["pos"]), tags=torch.cat(out["tags"]),
              cell=torch.cat(out["cell"]), edge_index=torch.cat(out["edge_index"], 1),
              cell_offsets=torch.cat(out["cell_offsets"]), neighbors=torch.cat(out["neighbors"]))
    b.batch = torch.repeat_interleave(torch.arange(nsys), torch.tensor(sizes))
    b.natoms = torch.tensor(sizes)
    b.num_graphs = nsys
    return b


def synthetic_batch(nmol, shape="qm9", seed=0, **kw):
    return collate(synthetic_molecules(nmol, shape, seed, **kw))


def synthetic_proteins(nprot, length=60, seed=0, variable=True):
    """Seeded protein-shaped batch for ProNet (reference pronet.py:365-372 reads `.x .coords_ca .coords_n .coords_c
    .bb_embs .side_chain_embs .batch`): a C-alpha random walk with 3.8 A steps, N / C atoms ~1.45 / 1.52 A from it,
    amino-acid types in [0, 26), backbone / side-chain torsion embeddings as sin / cos pairs."""
    gen = torch.Generator().manual_seed(seed)
    parts = {k: [] for k in ("x", "coords_ca", "coords_n", "coords_c", "bb_embs", "side_chain_embs")}
    sizes = []
    for _ in range(nprot):
        n = length
        if variable:
            n = int(torch.randint(max(4, length // 2), length + length // 2 + 1, (1,), generator=gen))
        step = torch.randn(n, 3, generator=gen)
        step = 3.8 * step / step.norm(dim=1, keepdim=True)
        ca = torch.cumsum(step, 0)
        # keep the walk compact (fold it back towards the origin) so that the 10 A radius graph is not a chain
        ca = ca * (6.0 * n ** (1.0 / 3.0) / ca.norm(dim=1).max().clamp_min(1.0))

        def offset(r):
            v = torch.randn(n, 3, generator=gen)
            return r * v / v.norm(dim=1, keepdim=True)
        parts["coords_ca"].append(ca)
        parts["coords_n"].append(ca + offset(1.45))
        parts["coords_c"].append(ca + offset(1.52))
        parts["x"].append(torch.randint(0, 26, (n, 1), generator=gen))
        ang = (torch.rand(n, 3, generator=gen) * 2 - 1) * math.pi
        parts["bb_embs"].append(torch.cat([torch.sin(ang), torch.cos(ang)], 1))
        ang = (torch.rand(n, 4, generator=gen) * 2 - 1) * math.pi
        parts["side_chain_embs"].append(torch.cat([torch.sin(ang), torch.cos(ang)], 1))
        sizes.append(n)
    out = Batch(**{k: torch.cat(v, 0) for k, v in parts.items()})
    sz = torch.tensor(sizes, dtype=torch.long)
    out.batch = torch.repeat_interleave(torch.arange(nprot), sz)
    out.num_graphs = nprot
    out.y = torch.randn(nprot, generator=gen)
    return out
