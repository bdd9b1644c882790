"""CPU tests of the dataset readers against the pins of the reference's own tests
(test/threedgraph/dataset/test_QM93D.py:11-34, test_MD17.py:8-18) on small npz files written with the reference's keys."""
import os

import numpy as np
import pytest
import torch

from dig_b200.data import DataLoader
from dig_b200.threedgraph.dataset import MD17, QM93D
from dig_b200.threedgraph.dataset.datasets import QM9_TARGETS


def _write_qm9(root, sizes, seed=0):
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "qm9", "raw"))
    n = np.asarray(sizes)
    data = {"N": n, "R": rng.normal(size=(int(n.sum()), 3)), "Z": rng.integers(1, 10, size=int(n.sum()))}
    for t in QM9_TARGETS:
        data[t] = rng.normal(size=len(sizes))
    np.savez(os.path.join(root, "qm9", "raw", "qm9_eV.npz"), **data)
    return data


def test_qm93d_reader_matches_reference_contract(tmp_path):
    raw = _write_qm9(str(tmp_path), [5, 4, 7, 3, 9, 6])
    ds = QM93D(root=str(tmp_path))
    ds.data.y = ds.data['U0']                                  # the reference's usage (PygQM93D.py:44-45)
    assert len(ds) == 6 and repr(ds) == 'QM93D(6)'
    m = ds[0]
    assert len(vars(m)) == 15                                  # test_QM93D.py:16: pos, z, y + 12 targets
    assert m.y.size() == (1,) and m.z.size() == (5,) and m.pos.size() == (5, 3) and m.Cv.size() == (1,)
    assert m.z.dtype == torch.int64 and m.pos.dtype == torch.float32
    assert float(m.y) == pytest.approx(float(raw['U0'][0]), rel=1e-6)
    assert torch.equal(ds[2].pos, torch.tensor(raw['R'][9:16], dtype=torch.float32))
    split = ds.get_idx_split(len(ds.data.y), train_size=3, valid_size=2, seed=42)
    sub = ds[split['train']]
    assert len(sub) == 3 and torch.equal(sub[1].z, ds[int(split['train'][1])].z)
    assert len(sub[torch.tensor([2, 0])]) == 2 and torch.equal(sub[torch.tensor([2, 0])][0].z, sub[2].z)
    batch = next(iter(DataLoader(sub, batch_size=3, shuffle=False)))
    want_atoms = sum(int(raw['N'][int(i)]) for i in split['train'])
    assert batch.z.shape == (want_atoms,) and batch.pos.shape == (want_atoms, 3) and batch.y.shape == (3,)
    assert batch.batch.shape == (want_atoms,) and batch.num_graphs == 3 and batch.mu.shape == (3,)
    with pytest.raises(FileNotFoundError, match="no network"):
        QM93D(root=str(tmp_path / "nowhere"))


def test_split_indices_equal_the_reference_test_pins():
    """test_QM93D.py:30-33 and test_MD17.py:15-18 pin the first index of each split for seed 42 on the real dataset sizes;
    the split only depends on (size, seed), so it can be checked without the data."""
    ds = object.__new__(QM93D)
    s = ds.get_idx_split(130831, train_size=1000, valid_size=10000, seed=42)
    assert int(s['train'][0]) == 112526 and int(s['valid'][0]) == 120798 and int(s['test'][0]) == 107901
    assert len(s['train']) == 1000 and len(s['valid']) == 10000 and len(s['test']) == 130831 - 11000


def test_md17_reader(tmp_path):
    rng = np.random.default_rng(1)
    os.makedirs(os.path.join(tmp_path, "aspirin", "raw"))
    frames, atoms = 12, 21
    raw = {"E": rng.normal(size=(frames, 1)), "F": rng.normal(size=(frames, atoms, 3)),
           "R": rng.normal(size=(frames, atoms, 3)), "z": rng.integers(1, 9, size=atoms)}
    np.savez(os.path.join(tmp_path, "aspirin", "raw", "aspirin_dft.npz"), **raw)
    ds = MD17(root=str(tmp_path), name='aspirin')
    assert len(ds) == frames and repr(ds) == 'MD17(12)'
    m = ds[3]
    assert len(vars(m)) == 4 and m.z.size() == (21,) and m.pos.size() == (21, 3) and m.force.size() == (21, 3)
    assert m.y.size() == (1,) and torch.equal(m.force, torch.tensor(raw['F'][3], dtype=torch.float32))
    batch = next(iter(DataLoader(ds[torch.arange(4)], batch_size=4)))
    assert batch.force.shape == (84, 3) and batch.y.shape == (4,) and batch.num_graphs == 4


def test_collate_and_synthetic_generators_are_deterministic():
    """Host-side batch plumbing (dig_b200/data.py): seeded generators reproduce, collate builds the sorted batch vector /
    ptr, per-graph targets stack, Batch.to / pin-free round trip keeps python attributes."""
    from dig_b200.data import Batch, collate, synthetic_batch, synthetic_molecules, synthetic_proteins
    a, b = synthetic_batch(5, "qm9", seed=3, variable=True), synthetic_batch(5, "qm9", seed=3, variable=True)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.z, b.z) and torch.equal(a.batch, b.batch)
    assert not torch.equal(a.pos, synthetic_batch(5, "qm9", seed=4, variable=True).pos)
    mols = synthetic_molecules(4, "md17-aspirin", seed=1)
    c = collate(mols)
    assert c.num_graphs == 4 and c.ptr.tolist() == [0, 21, 42, 63, 84] and c.y.shape == (4,) and c.force.shape == (84, 3)
    assert torch.equal(c.batch, torch.arange(4).repeat_interleave(21))
    assert c.z[:21].tolist() == [6] * 9 + [8] * 4 + [1] * 8               # aspirin pattern
    d = (c.pos[:21].unsqueeze(0) - c.pos[:21].unsqueeze(1)).norm(dim=-1) + torch.eye(21) * 10
    assert float(d.min()) >= 0.95 - 1e-6                                  # rejection-sampled minimum distance
    moved = c.to("cpu")
    assert isinstance(moved, Batch) and moved.num_graphs == 4 and set(moved.keys()) >= {"z", "pos", "batch", "y", "force"}
    p, q = synthetic_proteins(3, length=20, seed=2), synthetic_proteins(3, length=20, seed=2)
    assert torch.equal(p.coords_ca, q.coords_ca) and p.x.shape[1] == 1 and p.bb_embs.shape[1] == 6
    assert p.side_chain_embs.shape[1] == 8 and int(p.x.max()) < 26 and p.batch.numel() == p.x.size(0)
    assert torch.allclose((p.coords_n - p.coords_ca).norm(dim=1), torch.full((p.x.size(0),), 1.45), atol=1e-4)


def test_size_sorted_collate_keeps_every_tensor_with_its_molecule():
    """collate(sort_by_size=True): molecules ordered by atom count (largest first, stable); y / pos / z move with their
    molecule, `perm` maps back, and the unsorted collate is unchanged (reference order)."""
    from dig_b200.data import DataLoader, collate, synthetic_molecules
    mols = synthetic_molecules(7, "qm9", seed=3, variable=True)
    for i, m in enumerate(mols):
        m.y = torch.tensor([float(i)])
    plain, srt = collate(mols), collate(mols, sort_by_size=True)
    sizes = [int(m.z.numel()) for m in mols]
    assert plain.y.tolist() == [float(i) for i in range(7)] and not hasattr(plain, "perm")
    order = srt.perm.tolist()
    assert sorted(order) == list(range(7)) and [sizes[i] for i in order] == sorted(sizes, reverse=True)
    assert srt.y.tolist() == [float(i) for i in order]
    assert torch.equal(srt.batch, torch.repeat_interleave(torch.arange(7), torch.tensor([sizes[i] for i in order])))
    lo = 0
    for slot, i in enumerate(order):
        assert torch.equal(srt.pos[lo:lo + sizes[i]], mols[i].pos) and torch.equal(srt.z[lo:lo + sizes[i]], mols[i].z)
        lo += sizes[i]
    assert torch.equal(srt.y[srt.perm.argsort()], plain.y)
    b = next(iter(DataLoader(mols, batch_size=4, shuffle=False, sort_by_size=True, pin_memory=False)))
    assert b.num_graphs == 4 and sorted(b.perm.tolist()) == [0, 1, 2, 3]
