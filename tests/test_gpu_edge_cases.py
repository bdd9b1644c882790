"""GPU edge cases: empty / ragged inputs, isolated atoms, the neighbour cap, error behaviour."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_isolated_atoms_and_single_atom_graphs():
    from dig_b200 import ops
    from oracle import restated
    dev = torch.device("cuda:0")
    # graph 0: two atoms 20 A apart (no edges); graph 1: one atom; graph 2: a bonded triple
    pos = torch.tensor([[0, 0, 0], [20, 0, 0], [5, 5, 5], [0, 0, 0], [1, 0, 0], [0, 1.2, 0]],
                       dtype=torch.float32, device=dev)
    batch = torch.tensor([0, 0, 1, 2, 2, 2], device=dev)
    gr = ops.build_graph(pos, batch, 5.0)
    ops.triplet_geometry(gr, pos, use_torsion=True, want_idx64=True)
    ei = restated.radius_graph(pos, 5.0, batch)
    res = restated.xyz_to_dat(pos, ei, 6, use_torsion=True)
    assert torch.equal(gr.edge_index, ei) and gr.n_edges == 6
    assert torch.equal(gr.idx_kj64, res[-2]) and torch.equal(gr.torsion, res[2])


def test_no_edges_at_all():
    from dig_b200.threedgraph.method import DimeNetPP
    from dig_b200.data import Batch
    dev = torch.device("cuda:0")
    model = DimeNetPP().to(dev)
    b = Batch(z=torch.tensor([1, 6], device=dev), pos=torch.tensor([[0., 0, 0], [50, 0, 0]], device=dev),
              batch=torch.tensor([0, 1], device=dev))
    with torch.no_grad():
        u = model(b)
    assert u.shape == (2, 1) and torch.isfinite(u).all()


def test_neighbour_cap_matches_torch_cluster_semantics():
    from dig_b200 import ops
    from oracle import restated
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    pos = (torch.rand(90, 3) * 3.0).to(dev)       # dense: > 33 candidates per node
    batch = torch.zeros(90, dtype=torch.long, device=dev)
    gr = ops.build_graph(pos, batch, 6.0)
    ei = restated.radius_graph(pos, 6.0, batch)
    assert torch.equal(gr.edge_index, ei)
    deg = torch.bincount(ei[1], minlength=90)
    assert int(deg.max()) == 33 and int(deg.min()) == 32
    ops.triplet_geometry(gr, pos, use_torsion=True, want_idx64=True)
    res = restated.xyz_to_dat(pos, ei, 90, use_torsion=True)
    assert torch.equal(gr.idx_kj64, res[-2]) and torch.equal(gr.idx_ji64, res[-1])
    assert torch.equal(gr.angle, res[1]) and torch.equal(gr.torsion, res[2])


def test_ragged_batch_with_empty_graph_slot():
    from dig_b200 import ops
    dev = torch.device("cuda:0")
    pos = torch.rand(7, 3, device=dev)
    batch = torch.tensor([0, 0, 0, 2, 2, 2, 2], device=dev)     # graph 1 is empty
    gr = ops.build_graph(pos, batch, 5.0, num_graphs=3)
    assert gr.graph_ptr.tolist() == [0, 3, 3, 7]
    assert gr.n_edges == 3 * 2 + 4 * 3


def test_errors_are_loud():
    from dig_b200 import ops
    from dig_b200._lib import Dig3dError
    dev = torch.device("cuda:0")
    with pytest.raises(ValueError):
        ops.build_graph(torch.rand(4, 2, device=dev), None, 5.0)
    with pytest.raises(TypeError):
        ops.build_graph(torch.rand(4, 3, device=dev, dtype=torch.float64), None, 5.0)
    with pytest.raises(Dig3dError):
        ops.build_graph(torch.rand(4, 3, device=dev), torch.zeros(4, dtype=torch.long, device=dev), 5.0,
                        max_num_neighbors=200)
