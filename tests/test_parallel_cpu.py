"""world_size-2 gloo tests (CPU) of the graph-sharded data-parallel plumbing (dig_b200/parallel.py)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dig_b200 import parallel
from dig_b200.data import collate, synthetic_molecules


def test_shard_bounds_partition():
    for n in (0, 1, 7, 128, 1025):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mols = synthetic_molecules(9, "qm9", seed=5, variable=True)
        mine = parallel.shard_molecules(mols)
        b = collate(mine)
        # stand-in "energies": something every rank can compute on CPU per molecule
        e_local = torch.stack([m.pos.sum() + m.z.sum() for m in mine]).unsqueeze(1)
        e_all = parallel.gather_per_molecule(e_local)
        want = torch.stack([m.pos.sum() + m.z.sum() for m in mols]).unsqueeze(1)
        ok = torch.equal(e_all, want) and b.num_graphs == len(mine)
        step_ms = parallel.max_over_ranks(10.0 + rank)              # slowest rank defines the step
        total = parallel.sum_over_ranks(len(mine))
        ok = ok and step_ms == 10.0 + world - 1 and total == len(mols)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_equal_shards_give_equal_step_counts():
    """961 molecules on 2 ranks at batch_size 32 must give the same number of steps on both ranks."""
    from dig_b200 import parallel
    from dig_b200.data import DataLoader, synthetic_molecules
    mols = synthetic_molecules(13, "qm9", seed=0)
    for world in (2, 3, 4):
        shards = [parallel.shard_molecules_equal(mols, r, world) for r in range(world)]
        assert len({len(s) for s in shards}) == 1 and sum(len(s) for s in shards) == 13 - 13 % world
        assert shards[0][0] is mols[0] and shards[-1][-1] is mols[world * (13 // world) - 1]
        assert len({len(DataLoader(s, 2, shuffle=True)) for s in shards}) == 1


def test_two_rank_gloo_sharding_and_reductions():
    world = 2
    ret = mp.Manager().dict()
    port = 29000 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world


def _grad_worker(rank, world, port, ret):
    """Each rank backpropagates its shard; after allreduce_gradients every rank holds the full-batch gradient.
    (A plain torch layer stands in for the model: the CUDA kernels cannot run here, the exchange logic can.)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 1))
        frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
        unused = torch.nn.Parameter(torch.ones(2))                  # never reaches the loss: grad None -> zeros
        x = torch.randn(8, 5)
        y = torch.randn(8, 1)
        lo, hi = parallel.shard_bounds(8, rank, world)
        loss = torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi])        # equal shards: mean of means = mean
        loss.backward()
        nbytes = parallel.allreduce_gradients(list(model.parameters()) + [frozen, unused])
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 1))
        ref.load_state_dict(model.state_dict())
        torch.nn.functional.mse_loss(ref(x), y).backward()
        ok = all(torch.allclose(p.grad, q.grad, atol=1e-6) for p, q in zip(model.parameters(), ref.parameters()))
        ok = ok and frozen.grad is None and torch.equal(unused.grad, torch.zeros(2))
        ok = ok and nbytes == 4 * (sum(p.numel() for p in model.parameters()) + 2)
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce():
    world = 2
    ret = mp.Manager().dict()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_grad_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world


def _reducer_worker(rank, world, port, ret):
    """Overlapped bucketed all-reduce over the flat gradient buffer: hooks fire during backward, a parameter that gets
    no gradient this step is reduced by finish(); result = the full-batch gradient on every rank, for 2 steps."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.a, self.b = torch.nn.Linear(5, 7), torch.nn.Linear(7, 1)
                self.unused = torch.nn.Parameter(torch.ones(3))

            def forward(self, x):
                return self.b(torch.tanh(self.a(x)))
        model, ref = Net(), Net()
        ref.load_state_dict(model.state_dict())
        flat = parallel.FlatParameters(model)
        reducer = parallel.GradientReducer(flat, n_buckets=2)
        ok = len(reducer.buckets) == 2 and reducer.buckets[-1][1] == flat.numel
        ok = ok and all(p.data_ptr() == flat.data.data_ptr() + 4 * o for p, o in zip(flat.params, flat.offsets))
        for step in range(2):
            x, y = torch.randn(8, 5), torch.randn(8, 1)
            lo, hi = parallel.shard_bounds(8, rank, world)
            flat.zero_grad()
            reducer.arm()
            torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
            reducer.finish()
            ref.zero_grad()
            torch.nn.functional.mse_loss(ref(x), y).backward()
            for (n, p), q in zip(model.named_parameters(), ref.parameters()):
                want = q.grad if q.grad is not None else torch.zeros_like(q)
                ok = ok and torch.allclose(p.grad, want, atol=1e-6)
                ok = ok and p.grad.data_ptr() >= flat.grad.data_ptr()          # still a view of the flat buffer
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_overlapped_flat_gradient_reducer():
    world = 2
    ret = mp.Manager().dict()
    port = 28000 + (os.getpid() % 400)
    mp.spawn(_reducer_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)) and len(ret) == world


def test_run_driver_contract_on_cpu(tmp_path, capsys):
    """run().run(...) (reference run.py:20-101) with a plain torch model on CPU: the host-side driver logic (loaders,
    epoch loop, printed lines, checkpoint keys, best-validation bookkeeping) does not need the CUDA kernels."""
    from dig_b200.threedgraph.evaluation import ThreeDEvaluator
    from dig_b200.threedgraph.method import run

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(10, 4)
            self.lin = torch.nn.Linear(4, 1)

        def forward(self, b):
            h = self.lin(self.emb(b.z))
            return torch.zeros(b.num_graphs, 1).index_add_(0, b.batch, h)

    mols = synthetic_molecules(12, "qm9", seed=1, natoms=5)
    torch.manual_seed(0)
    run().run(torch.device("cpu"), mols[:8], mols[8:10], mols[10:], Tiny(), loss_func=torch.nn.L1Loss(),
              evaluation=ThreeDEvaluator(), epochs=3, batch_size=4, vt_batch_size=2, lr=1e-2, lr_decay_step_size=1,
              save_dir=str(tmp_path / "ck"), log_dir='')
    out = capsys.readouterr().out
    assert "#Params: 45" in out and out.count("=====Epoch") == 3 and "Training..." in out and "Evaluating..." in out
    assert "Testing..." in out and "'Train':" in out and "Best validation MAE so far:" in out
    assert "Test MAE when got best validation result:" in out
    ck = torch.load(str(tmp_path / "ck" / "valid_checkpoint.pt"), weights_only=False)
    assert set(ck) == {'epoch', 'model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict', 'best_valid_mae',
                       'num_params'} and ck['num_params'] == 45


def test_paced_gc_disables_and_restores_collection():
    """parallel.PacedGC: automatic cycle collection is off inside the context, tick() collects every `every` steps (the
    same steps on every rank), and the previous state comes back on exit."""
    import gc
    from dig_b200 import parallel
    assert gc.isenabled()
    with parallel.PacedGC(every=3) as pace:
        assert not gc.isenabled()

        class Node:
            pass
        a, b = Node(), Node()
        a.other, b.other = b, a                 # a reference cycle only the collector can free
        import weakref
        probe = weakref.ref(a)
        del a, b
        pace.tick(); pace.tick()
        assert probe() is not None              # nothing collected yet
        pace.tick()                             # third step: paced collection
        assert probe() is None
    assert gc.isenabled()


def test_inference_plan_follows_the_parameters(monkeypatch):
    """Host logic of the cached inference plan with the packing ops stubbed out (no GPU): the plan is
    reused while nothing changed and rebuilt after an in-place update (tensor._version), a `.data` write followed by
    invalidate_packed(), load_state_dict, and a REPLACED parameter object (parameter-registration hook); copies of the
    model do not carry it."""
    import copy
    import torch
    from torch import nn
    from dig_b200 import ops
    from dig_b200.threedgraph.method import SphereNet, ComENet

    built = []
    monkeypatch.setattr(ops, "update_v_h16_supported", lambda *a, **k: True)
    monkeypatch.setattr(ops, "pack_update_v_h16", lambda holders, cache: ("parr", "varr", len(holders[0].lins)))
    monkeypatch.setattr(ops, "pack_init_e", lambda m: "init_w")
    monkeypatch.setattr(ops, "tc_pack_matrix", lambda w, cache, key, kind="tc": torch.zeros(1))
    monkeypatch.setattr(ops, "init_e_tables", lambda m, cache: ("ti", "tj", "packed"))
    monkeypatch.setattr(ops, "tc_pack_update_e", lambda m, tors, cache, kind="tc": built.append(m.lin_up.weight) or "layer")
    model = SphereNet()
    p1 = model._inference_plan()
    assert p1 is model._inference_plan() and len(built) == 4
    with torch.no_grad():
        model.update_es[2].lin.weight.add_(1.0)                    # in place: version bump
    p2 = model._inference_plan()
    assert p2 is not p1 and p2 is model._inference_plan()
    model.update_es[1].lin_kj.weight.data.mul_(2.0)                # behind autograd's back: the version does not move
    assert model._inference_plan() is p2
    model.invalidate_packed()
    p3 = model._inference_plan()
    assert p3 is not p2
    model.load_state_dict(model.state_dict())
    p4 = model._inference_plan()
    assert p4 is not p3
    fresh = nn.Parameter(torch.zeros_like(model.update_es[0].lin_up.weight))
    model.update_es[0].lin_up.weight = fresh                       # a NEW parameter object
    built.clear()
    p5 = model._inference_plan()
    assert p5 is not p4 and built[0] is fresh
    assert any(p is fresh for p in model.__dict__["_plan_params"])
    with torch.no_grad():
        fresh.add_(1.0)                                            # ... whose later updates are tracked too
    assert model._inference_plan() is not p5
    clone = copy.deepcopy(model)
    assert "_plan" not in clone.__dict__ and "_plan_params" not in clone.__dict__ and "_plan" in model.__dict__

    # ComENet: same key logic (ops.plan_key); its plan builder needs the packed-weight registry, only the key is checked
    cm = ComENet(cutoff=6.0)
    k1 = ops.plan_key(cm)
    assert k1 == ops.plan_key(cm)
    with torch.no_grad():
        cm.lin_out.weight.mul_(0.5)
    assert ops.plan_key(cm) != k1
