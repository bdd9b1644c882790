"""dig_b200.pipeline.InferencePipeline: two batches in flight give the plain loop's energies bit for bit, and run.val
(which uses it) returns the plain loop's MAE."""
import pytest
import torch

from helpers import formula_state_dict

pytestmark = pytest.mark.gpu


def test_two_batches_in_flight_equal_the_plain_loop():
    from dig_b200.data import synthetic_batch
    from dig_b200.pipeline import InferencePipeline
    from dig_b200.threedgraph.method import SphereNet
    from dig_b200 import ops
    dev = torch.device("cuda:0")
    model = SphereNet()
    model.load_state_dict(formula_state_dict(model.state_dict(), seed=2))
    model = model.to(dev).eval()
    host = [synthetic_batch(n, "qm9", seed=10 + i, variable=True).pin_memory() for i, n in enumerate((24, 7, 40, 1, 33, 16, 9))]
    with torch.no_grad():
        plain = [model(b.to(dev)).cpu() for b in host]
    pipe = InferencePipeline(model, dev, depth=2)
    for rep in range(3):                       # slots and pinned buffers are reused across passes
        got = [o.clone() for o in pipe.map(host)]
        assert len(got) == len(plain)
        for a, b in zip(got, plain):
            assert a.shape == b.shape and torch.equal(a, b), rep
    torch.cuda.synchronize()
    assert ops.tc_timeouts() == 0
    t0 = pipe.submit(host[0])
    t1 = pipe.submit(host[1])
    with pytest.raises(RuntimeError, match="never taken"):
        pipe.submit(host[2])
    assert torch.equal(pipe.result(t0), plain[0]) and torch.equal(pipe.result(t1), plain[1])
    with pytest.raises(RuntimeError, match="not in flight"):
        pipe.result(t0)


def test_run_val_uses_the_pipeline_and_returns_the_plain_mae():
    from dig_b200.data import DataLoader, synthetic_molecules
    from dig_b200.threedgraph.evaluation import ThreeDEvaluator
    from dig_b200.threedgraph.method import DimeNetPP, run
    dev = torch.device("cuda:0")
    mols = synthetic_molecules(22, "qm9", seed=3, variable=True)
    torch.manual_seed(0)
    model = DimeNetPP(num_layers=2).to(dev)
    mae = run().val(model, DataLoader(mols, 5, shuffle=False), False, 100, ThreeDEvaluator(), dev)
    model.eval()
    with torch.no_grad():
        preds = torch.cat([model(b.to(dev)) for b in DataLoader(mols, 5, shuffle=False)])
    y = torch.cat([m.y.reshape(-1) for m in mols]).unsqueeze(1).to(dev)
    assert abs(mae - (preds - y).abs().mean().item()) < 1e-6 * max(1.0, abs(mae))
