"""CPU check of the force-training plumbing (dig_b200/autograd_jvp.py): `energy_with_force` makes
grad(E, pos, create_graph=True) differentiable in the parameters by differentiating the DIRECTIONAL derivative of E
along c = dL/d(dE/dpos).  Here the two callables are plain torch functions (fp64), so torch's own double backward is the
reference: the identity  d/dtheta [c . dE/dpos] = d/dtheta JVP_pos(E; c)  and the autograd wiring are what is tested;
the CUDA tangent kernels are tested on the GPU (tests/test_gpu_train.py)."""
import pytest
import torch

from dig_b200 import autograd_jvp as jv


def _toy():
    torch.manual_seed(0)
    w1 = torch.randn(8, 3, requires_grad=True, dtype=torch.float64)
    w2 = torch.randn(1, 8, requires_grad=True, dtype=torch.float64)
    batch = torch.tensor([0, 0, 0, 1, 1])

    def energy(pos):
        h = torch.tanh(pos @ w1.t()) @ w2.t()
        return torch.zeros(2, 1, dtype=pos.dtype).index_add(0, batch, h)

    def dual(pos, c):
        return torch.func.jvp(energy, (pos,), (c,))
    return (w1, w2), energy, dual


@pytest.mark.parametrize("force_loss", ["l1", "mse"])
def test_energy_with_force_gives_the_double_backward_gradients(force_loss):
    params, energy, dual = _toy()
    pos = torch.randn(5, 3, dtype=torch.float64)
    y = torch.randn(2, 1, dtype=torch.float64)
    f_t = torch.randn(5, 3, dtype=torch.float64)
    lf = torch.nn.functional.l1_loss if force_loss == "l1" else torch.nn.functional.mse_loss

    def total(e, p):
        f = -torch.autograd.grad(e, p, torch.ones_like(e), create_graph=True, retain_graph=True)[0]
        return torch.nn.functional.l1_loss(e, y) + 100 * lf(f, f_t), f

    p1 = pos.clone().requires_grad_(True)
    l1, f1 = total(energy(p1), p1)
    l1.backward()
    ref = [w.grad.clone() for w in params]
    for w in params:
        w.grad = None
    p2 = pos.clone().requires_grad_(True)
    e2 = jv.energy_with_force(energy, dual, p2, params)
    l2, f2 = total(e2, p2)
    assert f2.requires_grad and torch.equal(f1.detach(), f2.detach()) and l1.item() == l2.item()
    l2.backward()
    for a, w in zip(ref, params):
        assert torch.allclose(a, w.grad, rtol=1e-12, atol=1e-12)


def test_plain_backward_without_create_graph_still_works():
    params, energy, dual = _toy()
    pos = torch.randn(5, 3, dtype=torch.float64, requires_grad=True)
    e = jv.energy_with_force(energy, dual, pos, params)
    e.sum().backward()                     # first order only: parameter and position gradients of sum(E)
    p2 = pos.detach().clone().requires_grad_(True)
    ref = torch.autograd.grad(energy(p2).sum(), (p2,) + params)
    assert torch.allclose(pos.grad, ref[0]) and all(torch.allclose(w.grad, r) for w, r in zip(params, ref[1:]))
