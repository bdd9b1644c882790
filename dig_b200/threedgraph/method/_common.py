"""Pieces shared by the drop-in model classes (parameter holders + initialisers)."""
import math

import torch
from torch import nn


def swish(x):
    """Marker for the default activation (reference spherenet.py:14); the fused kernels implement
    x * sigmoid(x) on the device, this Python function is only compared by identity."""
    return x * torch.sigmoid(x)


def glorot_orthogonal(tensor, scale):
    """torch_geometric.nn.inits.glorot_orthogonal as used at reference spherenet.py:44-47:
    orthogonal init rescaled so that Var(W) = scale / (fan_in + fan_out)."""
    if tensor is not None:
        torch.nn.init.orthogonal_(tensor.data)
        scale /= ((tensor.size(-2) + tensor.size(-1)) * tensor.var())
        tensor.data *= scale.sqrt()


def glorot(tensor):
    """torch_geometric.nn.inits.glorot (reference comenet.py:52-53)."""
    if tensor is not None:
        stdv = math.sqrt(6.0 / (tensor.size(-2) + tensor.size(-1)))
        tensor.data.uniform_(-stdv, stdv)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: dig_b200 models run on CUDA (sm_100a) tensors only -- got {t.device}. "
            "There is no CPU fallback; move the model and the batch to a B200 device.")


def wants_grad(module):
    """True when autograd is recording and the module has trainable parameters: forward() then takes the
    differentiable training path (dig_b200/autograd.py) instead of the fused inference kernels."""
    return torch.is_grad_enabled() and (bool(getattr(module, "energy_and_force", False)) or
                                        any(p.requires_grad for p in module.parameters()))


class ResidualLayer(nn.Module):
    """Parameter holder for reference ResidualLayer (spherenet.py:34-50); evaluated inside the
    fused update_e kernel."""

    def __init__(self, hidden_channels):
        super().__init__()
        self.lin1 = nn.Linear(hidden_channels, hidden_channels)
        self.lin2 = nn.Linear(hidden_channels, hidden_channels)
        self.reset_parameters()

    def reset_parameters(self):
        glorot_orthogonal(self.lin1.weight, scale=2.0)
        self.lin1.bias.data.fill_(0)
        glorot_orthogonal(self.lin2.weight, scale=2.0)
        self.lin2.bias.data.fill_(0)
