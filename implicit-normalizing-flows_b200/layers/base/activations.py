"""Lipschitz activations of the residual branches, evaluated by the act_mul CUDA kernel.

API mirror of lib/layers/base/activations.py (Sin :7-12, Swish :64-71, Identity :15-18,
Zero :20-23).  Only the activations the hot-path configs use run custom kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops

__all__ = ['Sin', 'Swish', 'Identity', 'Zero', 'ReLU']


class Sin(nn.Module):
    """sin(2 pi x) / (2 pi)  (activations.py:11-12)."""
    act_kind = ops.ACT_SIN

    def forward(self, x):
        return ops.activation(x, ops.ACT_SIN)


class Swish(nn.Module):
    """LipSwish: x * sigmoid(x * softplus(beta)) / 1.1, learnable beta (activations.py:64-71)."""
    act_kind = ops.ACT_LIPSWISH

    def __init__(self):
        super(Swish, self).__init__()
        self.beta = nn.Parameter(torch.tensor([0.5]))

    def beta_sp(self):
        return F.softplus(self.beta)

    def forward(self, x):
        return ops.activation(x, ops.ACT_LIPSWISH, self.beta_sp())


class ReLU(nn.Module):
    """Kernel-backed ReLU; torch.nn.ReLU modules inside a branch are accepted as well."""
    act_kind = ops.ACT_RELU

    def __init__(self, inplace=False):
        super(ReLU, self).__init__()

    def forward(self, x):
        return ops.activation(x, ops.ACT_RELU)


class Identity(nn.Module):

    def forward(self, x):
        return x


class Zero(nn.Module):

    def forward(self, x):
        return torch.zeros_like(x)
