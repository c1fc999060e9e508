"""Lipschitz activations of the residual branches, evaluated by the act_mul CUDA kernel.

API mirror of lib/layers/base/activations.py (Sin :7-12, Swish :64-71, Identity :15-18,
Zero :20-23, FullSort :25-28, MaxMin :31-38, LipschitzCube :41-44).  Only the activations the hot-path
configs use (Sin, Swish, ReLU) run custom kernels; the group-sort / cube activations exist because the train
scripts build their ACTIVATION_FNS tables from them at import (train_toy.py:27-30, train_tabular.py:29-32) and
run as plain tensor expressions (a branch that uses one is evaluated through the module / autograd path)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops

__all__ = ['Sin', 'Swish', 'Identity', 'Zero', 'ReLU', 'FullSort', 'MaxMin', 'LipschitzCube']


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


class FullSort(nn.Module):
    """Sort the features of every sample (1-Lipschitz, gradient-norm preserving)."""

    def forward(self, x):
        return torch.sort(x, 1)[0]


class MaxMin(nn.Module):
    """Pairs of neighbouring features -> (all maxima, all minima)."""

    def forward(self, x):
        pairs = x.view(x.shape[0], x.shape[1] // 2, 2)
        return torch.cat([pairs.max(2)[0], pairs.min(2)[0]], 1)


class LipschitzCube(nn.Module):
    """x^3/3 inside (-1, 1), continued with slope 1 outside."""

    def forward(self, x):
        return torch.where(x >= 1, x - 2 / 3, torch.where(x <= -1, x + 2 / 3, x ** 3 / 3))
