"""Layers the reference exports from `lib.layers` that are NOT on the ImpFlow hot path but must stay
importable / constructible for the train scripts to load:

  ZeroMeanTransform, Normalize          lib/layers/elemwise.py:8-56   (train_img.py:230 builds Normalize)
  MovingBatchNorm1d / 2d                lib/layers/normalization.py:8-99 (train_toy.py:222,249)
  CouplingBlock & co, InvertibleLinear / InvertibleConv2d   lib/layers/coupling.py, glow.py

The first two groups are small elementwise flows and are implemented here as plain tensor expressions.
Coupling / Glow layers belong to the RealNVP / Glow baselines of the reference (`--arch realnvp`), not to
ImpFlow: the classes exist so that `layers.CouplingBlock` resolves, and constructing one says where it lives."""
import torch
import torch.nn as nn
from torch.nn import Parameter

__all__ = ['ZeroMeanTransform', 'Normalize', 'MovingBatchNorm1d', 'MovingBatchNorm2d', 'CouplingBlock',
           'ChannelCouplingBlock', 'MaskedCouplingBlock', 'InvertibleLinear', 'InvertibleConv2d']


class ZeroMeanTransform(nn.Module):
    """x -> x - 0.5 (volume preserving)."""

    def forward(self, x, logpx=None, restore=False):
        y = x - .5
        return y if logpx is None else (y, logpx)

    def inverse(self, y, logpy=None):
        x = y + .5
        return x if logpy is None else (x, logpy)


class Normalize(nn.Module):
    """Per-channel (x - mean) / std on the first len(mean) channels; logdet = -sum log|std| per pixel."""

    def __init__(self, mean, std):
        nn.Module.__init__(self)
        self.register_buffer('mean', torch.as_tensor(mean, dtype=torch.float32))
        self.register_buffer('std', torch.as_tensor(std, dtype=torch.float32))

    def _affine(self, t, inverse):
        c = len(self.mean)
        m, s = self.mean.view(1, -1, 1, 1), self.std.view(1, -1, 1, 1)
        head = t[:, :c] * s + m if inverse else (t[:, :c] - m) / s
        return head if c == t.shape[1] else torch.cat([head, t[:, c:]], 1)

    def forward(self, x, logpx=None, restore=False):
        y = self._affine(x, False)
        return y if logpx is None else (y, logpx - self._logdetgrad(x))

    def inverse(self, y, logpy=None):
        x = self._affine(y, True)
        return x if logpy is None else (x, logpy + self._logdetgrad(x))

    def _logdetgrad(self, x):
        per_image = -self.std.abs().log().sum() * (x.shape[2] * x.shape[3])
        return per_image.expand(x.shape[0], 1)


class _MovingBatchNorm(nn.Module):
    """Mean-only batch norm with a running mean (volume preserving: the log-density is untouched)."""

    def __init__(self, num_features, eps=1e-4, decay=0.1, bn_lag=0., affine=True):
        super(_MovingBatchNorm, self).__init__()
        self.num_features, self.affine, self.eps, self.decay, self.bn_lag = num_features, affine, eps, decay, bn_lag
        self.register_buffer('step', torch.zeros(1))
        if affine:
            self.bias = Parameter(torch.zeros(num_features))
        else:
            self.register_parameter('bias', None)
        self.register_buffer('running_mean', torch.zeros(num_features))

    def _bc(self, v, like):
        return v.view(1, -1, *([1] * (like.dim() - 2)))

    def forward(self, x, logpx=None, restore=False):
        mean = self.running_mean.clone().detach()
        if self.training:
            batch_mean = x.transpose(0, 1).reshape(x.size(1), -1).mean(1)
            if self.bn_lag > 0:
                mean = batch_mean - (1 - self.bn_lag) * (batch_mean - mean)
                mean = mean / (1. - self.bn_lag ** (self.step[0] + 1))
            with torch.no_grad():
                self.running_mean -= self.decay * (self.running_mean - batch_mean.detach())
                self.step += 1
        y = x - self._bc(mean, x)
        if self.affine:
            y = y + self._bc(self.bias, x)
        return y if logpx is None else (y, logpx)

    def inverse(self, y, logpy=None):
        x = y
        if self.affine:
            x = x - self._bc(self.bias, y)
        x = x + self._bc(self.running_mean, y)
        return x if logpy is None else (x, logpy)

    def __repr__(self):
        return '{}({}, eps={}, decay={}, bn_lag={}, affine={})'.format(type(self).__name__, self.num_features,
                                                                       self.eps, self.decay, self.bn_lag, self.affine)


class MovingBatchNorm1d(_MovingBatchNorm):
    pass


class MovingBatchNorm2d(_MovingBatchNorm):
    pass


class _BaselineOnly(nn.Module):
    def __init__(self, *args, **kwargs):
        super(_BaselineOnly, self).__init__()
        raise NotImplementedError(
            'impflow_b200: %s is part of the RealNVP / Glow baselines of the reference (lib/layers/coupling.py, '
            'glow.py), not of the ImpFlow hot path this package replaces; build it from the reference tree'
            % type(self).__name__)


class CouplingBlock(_BaselineOnly):
    pass


class ChannelCouplingBlock(_BaselineOnly):
    pass


class MaskedCouplingBlock(_BaselineOnly):
    pass


class InvertibleLinear(_BaselineOnly):
    pass


class InvertibleConv2d(_BaselineOnly):
    pass
