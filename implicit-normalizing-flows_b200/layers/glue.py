"""Elementwise invertible layers chained around imBlock: ActNorm1d/2d (act_norm.py:9-79),
SqueezeLayer (squeeze.py:7-45), LogitTransform (elemwise.py:58-88).

These are the "next" rows of SURVEY.md §8(f), kept API- and state-dict-compatible.  ActNorm's training forward
and backward run as fused kernels through the C ABI (ops.actnorm); its data-dependent init, the inverses, Squeeze
(a pure permutation) and the one-off LogitTransform are plain tensor expressions on the GPU."""
import math

import torch
import torch.nn as nn
from torch.nn import Parameter

from .. import ops
from .extras import Normalize, ZeroMeanTransform  # noqa: F401  (lib/layers/elemwise.py exports them too)

FUSED_ACTNORM = {'on': True}     # A/B switch: fused kernels against the plain tensor expressions

__all__ = ['ActNorm1d', 'ActNorm2d', 'SqueezeLayer', 'LogitTransform']


class ActNormNd(nn.Module):

    def __init__(self, num_features, eps=1e-12):
        super(ActNormNd, self).__init__()
        self.num_features = num_features
        self.eps = eps
        self.weight = Parameter(torch.Tensor(num_features))
        self.bias = Parameter(torch.Tensor(num_features))
        self.register_buffer('initialized', torch.tensor(0))
        self._init_known = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._init_known = None
        return super(ActNormNd, self)._load_from_state_dict(*args, **kwargs)

    @property
    def shape(self):
        raise NotImplementedError

    def _maybe_init(self, x):
        if self._init_known is None or not self._init_known:
            self._init_known = bool(self.initialized)     # one device read until initialised
        if self._init_known:
            return
        with torch.no_grad():      # data-dependent init (act_norm.py:25-37)
            c = x.size(1)
            x_t = x.transpose(0, 1).contiguous().view(c, -1)
            batch_mean = torch.mean(x_t, dim=1)
            batch_var = torch.max(torch.var(x_t, dim=1), torch.tensor(0.2).to(x_t))
            self.bias.data.copy_(-batch_mean)
            self.weight.data.copy_(-0.5 * torch.log(batch_var))
            self.initialized.fill_(1)
            self._init_known = True

    def forward(self, x, logpx=None, restore=None):
        self._maybe_init(x)
        if FUSED_ACTNORM['on'] and x.dtype == torch.float32 and (
                logpx is None or (torch.is_tensor(logpx) and logpx.dtype == torch.float32
                                  and logpx.numel() == x.size(0))):
            # one kernel forward (y and the log-density update), one C call backward (csrc/elementwise.cu)
            return ops.actnorm(x, self.bias, self.weight, logpx)
        y = (x + self.bias.view(*self.shape)) * torch.exp(self.weight.view(*self.shape))
        if logpx is None:
            return y
        return y, logpx - self._logdetgrad(x)

    def inverse(self, y, logpy=None):
        x = y * torch.exp(-self.weight.view(*self.shape)) - self.bias.view(*self.shape)
        if logpy is None:
            return x
        return x, logpy + self._logdetgrad(x)

    def _logdetgrad(self, x):
        per_sample = x.numel() // x.size(0) // self.num_features
        return (self.weight.sum() * per_sample).expand(x.size(0), 1)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.num_features)


class ActNorm1d(ActNormNd):
    @property
    def shape(self):
        return [1, -1]


class ActNorm2d(ActNormNd):
    @property
    def shape(self):
        return [1, -1, 1, 1]


class SqueezeLayer(nn.Module):

    def __init__(self, downscale_factor):
        super(SqueezeLayer, self).__init__()
        self.downscale_factor = downscale_factor

    def forward(self, x, logpx=None, restore=False):
        y = squeeze(x, self.downscale_factor)
        return y if logpx is None else (y, logpx)

    def inverse(self, y, logpy=None):
        x = torch.pixel_shuffle(y, self.downscale_factor)
        return x if logpy is None else (x, logpy)


def squeeze(x, r=2):
    """[:, C, H*r, W*r] -> [:, C*r^2, H, W]  (squeeze.py:32-45)."""
    b, c, h, w = x.shape
    v = x.reshape(b, c, h // r, r, w // r, r).permute(0, 1, 3, 5, 2, 4)
    return v.reshape(b, c * r * r, h // r, w // r)


class LogitTransform(nn.Module):
    """x -> logit(alpha + (1 - 2 alpha) x)  (elemwise.py:58-88)."""

    def __init__(self, alpha=1e-6):
        nn.Module.__init__(self)
        self.alpha = alpha

    def forward(self, x, logpx=None, restore=False):
        s = self.alpha + (1 - 2 * self.alpha) * x
        y = torch.log(s) - torch.log(1 - s)
        if logpx is None:
            return y
        return y, logpx - self._logdetgrad(x).view(x.size(0), -1).sum(1, keepdim=True)

    def inverse(self, y, logpy=None):
        x = (torch.sigmoid(y) - self.alpha) / (1 - 2 * self.alpha)
        if logpy is None:
            return x
        return x, logpy + self._logdetgrad(x).view(x.size(0), -1).sum(1, keepdim=True)

    def _logdetgrad(self, x):
        s = self.alpha + (1 - 2 * self.alpha) * x
        return -torch.log(s - s * s) + math.log(1 - 2 * self.alpha)

    def __repr__(self):
        return '{}({})'.format(self.__class__.__name__, self.alpha)
