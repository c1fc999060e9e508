"""Factories of the Lipschitz layers — mirror of lib/layers/base/lipschitz.py:510-531
(get_linear / get_conv2d).  The Lop*/SpectralNorm* classes are never built by any shipped
config (all vnorms are '2'); they exist only for isinstance() checks made
by the train scripts (train_img.py:567-579, 786-792)."""
import torch.nn as nn

from .mixed_lipschitz import InducedNormConv2d, InducedNormLinear

__all__ = ['get_linear', 'get_conv2d', 'SpectralNormLinear', 'SpectralNormConv2d', 'LopLinear', 'LopConv2d']


class _OutOfScope(nn.Module):
    def __init__(self, *a, **k):
        super(_OutOfScope, self).__init__()
        raise NotImplementedError('impflow_b200: %s is outside the hot-path scope (only induced 2-norm layers '
                                  'are used by the shipped configs)' % type(self).__name__)


class SpectralNormLinear(_OutOfScope):
    pass


class SpectralNormConv2d(_OutOfScope):
    pass


class LopLinear(_OutOfScope):
    pass


class LopConv2d(_OutOfScope):
    pass


def get_linear(in_features, out_features, bias=True, coeff=0.97, domain=None, codomain=None, **kwargs):
    _linear = InducedNormLinear
    if domain == 1:
        if codomain in [1, 2, float('inf')]:
            _linear = LopLinear
    elif codomain == float('inf'):
        if domain in [2, float('inf')]:
            _linear = LopLinear
    return _linear(in_features, out_features, bias, coeff, domain, codomain, **kwargs)


def get_conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=True, coeff=0.97, domain=None,
               codomain=None, **kwargs):
    _conv2d = InducedNormConv2d
    if domain == 1:
        if codomain in [1, 2, float('inf')]:
            _conv2d = LopConv2d
    elif codomain == float('inf'):
        if domain in [2, float('inf')]:
            _conv2d = LopConv2d
    return _conv2d(in_channels, out_channels, kernel_size, stride, padding, bias, coeff, domain, codomain, **kwargs)
