"""Factories of the Lipschitz layers and the closed-form operator-norm layers — mirror of
lib/layers/base/lipschitz.py:274-366 (LopLinear / LopConv2d), :462-507 (operator_norm_settings) and :510-531
(get_linear / get_conv2d).

get_linear / get_conv2d hand every (domain, codomain) pair with a closed-form operator norm — 1 -> {1, 2, inf},
{2, inf} -> inf — to the Lop* layers (e.g. the first and last layer of a branch under vnorms '122f'); everything else
is an InducedNorm* layer (power iteration, mixed_lipschitz.py).  The SpectralNorm* classes are never built by the
factories; they exist only for the isinstance() checks of the train scripts
(train_img.py:567-579, 786-792) and refuse construction."""
import torch
import torch.nn as nn

from ... import _cabi, ops
from .mixed_lipschitz import InducedNormConv2d, InducedNormLinear, _from_nhwc, _to_nhwc

__all__ = ['get_linear', 'get_conv2d', 'SpectralNormLinear', 'SpectralNormConv2d', 'LopLinear', 'LopConv2d',
           'operator_norm_settings']


class _OutOfScope(nn.Module):
    def __init__(self, *a, **k):
        super(_OutOfScope, self).__init__()
        raise NotImplementedError('impflow_b200: %s is outside the hot-path scope (no factory builds it; the 2 -> 2 '
                                  'layers are InducedNorm*)' % type(self).__name__)


class SpectralNormLinear(_OutOfScope):
    pass


class SpectralNormConv2d(_OutOfScope):
    pass


def operator_norm_settings(domain, codomain):
    """(maximum over the input dimension?, vector norm taken over the other one) of the operator norms that have a
    closed form (lipschitz.py:483-507): the largest column 1- / 2- / inf-norm for domain 1, the largest row 2- / 1-norm
    for codomain inf."""
    inf = float('inf')
    table = {(1, 1): (True, 1), (1, 2): (True, 2), (1, inf): (True, inf), (2, inf): (False, 2), (inf, inf): (False, 1)}
    if (domain, codomain) not in table:
        raise ValueError('Unknown combination of domain "{}" and codomain "{}"'.format(domain, codomain))
    return table[(domain, codomain)]


def _norm_except_dim(w, norm_type, dim):
    """The norm over every dimension but `dim`, kept as size-1 axes (lipschitz.py:467-480).  A reduction over the
    layer's own weights (<= 1 MB, once per forward): library reductions, differentiable."""
    other = [a for a in range(w.dim()) if a != dim]
    if norm_type == float('inf'):
        # the reference's _max_except_dim takes the SIGNED maximum (no abs); kept (lipschitz.py:474-480)
        return w.amax(dim=other, keepdim=True)
    if norm_type == 1:
        return w.abs().sum(dim=other, keepdim=True)
    return w.pow(2).sum(dim=other, keepdim=True).sqrt()


class _LopMixin(object):
    """Soft rescale by the closed-form operator norm: per output row (domain 1: max_across_dim = 1 keeps the INPUT
    axis, so every input column is scaled on its own) or per input column, or by the single largest one when
    local_constraint is off (lipschitz.py:298-307, 347-356)."""

    def _lop_init(self, coeff, domain, codomain, local_constraint):
        self.coeff = coeff
        self.domain = domain
        self.codomain = codomain
        self.local_constraint = local_constraint
        max_across_input_dims, self.norm_type = operator_norm_settings(self.domain, self.codomain)
        self.max_across_dim = 1 if max_across_input_dims else 0
        self.register_buffer('scale', torch.tensor(0.))

    def compute_weight(self):
        scale = _norm_except_dim(self.weight, self.norm_type, dim=self.max_across_dim)
        if not self.local_constraint:
            scale = scale.max()
        with torch.no_grad():
            self.scale.copy_(scale.max())
        factor = torch.max(torch.ones(1, device=self.weight.device), scale / self.coeff)      # soft normalisation
        return self.weight / factor


class LopLinear(_LopMixin, nn.Linear):
    """Linear layer whose Lipschitz constant is bounded through a closed-form operator norm (lipschitz.py:274-317);
    the product runs on the impflow GEMM kernels."""

    def __init__(self, in_features, out_features, bias=True, coeff=0.97, domain=float('inf'), codomain=float('inf'),
                 local_constraint=True, **unused_kwargs):
        del unused_kwargs
        nn.Linear.__init__(self, in_features, out_features, bias)
        self._lop_init(coeff, domain, codomain, local_constraint)

    def forward(self, input):
        _cabi.require_device(input, 'LopLinear input')
        weight = self.compute_weight()
        shape = input.shape
        y = ops.linear(input.reshape(-1, shape[-1]), weight, self.bias)
        return y.view(*shape[:-1], self.out_features)

    def extra_repr(self):
        return nn.Linear.extra_repr(self) + ', coeff={}, domain={}, codomain={}, local={}'.format(
            self.coeff, self.domain, self.codomain, self.local_constraint)


class LopConv2d(_LopMixin, nn.Conv2d):
    """Convolution bounded the same way (lipschitz.py:320-366).  The kernels cover what the flows build: 1x1 and
    3x3, stride 1, 'same' padding (implicit_flow.py:359-398)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=True, coeff=0.97,
                 domain=float('inf'), codomain=float('inf'), local_constraint=True, **unused_kwargs):
        del unused_kwargs
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=bias)
        self._lop_init(coeff, domain, codomain, local_constraint)
        k = self.kernel_size
        if k not in ((1, 1), (3, 3)) or self.stride != (1, 1) or self.padding != (k[0] // 2, k[1] // 2):
            raise NotImplementedError('impflow_b200: LopConv2d covers 1x1 and 3x3 kernels with stride 1 and same '
                                      'padding (got kernel {}, stride {}, padding {})'.format(k, self.stride,
                                                                                              self.padding))

    def forward(self, input):
        _cabi.require_device(input, 'LopConv2d input')
        weight = self.compute_weight()
        x = _to_nhwc(input)
        if self.kernel_size == (1, 1):
            y = ops.conv1x1_nhwc(x, weight, self.bias)
        else:
            y = ops.conv3x3_nhwc(x, weight, self.bias)
        return _from_nhwc(y)

    def extra_repr(self):
        return nn.Conv2d.extra_repr(self) + ', coeff={}, domain={}, codomain={}, local={}'.format(
            self.coeff, self.domain, self.codomain, self.local_constraint)


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
