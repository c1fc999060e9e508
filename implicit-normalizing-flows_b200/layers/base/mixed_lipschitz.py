"""Induced-norm constrained Linear / Conv2d — API and state-dict mirror of
lib/layers/base/mixed_lipschitz.py (InducedNormLinear :12-146, InducedNormConv2d :149-403).

domain = codomain = 2 (every shipped config) runs on the power-iteration kernels
(csrc/spectral*.cu).  Other induced p -> q norms and learnable orders (`learn_p`) take the
general iteration below: the same W v / W^T u products on the impflow GEMM / conv kernels, the dual-norm
normalisations of mixed_lipschitz.py:414-444 on the (small) vectors between them, the reference's random restarts
at initialisation.  The soft rescale W / max(1, sigma/coeff) with sigma = u^T W v does not depend on the norm, so
everything above the layer (branch programs, weight gradients) is unchanged.  There is no CPU execution path except
the constructor-time power iteration that the reference also performs on the host before the model is moved to the
GPU."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

from ... import _cabi, ops

__all__ = ['InducedNormLinear', 'InducedNormConv2d', 'update_lipschitz', 'sigma_of', 'normalize_u', 'normalize_v',
           'projmax_', 'vector_norm', 'asym_squash', 'leaky_elu']


def _is_two(p):
    """The reference's test `not torch.is_tensor(p) and p == 2` (a learnable order is never "2")."""
    return (not torch.is_tensor(p)) and p == 2


def _plain_l2(domain, codomain):
    return _is_two(domain) and _is_two(codomain)


# ---- dual-norm normalisations of the general power iteration (mixed_lipschitz.py:406-444; the algorithm of
# http://www.qetlab.com/InducedMatrixNorm that the reference cites) --------------------------------------------------

def vector_norm(x, p):
    x = x.reshape(-1)
    return torch.sum(x ** p) ** (1 / p)


def projmax_(v):
    """In place: the unit vector at the entry of largest magnitude."""
    ind = torch.argmax(torch.abs(v))
    v.zero_()
    v[ind] = 1
    return v


def _dual_direction(x, power, norm_p):
    """sign(x) * m / ||m||_{norm_p} with m = (|x| / max|x|)^power, sign(0) = 1."""
    mag = torch.abs(x)
    sgn = x / mag
    sgn[torch.isnan(sgn)] = 1
    mag = (mag / torch.max(mag)) ** power
    return sgn * mag / vector_norm(mag, norm_p)


def normalize_v(v, domain, out=None):
    if _is_two(domain):
        return F.normalize(v, p=2, dim=0, out=out)
    if domain == 1:
        return projmax_(v)
    return _dual_direction(v, 1 / (domain - 1), domain)


def normalize_u(u, codomain, out=None):
    if _is_two(codomain):
        return F.normalize(u, p=2, dim=0, out=out)
    if codomain == float('inf'):
        return projmax_(u)
    if codomain == 1:
        return _dual_direction(u, codomain - 1, float('inf'))
    return _dual_direction(u, codomain - 1, codomain / (codomain - 1))


def leaky_elu(x, a=0.3):
    return a * x + (1 - a) * F.elu(x)


def asym_squash(x):
    """Learnable order -> (1, 5), 2 at x = 0 (mixed_lipschitz.py:451-452)."""
    return torch.tanh(-leaky_elu(-x + 0.5493061829986572)) * 2 + 3


def _converged(u, v, old_u, old_v, atol, rtol):
    """The reference's stopping rule (:115-120); one host read per iteration, as there."""
    err_u = torch.norm(u - old_u) / (u.nelement() ** 0.5)
    err_v = torch.norm(v - old_v) / (v.nelement() ** 0.5)
    return bool((err_u < atol + rtol * torch.max(u)) & (err_v < atol + rtol * torch.max(v)))


def _require_cuda(t, what):
    _cabi.require_device(t, what)


def _state_key(m):
    """Identifies the (weight, u, v) contents a cached sigma / d sigma/dW belongs to."""
    return (m.weight._version, m.u._version, m.v._version, m.weight.data_ptr(), m.u.data_ptr(), m.v.data_ptr())


class _Sigma(torch.autograd.Function):
    """sigma = u^T W v on the device (mixed_lipschitz.py:125, 320); linear in W."""

    @staticmethod
    def forward(ctx, W2d, u, v):
        ctx.save_for_backward(u, v)
        sigma, _ = ops.sn_power_iter(W2d, u, v, 0, 0.0, 0.0)   # 0 iterations: u, v untouched
        return sigma

    @staticmethod
    def backward(ctx, g):
        u, v = ctx.saved_tensors
        return g * torch.outer(u, v), None, None


def _soft_rescale(weight, sigma, coeff):
    # soft normalisation: only when sigma is larger than coeff (mixed_lipschitz.py:128-131)
    factor = torch.max(torch.ones(1, device=weight.device), sigma / coeff)
    return weight / factor


def _mv(W2d, v):
    """W v for a dense (out, in) matrix: the impflow strided GEMM on the device, torch.mv at construction time on
    the host (the reference's constructor runs there too)."""
    if W2d.is_cuda:
        return ops.gemm_strided(v.reshape(1, -1), False, W2d, True).reshape(-1)
    return torch.mv(W2d, v)


def _mtv(W2d, u):
    if W2d.is_cuda:
        return ops.gemm_strided(u.reshape(1, -1), False, W2d, False).reshape(-1)
    return torch.mv(W2d.t(), u)


def _general_power_iteration(apply_w, apply_wt, u, v, domain, codomain, n_iterations, atol, rtol):
    """u <- normalize_u(W v), v <- normalize_v(W^T u) until the reference's tolerance rule or the iteration cap
    (mixed_lipschitz.py:104-120, :293-311, :343-366).  Returns (u, v, iterations used)."""
    max_itrs = 200 if n_iterations is None else n_iterations
    tol_mode = n_iterations is None and atol is not None and rtol is not None
    used = 0
    for _ in range(max_itrs):
        old_u, old_v = u, v
        u = normalize_u(apply_w(v), codomain)
        v = normalize_v(apply_wt(u), domain)
        used += 1
        if tol_mode and _converged(u, v, old_u, old_v, atol, rtol):
            break
    return u, v, used


class InducedNormLinear(nn.Module):

    def __init__(self, in_features, out_features, bias=True, coeff=0.97, domain=2, codomain=2, n_iterations=None,
                 atol=None, rtol=None, zero_init=False, **unused_kwargs):
        del unused_kwargs
        super(InducedNormLinear, self).__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.coeff = coeff
        self.n_iterations = n_iterations
        self.atol = atol
        self.rtol = rtol
        self.domain = domain
        self.codomain = codomain
        self.weight = nn.Parameter(torch.Tensor(out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_features))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters(zero_init)
        with torch.no_grad():
            dom, cod = self.compute_domain_codomain()
        h, w = self.weight.shape
        self.register_buffer('scale', torch.tensor(0.))
        self.register_buffer('u', normalize_u(self.weight.new_empty(h).normal_(0, 1), cod))
        self.register_buffer('v', normalize_v(self.weight.new_empty(w).normal_(0, 1), dom))
        if _plain_l2(dom, cod):
            self._init_power_iteration_host(200)
        else:
            self._init_general(dom, cod, h, w)

    def reset_parameters(self, zero_init=False):
        # same draws as the reference (mixed_lipschitz.py:58-66)
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if zero_init:
            self.weight.data.div_(1000)
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def _init_power_iteration_host(self, n):
        """Constructor-time 200 power iterations (mixed_lipschitz.py:44-46) on the construction
        device; not part of the hot path."""
        with torch.no_grad():
            if self.weight.is_cuda:
                sigma, _ = ops.sn_power_iter(self.weight, self.u, self.v, n, 0.0, 0.0)
                self.scale.copy_(sigma[0])
                return
            W, u, v = self.weight, self.u, self.v
            for _ in range(n):
                u = F.normalize(torch.mv(W, v), dim=0)
                v = F.normalize(torch.mv(W.t(), u), dim=0)
            self.u.copy_(u)
            self.v.copy_(v)
            self.scale.copy_(torch.dot(u, torch.mv(W, v)))

    def _init_general(self, dom, cod, h, w):
        """Non-2 norms: 200 iterations from the first start, then the reference's ten random restarts, which keep the
        start whose sigma beats the FIRST one (its best_scale is never raised, mixed_lipschitz.py:44-56)."""
        with torch.no_grad():
            W = self.weight.detach()
            sigma = self._iterate_general(W, dom, cod, 200, None, None)
            best_scale, best_u, best_v = sigma.clone(), self.u.clone(), self.v.clone()
            for _ in range(10):
                self.u.copy_(normalize_u(W.new_empty(h).normal_(0, 1), cod))
                self.v.copy_(normalize_v(W.new_empty(w).normal_(0, 1), dom))
                sigma = self._iterate_general(W, dom, cod, 200, None, None)
                if sigma > best_scale:
                    best_u, best_v = self.u.clone(), self.v.clone()
            self.u.copy_(best_u)
            self.v.copy_(best_v)

    def _iterate_general(self, W, dom, cod, n_iterations, atol, rtol):
        """General-norm power iteration on the buffers; returns sigma = u^T W v (also left in `scale`)."""
        u, v, _ = _general_power_iteration(lambda x: _mv(W, x), lambda x: _mtv(W, x), self.u.clone(), self.v.clone(),
                                           dom, cod, n_iterations, atol, rtol)
        self.u.copy_(u)
        self.v.copy_(v)
        sigma = torch.dot(self.u, _mv(W, self.v))
        self.scale.copy_(sigma)
        return sigma

    def compute_domain_codomain(self):
        if torch.is_tensor(self.domain):          # learnable orders (learn_p): asym_squash of the raw parameters
            return asym_squash(self.domain), asym_squash(self.codomain)
        return self.domain, self.codomain

    def sigma_gradient(self):
        """D = d sigma / d W = u v^T (sigma = u^T W v is linear in W), cached until u or v change."""
        key = (self.u._version, self.v._version, self.u.data_ptr(), self.v.data_ptr())
        cached = getattr(self, '_sigma_grad', None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                cached = self._sigma_grad = (key, torch.outer(self.u, self.v).contiguous())
        return cached[1]

    def compute_one_iter(self):
        _require_cuda(self.weight, 'InducedNormLinear.weight')
        dom, cod = self.compute_domain_codomain()
        if not _plain_l2(dom, cod):
            with torch.no_grad():
                W = self.weight.detach()
                u = normalize_u(_mv(W, self.v.detach()), cod)
                v = normalize_v(_mtv(W, u), dom)
                return torch.dot(u, _mv(W, v))
        u, v = self.u.clone(), self.v.clone()
        sigma, _ = ops.sn_power_iter(self.weight.detach(), u, v, 1, 0.0, 0.0)
        return sigma[0]

    def compute_weight(self, update=True, n_iterations=None, atol=None, rtol=None):
        _require_cuda(self.weight, 'InducedNormLinear.weight')
        if update:
            n_iterations = self.n_iterations if n_iterations is None else n_iterations
            atol = self.atol if atol is None else atol
            rtol = self.rtol if rtol is None else atol      # reference quirk (mixed_lipschitz.py:94)
            if n_iterations is None and (atol is None or rtol is None):
                raise ValueError('Need one of n_iteration or (atol, rtol).')
            dom, cod = self.compute_domain_codomain()
            if not _plain_l2(dom, cod):
                # general p -> q norm: host-driven iteration over the GEMM kernels; sigma below as always
                with torch.no_grad():
                    self._iterate_general(self.weight.detach(), dom, cod, n_iterations, atol, rtol)
            else:
                only = UPDATE_ONLY['on'] and not torch.is_grad_enabled()
                with torch.no_grad():
                    sig, _ = ops.sn_power_iter(self.weight.detach(), self.u, self.v, n_iterations, atol, rtol,
                                               sigma_out=self.scale if only else None)
                if only:         # update_lipschitz: the kernel left sigma in `scale`; nobody reads the weight
                    self._scale_for = _state_key(self)
                    return None
                if not torch.is_grad_enabled():
                    out = ops.sn_rescale(self.weight.detach(), sig, self.coeff, scale_out=self.scale)
                    self._scale_for = _state_key(self)
                    return out
        sigma = _Sigma.apply(self.weight, self.u, self.v)
        with torch.no_grad():
            self.scale.copy_(sigma[0])
        return _soft_rescale(self.weight, sigma, self.coeff)

    def forward(self, input):
        weight = self.compute_weight(update=False)
        shape = input.shape
        y = ops.linear(input.reshape(-1, shape[-1]), weight, self.bias)
        return y.view(*shape[:-1], self.out_features)

    def extra_repr(self):
        return ('in_features={}, out_features={}, bias={}, coeff={}, domain={:.2f}, codomain={:.2f}, n_iters={}, '
                'atol={}, rtol={}'.format(self.in_features, self.out_features, self.bias is not None, self.coeff,
                                          self.domain, self.codomain, self.n_iterations, self.atol, self.rtol))


def _pair(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _to_nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _from_nhwc(y):
    return y.permute(0, 3, 1, 2)


class InducedNormConv2d(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, bias=True, coeff=0.97, domain=2,
                 codomain=2, n_iterations=None, atol=None, rtol=None, **unused_kwargs):
        del unused_kwargs
        super(InducedNormConv2d, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride = _pair(stride)
        self.padding = _pair(padding)
        if self.kernel_size not in ((1, 1), (3, 3)) or self.stride != (1, 1) or \
                self.padding != (self.kernel_size[0] // 2,) * 2:
            raise NotImplementedError('impflow_b200: conv kernels are 1x1 or 3x3, stride 1, same padding '
                                      '(all shipped configs); got k=%s s=%s p=%s'
                                      % (self.kernel_size, self.stride, self.padding))
        self.coeff = coeff
        self.n_iterations = n_iterations
        self.domain = domain
        self.codomain = codomain
        self.atol = atol
        self.rtol = rtol
        self.weight = nn.Parameter(torch.Tensor(out_channels, in_channels, *self.kernel_size))
        if bias:
            self.bias = nn.Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()
        self.register_buffer('initialized', torch.tensor(0))
        self.register_buffer('spatial_dims', torch.tensor([1., 1.]))
        self.register_buffer('scale', torch.tensor(0.))
        self.register_buffer('u', self.weight.new_empty(self.out_channels))
        self.register_buffer('v', self.weight.new_empty(self.in_channels))
        self._hw = None
        self._init_known = None     # python mirror of the `initialized` buffer (reading it syncs the stream)

    def compute_domain_codomain(self):
        if torch.is_tensor(self.domain):          # learnable orders (learn_p)
            return asym_squash(self.domain), asym_squash(self.codomain)
        return self.domain, self.codomain

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    # --- helpers -------------------------------------------------------------------------------
    def is_initialized(self):
        """`bool(self.initialized)` without a device->host read on every call."""
        if self._init_known is None or not self._init_known:
            self._init_known = bool(self.initialized)
        return self._init_known

    def _spatial(self):
        if self._hw is None:
            self._hw = (int(self.spatial_dims[0].item()), int(self.spatial_dims[1].item()))
        return self._hw

    def _load_from_state_dict(self, *args, **kwargs):
        self._hw = None
        self._init_known = None
        return super(InducedNormConv2d, self)._load_from_state_dict(*args, **kwargs)

    def _conv_vec(self, vec_chw, weight, transpose=False):
        """3x3 conv (or its adjoint) applied to one image given/returned as a flat CHW vector."""
        h, w = self._spatial()
        if not transpose:
            x = vec_chw.view(1, self.in_channels, h, w)
            y = ops.conv3x3_nhwc(_to_nhwc(x), weight)
        else:
            x = vec_chw.view(1, self.out_channels, h, w)
            wt = weight.flip(2, 3).transpose(0, 1)        # conv_transpose2d, stride 1, pad 1
            y = ops.conv3x3_nhwc(_to_nhwc(x), wt)
        return _from_nhwc(y).reshape(-1)

    def sigma_gradient(self):
        """D = d sigma / d W in the weight's layout; sigma = <u, conv(v; W)> is linear in W, so D
        depends on u and v only (the correlation of u with the patches of v) and is cached until they
        change (once per update_lipschitz)."""
        key = (self.u._version, self.v._version, self.u.data_ptr(), self.v.data_ptr())
        cached = getattr(self, '_sigma_grad', None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                if self.kernel_size == (1, 1):
                    D = torch.outer(self.u, self.v).view(self.out_channels, self.in_channels, 1, 1).contiguous()
                else:
                    h, w = self._spatial()
                    co, ci = self.out_channels, self.in_channels
                    v_nhwc = self.v.view(1, ci, h, w).permute(0, 2, 3, 1).contiguous()
                    col_t = ops.transpose2d(ops.im2col3x3(v_nhwc))                 # (9ci, hw)
                    Dr, _, _ = ops.gemm_nt(self.u.view(co, h * w), col_t)          # (co, (ky,kx,ci))
                    D = Dr.view(co, 3, 3, ci).permute(0, 3, 1, 2).contiguous()
                cached = self._sigma_grad = (key, D)
        return cached[1]

    def _initialize_u_v(self):
        # mixed_lipschitz.py:195-239: one start for domain = codomain = 2, ten more random starts otherwise
        with torch.no_grad():
            dom, cod = self.compute_domain_codomain()
            if self.kernel_size == (1, 1):
                n_u, n_v = self.out_channels, self.in_channels
            else:
                h, w = self._spatial()
                n_u, n_v = self.out_channels * h * w, self.in_channels * h * w
            # the reference draws v first for k x k kernels, u first for 1 x 1 ones
            if self.kernel_size == (1, 1):
                self.u.resize_(n_u).normal_(0, 1)
                self.u.copy_(normalize_u(self.u, cod))
                self.v.resize_(n_v).normal_(0, 1)
                self.v.copy_(normalize_v(self.v, dom))
            else:
                self.v.resize_(n_v).normal_(0, 1)
                self.v.copy_(normalize_v(self.v, dom))
                self.u.resize_(n_u).normal_(0, 1)
                self.u.copy_(normalize_u(self.u, cod))
            self.initialized.fill_(1)
            self._init_known = True
            self.compute_weight(True)
            if not _plain_l2(dom, cod):
                # restarts keep the start whose sigma beats the FIRST one (best_scale is never raised, :223-235)
                best_scale = self.scale.clone()
                best_u, best_v = self.u.clone(), self.v.clone()
                for _ in range(10):
                    if self.kernel_size == (1, 1):
                        self.u.copy_(normalize_u(self.weight.new_empty(n_u).normal_(0, 1), cod))
                        self.v.copy_(normalize_v(self.weight.new_empty(n_v).normal_(0, 1), dom))
                    else:     # host generator, as in the reference (torch.randn(...).to(weight))
                        self.u.copy_(normalize_u(torch.randn(n_u).to(self.weight), cod))
                        self.v.copy_(normalize_v(torch.randn(n_v).to(self.weight), dom))
                    self.compute_weight(True, n_iterations=200)
                    if self.scale > best_scale:
                        best_u, best_v = self.u.clone(), self.v.clone()
                self.u.copy_(best_u)
                self.v.copy_(best_v)
            self.u = self.u.clone(memory_format=torch.contiguous_format)
            self.v = self.v.clone(memory_format=torch.contiguous_format)

    def _general_update(self, dom, cod, n_iterations, atol, rtol):
        """General p -> q power iteration on the buffers (mixed_lipschitz.py:293-318, :343-372): W v / W^T u on the
        GEMM / conv kernels, dual-norm normalisations in between.  Which buffer is written back follows the
        reference branch by branch (a 1x1 layer keeps its stored v for domain 1 and its stored u for codomain inf)."""
        if self.kernel_size == (1, 1):
            W2 = self.weight.detach().view(self.out_channels, self.in_channels)
            fw, bw = (lambda x: _mv(W2, x)), (lambda x: _mtv(W2, x))
        else:
            wt = self.weight.detach()
            fw, bw = (lambda x: self._conv_vec(x, wt)), (lambda x: self._conv_vec(x, wt, transpose=True))
        u, v, used = _general_power_iteration(fw, bw, self.u.clone(), self.v.clone(), dom, cod, n_iterations, atol,
                                              rtol)
        if used > 0:
            if self.kernel_size == (1, 1):
                keep_v, keep_u = (dom == 1 or _is_two(dom)), (_is_two(cod) or cod == float('inf'))
            else:
                keep_v, keep_u = _is_two(dom), _is_two(cod)
            # a 2-norm side is normalised IN PLACE in the reference (F.normalize(out=buffer)): written either way
            if not keep_v or _is_two(dom):
                self.v.copy_(v)
            if not keep_u or _is_two(cod):
                self.u.copy_(u)
        return u, v

    def compute_one_iter(self):
        if not self.is_initialized():
            raise ValueError('Layer needs to be initialized first.')
        _require_cuda(self.weight, 'InducedNormConv2d.weight')
        dom, cod = self.compute_domain_codomain()
        if not _plain_l2(dom, cod):
            with torch.no_grad():
                if self.kernel_size == (1, 1):
                    W2 = self.weight.detach().view(self.out_channels, self.in_channels)
                    fw, bw = (lambda x: _mv(W2, x)), (lambda x: _mtv(W2, x))
                else:
                    wt = self.weight.detach()
                    fw, bw = (lambda x: self._conv_vec(x, wt)), (lambda x: self._conv_vec(x, wt, transpose=True))
                u = normalize_u(fw(self.v.detach()), cod)
                v = normalize_v(bw(u), dom)
                return torch.dot(u, fw(v))
        if self.kernel_size == (1, 1):
            W2 = self.weight.detach().view(self.out_channels, self.in_channels)
            sigma, _ = ops.sn_power_iter(W2, self.u.clone(), self.v.clone(), 1, 0.0, 0.0)
            return sigma[0]
        with torch.no_grad():
            wt = self.weight.detach()
            u = F.normalize(self._conv_vec(self.v, wt), dim=0)
            v = F.normalize(self._conv_vec(u, wt, transpose=True), dim=0)
            return torch.dot(u, self._conv_vec(v, wt))

    def compute_weight(self, update=True, n_iterations=None, atol=None, rtol=None):
        _require_cuda(self.weight, 'InducedNormConv2d.weight')
        if not self.is_initialized():
            self._initialize_u_v()
        n_iterations = self.n_iterations if n_iterations is None else n_iterations
        atol = self.atol if atol is None else atol
        rtol = self.rtol if rtol is None else atol          # reference quirk (:279, :331)
        if n_iterations is None and (atol is None or rtol is None):
            raise ValueError('Need one of n_iteration or (atol, rtol).')
        if self.kernel_size == (1, 1):
            return self._compute_weight_1x1(update, n_iterations, atol, rtol)
        return self._compute_weight_kxk(update, n_iterations, atol, rtol)

    def _compute_weight_1x1(self, update, n_iterations, atol, rtol):
        W2 = self.weight.view(self.out_channels, self.in_channels)
        dom, cod = self.compute_domain_codomain()
        if not _plain_l2(dom, cod):
            u, v = self.u, self.v
            if update:
                with torch.no_grad():
                    u, v = self._general_update(dom, cod, n_iterations, atol, rtol)
            # sigma from the iterates of THIS call (the stored buffers may lag behind them, see _general_update)
            sigma = _Sigma.apply(W2, u.detach().contiguous(), v.detach().contiguous())
            with torch.no_grad():
                self.scale.copy_(sigma[0])
            return _soft_rescale(W2, sigma, self.coeff).view(self.out_channels, self.in_channels, 1, 1)
        if update:
            only = UPDATE_ONLY['on'] and not torch.is_grad_enabled()
            with torch.no_grad():
                sig, _ = ops.sn_power_iter(W2.detach(), self.u, self.v, n_iterations, atol, rtol,
                                           sigma_out=self.scale if only else None)
            if only:
                self._scale_for = _state_key(self)
                return None
            if not torch.is_grad_enabled():
                out = ops.sn_rescale(W2.detach(), sig, self.coeff, scale_out=self.scale).view(
                    self.out_channels, self.in_channels, 1, 1)
                self._scale_for = _state_key(self)
                return out
        sigma = _Sigma.apply(W2, self.u, self.v)
        with torch.no_grad():
            self.scale.copy_(sigma[0])
        return _soft_rescale(W2, sigma, self.coeff).view(self.out_channels, self.in_channels, 1, 1)

    def _compute_weight_kxk(self, update, n_iterations, atol, rtol):
        u, v = self.u, self.v
        dom, cod = self.compute_domain_codomain()
        if not _plain_l2(dom, cod):
            if update:
                with torch.no_grad():
                    u, v = self._general_update(dom, cod, n_iterations, atol, rtol)
            weight_v = self._conv_vec(v.detach(), self.weight)
            sigma = ops.rowdot_fn(u.detach().view(1, -1), weight_v.view(1, -1))
            with torch.no_grad():
                self.scale.copy_(sigma[0])
            return _soft_rescale(self.weight, sigma, self.coeff)
        if update and self.kernel_size == (3, 3):
            # the whole iteration (and sigma) in one cooperative launch; no host synchronisation
            h, w = self._spatial()
            tol_mode = n_iterations is None and atol is not None and rtol is not None
            res = None
            only = UPDATE_ONLY['on'] and not torch.is_grad_enabled()
            if tol_mode or n_iterations is not None:
                with torch.no_grad():
                    res = ops.sn_power_iter_conv(self.weight.detach(), self.u, self.v, h, w,
                                                 None if tol_mode else n_iterations, atol, rtol, want_D=True,
                                                 sigma_out=self.scale if only else None)
            if res is not None:
                update = False
                # the kernel also leaves d sigma / d W for the new (u, v): what sigma_gradient() would recompute
                self._sigma_grad = ((self.u._version, self.v._version, self.u.data_ptr(), self.v.data_ptr()), res[2])
                if only:         # update_lipschitz: sigma went straight into `scale`; nobody reads the weight
                    self._scale_for = _state_key(self)
                    return None
                if not torch.is_grad_enabled():      # nothing differentiates this result
                    out = ops.sn_rescale(self.weight.detach(), res[0], self.coeff, scale_out=self.scale)
                    self._scale_for = _state_key(self)      # `scale` now holds sigma of exactly this state
                    return out
        if update:
            max_itrs = 200 if n_iterations is None else n_iterations
            with torch.no_grad():
                wt = self.weight.detach()
                for _ in range(max_itrs):
                    old_u, old_v = u, v
                    u = F.normalize(self._conv_vec(v, wt), dim=0)
                    v = F.normalize(self._conv_vec(u, wt, transpose=True), dim=0)
                    if n_iterations is None and atol is not None and rtol is not None:
                        err_u = torch.norm(u - old_u) / (u.nelement() ** 0.5)
                        err_v = torch.norm(v - old_v) / (v.nelement() ** 0.5)
                        tol_u = atol + rtol * torch.max(u)
                        tol_v = atol + rtol * torch.max(v)
                        if bool((err_u < tol_u) & (err_v < tol_v)):
                            break
                self.u.copy_(u)
                self.v.copy_(v)
                u, v = self.u, self.v
        weight_v = self._conv_vec(v, self.weight)
        sigma = ops.rowdot_fn(u.view(1, -1), weight_v.view(1, -1))
        with torch.no_grad():
            self.scale.copy_(sigma[0])
        return _soft_rescale(self.weight, sigma, self.coeff)

    def forward(self, input):
        _require_cuda(input, 'InducedNormConv2d input')
        if not self.is_initialized():
            self.spatial_dims.copy_(torch.tensor(input.shape[2:4]).to(self.spatial_dims))
            self._hw = None
        weight = self.compute_weight(update=False)
        x = _to_nhwc(input)
        if self.kernel_size == (1, 1):
            y = ops.conv1x1_nhwc(x, weight, self.bias)
        else:
            y = ops.conv3x3_nhwc(x, weight, self.bias)
        return _from_nhwc(y)

    def extra_repr(self):
        s = '{}, {}, kernel_size={}, stride={}'.format(self.in_channels, self.out_channels, self.kernel_size,
                                                       self.stride)
        if self.bias is None:
            s += ', bias=False'
        s += ', coeff={}, domain={:.2f}, codomain={:.2f}, n_iters={}, atol={}, rtol={}'.format(
            self.coeff, self.domain, self.codomain, self.n_iterations, self.atol, self.rtol)
        return s


_side_streams = {}
UPDATE_ONLY = {'on': False}     # set by update_lipschitz: compute_weight(update=True) only refreshes u, v and sigma


def update_lipschitz(model, n_iterations=None, n_streams=8):
    """The per-step power-iteration refresh of every induced-norm layer (train_img.py:786-792,
    train_toy.py:174-179 with n_iterations): `compute_weight(update=True)` under no_grad.

    The layers are independent and each refresh is one or two latency-bound launches (a single-CTA matrix
    power iteration or the cooperative conv one), so they are fanned out over side streams and joined
    back into the current stream.  The frozen `*_copy` twins are skipped: imBlock.forward overwrites
    them from the live nets before their next use."""
    # the walk over every sub-module costs ~1 ms of host time per step at the CIFAR model: keep the list on the
    # model (rebuilt when sub-modules were added or removed)
    cached = model.__dict__.get('_lipschitz_layers')
    n_children = sum(1 for _ in model.children())
    if cached is None or cached[0] != n_children:
        mods = [m for name, m in model.named_modules()
                if '_copy' not in name and isinstance(m, (InducedNormConv2d, InducedNormLinear))]
        model.__dict__['_lipschitz_layers'] = (n_children, mods)
    else:
        mods = cached[1]
    if not mods:
        return
    dev = mods[0].weight.device
    UPDATE_ONLY['on'] = True
    try:
        _update_all(mods, dev, n_iterations, n_streams)
    finally:
        UPDATE_ONLY['on'] = False


BATCH_DENSE = {'on': True}      # update_lipschitz: all dense (Linear / 1x1) 2-norm layers in one launch

_desc_cache = {}


def _dense_settings(m, n_iterations):
    """(n_iterations or -1, atol, rtol) a dense layer's compute_weight(update=True) would run with, or None if the
    layer does not qualify for the batched launch."""
    if isinstance(m, InducedNormConv2d):
        if m.kernel_size != (1, 1) or not m.is_initialized():
            return None
    elif not isinstance(m, InducedNormLinear):
        return None
    if not _plain_l2(*m.compute_domain_codomain()):
        return None
    n_it = m.n_iterations if n_iterations is None else n_iterations
    atol = m.atol
    rtol = m.rtol if atol is None else atol             # the reference's quirk (mixed_lipschitz.py:94, :279)
    if n_it is None and (atol is None or rtol is None):
        return None                                      # the per-layer path raises the reference's ValueError
    return (-1 if n_it is None else int(n_it), float(atol or 0.0), float(rtol or 0.0))


def _update_dense_batched(mods, n_iterations):
    """Power iteration of every qualifying dense layer in ONE launch per (n_iterations, atol, rtol) group
    (csrc/spectral.cu k_sn_power_iter_batch, one CTA per layer).  Returns (the modules it does NOT handle, the
    launches as closures): the caller issues them on the current stream AFTER it has started the other layers on the
    side streams, so that the handful of single-CTA problems run next to the 3x3 layers' launches."""
    groups, rest = {}, []
    for m in mods:
        key = _dense_settings(m, n_iterations) if BATCH_DENSE['on'] else None
        if key is None:
            rest.append(m)
        else:
            groups.setdefault(key, []).append(m)
    lib = _cabi.load()
    launches = []
    for key, ms in groups.items():
        if len(ms) < 2:
            rest += ms
        else:
            launches.append(lambda key=key, ms=ms: _launch_dense_group(lib, key, ms))
    return rest, launches


def _launch_dense_group(lib, key, ms):
    """One k_sn_power_iter_batch launch for the layers `ms` (all with the settings `key`); the descriptor table lives
    on the device, cached by the layers' buffer addresses."""
    import ctypes
    import numpy as np
    n_it, atol, rtol = key
    Ws = [m.weight.detach().view(m.weight.shape[0], -1) for m in ms]
    ck = tuple((W.data_ptr(), m.u.data_ptr(), m.v.data_ptr(), m.scale.data_ptr(), W.shape) for W, m in zip(Ws, ms))
    cached = _desc_cache.get(ck)
    if cached is None:
        if len(_desc_cache) > 64:
            _desc_cache.clear()
        iters = torch.zeros(len(ms), device=Ws[0].device, dtype=torch.int32)
        host = (_cabi.SnDesc * len(ms))()
        for i, (W, m) in enumerate(zip(Ws, ms)):
            assert W.is_contiguous() and m.u.is_contiguous() and m.v.is_contiguous()
            host[i].W, host[i].u, host[i].v = W.data_ptr(), m.u.data_ptr(), m.v.data_ptr()
            host[i].sigma, host[i].iters = m.scale.data_ptr(), iters.data_ptr() + 4 * i
            host[i].out_f, host[i].in_f = W.shape[0], W.shape[1]
        raw = np.frombuffer(bytes(host), dtype=np.uint8).copy()
        descs = torch.from_numpy(raw).to(Ws[0].device)
        cached = _desc_cache[ck] = (descs, iters, max(W.shape[0] for W in Ws), max(W.shape[1] for W in Ws))
    descs, iters, max_out, max_in = cached
    _cabi.check(lib.impflow_sn_power_iter_batch(ctypes.c_void_p(descs.data_ptr()), len(ms), max_out, max_in, n_it,
                                                atol, rtol, _cabi.stream()), 'sn_power_iter_batch')
    touched = []
    for m in ms:
        touched += [m.u, m.v, m.scale] if n_it != 0 else [m.scale]
    torch.autograd.graph.increment_version(touched)      # written through raw pointers: version-keyed host caches
    for m in ms:
        m._scale_for = _state_key(m)


def _update_all(mods, dev, n_iterations, n_streams):
    with torch.no_grad():
        batched = []
        if UPDATE_ONLY['on']:
            _require_cuda(mods[0].weight, 'update_lipschitz: layer weights')
            mods, batched = _update_dense_batched(mods, n_iterations)
        ready = all((not isinstance(m, InducedNormConv2d)) or m.is_initialized() for m in mods)
        if dev.type != 'cuda' or n_streams <= 1 or len(mods) < 2 or not ready:
            for launch in batched:
                launch()
            for m in mods:
                m.compute_weight(update=True, n_iterations=n_iterations)
            return
        pool = _side_streams.setdefault(dev.index, [])
        while len(pool) < n_streams:
            pool.append(torch.cuda.Stream(device=dev))
        used = pool[:min(n_streams, len(mods))]
        main = torch.cuda.current_stream(dev)
        start = torch.cuda.Event()
        start.record(main)
        for s in used:
            s.wait_event(start)
        for i, m in enumerate(mods):
            with torch.cuda.stream(used[i % len(used)]):
                m.compute_weight(update=True, n_iterations=n_iterations)
        for launch in batched:          # on the main stream, next to the side streams' work
            launch()
        for s in used:
            done = torch.cuda.Event()
            done.record(s)
            main.wait_event(done)


def sigma_of(m):
    """Device scalar sigma = <W, d sigma/dW> of the layer's CURRENT weight (what compute_weight(update=False)
    uses).  When the weight, u and v are untouched since the last power iteration the value the kernel left in
    `m.scale` is that number and nothing is launched."""
    if getattr(m, '_scale_for', None) == _state_key(m):
        return m.scale
    with torch.no_grad():
        sigma = ops._flat_dot(m.weight.detach().contiguous(), m.sigma_gradient())
        m.scale.copy_(sigma[0])
    return sigma
