"""ImplicitFlow: multiscale stack of imBlocks — API mirror of lib/implicit_flow.py
(ImplicitFlow :20-251, StackedImplicitBlocks :254-434, FCNet :437-474, FCWrapper :477-501).

This module is a *caller* of the hot path: it only wires imBlocks, ActNorm,
Squeeze and LogitTransform together.  Options that are dead in the reference (quadratic,
batchnorm — SURVEY.md quirk #20) or outside the path (learn_p, dropout) raise."""
import numpy as np
import torch
import torch.nn as nn

from . import layers
from .layers import base as base_layers

ACT_FNS = {
    'softplus': lambda b: nn.Softplus(),
    'elu': lambda b: nn.ELU(inplace=b),
    'swish': lambda b: base_layers.Swish(),
    'identity': lambda b: base_layers.Identity(),
    'relu': lambda b: base_layers.ReLU(),
    'sin': lambda b: base_layers.Sin(),
    'zero': lambda b: base_layers.Zero(),
}


def _parse_vnorms(vnorms):
    ps = [float('inf') if p == 'f' else float(p) for p in vnorms]
    return ps[:-1], ps[1:]


class ImplicitFlow(nn.Module):
    _stack = None      # set below: StackedImplicitBlocks (lib/resflow.py's ResidualFlow uses StackediResBlocks)

    def __init__(self, input_size, n_blocks=[16, 16], intermediate_dim=64, factor_out=True, quadratic=False,
                 init_layer=None, actnorm=False, fc_actnorm=False, batchnorm=False, dropout=0, fc=False, coeff=0.9,
                 vnorms='122f', n_lipschitz_iters=None, sn_atol=None, sn_rtol=None, n_power_series=5,
                 n_dist='geometric', n_samples=1, kernels='3-1-3', activation_fn='elu', fc_end=True, fc_idim=128,
                 n_exact_terms=0, preact=False, neumann_grad=True, grad_in_forward=False, first_resblock=True,
                 learn_p=False, classification=False, classification_hdim=64, n_classes=10):
        super(ImplicitFlow, self).__init__()
        self.n_scale = min(len(n_blocks), self._calc_n_scale(input_size))
        self.n_blocks = n_blocks
        self.intermediate_dim = intermediate_dim
        self.factor_out = factor_out
        # every constructor argument is kept as an attribute, as in the reference (implicit_flow.py:58-89); the init
        # layer is thereby ALSO registered as `init_layer` (its buffers appear under both names in a state dict)
        self.quadratic, self.init_layer, self.actnorm, self.fc_actnorm = quadratic, init_layer, actnorm, fc_actnorm
        self.batchnorm, self.dropout, self.fc, self.coeff, self.vnorms = batchnorm, dropout, fc, coeff, vnorms
        self.n_lipschitz_iters, self.sn_atol, self.sn_rtol = n_lipschitz_iters, sn_atol, sn_rtol
        self.n_power_series, self.n_dist, self.n_samples, self.kernels = n_power_series, n_dist, n_samples, kernels
        self.activation_fn, self.fc_end, self.fc_idim, self.n_exact_terms = activation_fn, fc_end, fc_idim, n_exact_terms
        self.preact, self.neumann_grad, self.grad_in_forward = preact, neumann_grad, grad_in_forward
        self.first_resblock, self.learn_p = first_resblock, learn_p
        self.classification = classification
        self.classification_hdim = classification_hdim
        self.n_classes = n_classes
        if not self.n_scale > 0:
            raise ValueError('Could not compute number of scales for input of size (%d,%d,%d,%d)' % input_size)
        shared = dict(idim=intermediate_dim, quadratic=quadratic, actnorm=actnorm, fc_actnorm=fc_actnorm,
                      batchnorm=batchnorm, dropout=dropout, fc=fc, coeff=coeff, vnorms=vnorms,
                      n_lipschitz_iters=n_lipschitz_iters, sn_atol=sn_atol, sn_rtol=sn_rtol,
                      n_power_series=n_power_series, n_dist=n_dist, n_samples=n_samples, kernels=kernels,
                      activation_fn=activation_fn, fc_end=fc_end, fc_idim=fc_idim, n_exact_terms=n_exact_terms,
                      preact=preact, neumann_grad=neumann_grad, grad_in_forward=grad_in_forward, learn_p=learn_p)
        _, c, h, w = input_size
        transforms = []
        for i in range(self.n_scale):
            transforms.append(self._stack(
                initial_size=(c, h, w), squeeze=(i < self.n_scale - 1), init_layer=init_layer if i == 0 else None,
                n_blocks=n_blocks[i], first_resblock=first_resblock and (i == 0), **shared))
            c, h, w = (c * 2 if factor_out else c * 4), h // 2, w // 2
        self.transforms = nn.ModuleList(transforms)
        self.dims = [o[1:] for o in self.calc_output_size(input_size)]
        if self.classification:
            self.build_multiscale_classifier(input_size)

    def _calc_n_scale(self, input_size):
        _, _, h, w = input_size
        n_scale = 0
        while h >= 4 and w >= 4:
            n_scale += 1
            h, w = h // 2, w // 2
        return n_scale

    def calc_output_size(self, input_size):
        n, c, h, w = input_size
        if not self.factor_out:
            k = self.n_scale - 1
            return [[n, c * 4 ** k, h // 2 ** k, w // 2 ** k]]
        sizes = []
        for i in range(self.n_scale):
            if i < self.n_scale - 1:
                c, h, w = c * 2, h // 2, w // 2
            sizes.append((n, c, h, w))
        return tuple(sizes)

    def build_multiscale_classifier(self, input_size):
        n, c, h, w = input_size
        heads = []
        for i in range(self.n_scale):
            if i < self.n_scale - 1:
                c *= 2 if self.factor_out else 4
                h //= 2
                w //= 2
            heads.append(nn.Sequential(nn.Conv2d(c, self.classification_hdim, 3, 1, 1),
                                       layers.ActNorm2d(self.classification_hdim), nn.ReLU(inplace=True),
                                       nn.AdaptiveAvgPool2d((1, 1))))
        self.classification_heads = nn.ModuleList(heads)
        self.logit_layer = nn.Linear(self.classification_hdim * len(heads), self.n_classes)

    def forward(self, x, logpx=None, inverse=False, classify=False, restore=False):
        if inverse:
            return self.inverse(x, logpx)
        out, class_outs = [], []
        for idx, tr in enumerate(self.transforms):
            if logpx is not None:
                x, logpx = tr.forward(x, logpx, restore=restore)
            else:
                x = tr.forward(x, restore=restore)
            f = None
            if self.factor_out and (idx < len(self.transforms) - 1):
                d = x.size(1) // 2
                x, f = x[:, :d], x[:, d:]
                out.append(f)
            if classify:
                class_outs.append(self.classification_heads[idx](f if self.factor_out else x))
        out.append(x)
        out = torch.cat([o.reshape(o.size()[0], -1) for o in out], 1)
        output = out if logpx is None else (out, logpx)
        if classify:
            hcat = torch.cat(class_outs, dim=1).squeeze(-1).squeeze(-1)
            return output, self.logit_layer(hcat)
        return output

    def inverse(self, z, logpz=None):
        if self.factor_out:
            z = z.view(z.shape[0], -1)
            zs, i = [], 0
            for dims in self.dims:
                s = int(np.prod(dims))
                zs.append(z[:, i:i + s].view(z.size(0), *dims))
                i += s
            if logpz is None:
                z_prev = self.transforms[-1].inverse(zs[-1])
                for idx in range(len(self.transforms) - 2, -1, -1):
                    z_prev = self.transforms[idx].inverse(torch.cat((z_prev, zs[idx]), dim=1))
                return z_prev
            z_prev, logpz = self.transforms[-1].inverse(zs[-1], logpz)
            for idx in range(len(self.transforms) - 2, -1, -1):
                z_prev, logpz = self.transforms[idx].inverse(torch.cat((z_prev, zs[idx]), dim=1), logpz)
            return z_prev, logpz
        z = z.view(z.shape[0], *self.dims[-1])
        for idx in range(len(self.transforms) - 1, -1, -1):
            if logpz is None:
                z = self.transforms[idx].inverse(z)
            else:
                z, logpz = self.transforms[idx].inverse(z, logpz)
        return z if logpz is None else (z, logpz)


class StackedImplicitBlocks(layers.SequentialFlow):
    _implicit = True       # imBlock(nnet_x, nnet_z); StackediResBlocks (resflow.py) builds iResBlock(nnet)

    def __init__(self, initial_size, idim, squeeze=True, init_layer=None, n_blocks=1, quadratic=False, actnorm=False,
                 fc_actnorm=False, batchnorm=False, dropout=0, fc=False, coeff=0.9, vnorms='122f',
                 n_lipschitz_iters=None, sn_atol=None, sn_rtol=None, n_power_series=5, n_dist='geometric', n_samples=1,
                 kernels='3-1-3', activation_fn='elu', fc_end=True, fc_nblocks=None, fc_idim=128, n_exact_terms=0,
                 preact=False, neumann_grad=True, grad_in_forward=False, first_resblock=True, learn_p=False):
        if quadratic:
            # implicit_flow.py:309-315: Glow's InvertibleConv2d / InvertibleLinear in front of every block; those
            # layers belong to the Glow baseline (lib/layers/glow.py), no run script sets --quadratic
            raise NotImplementedError('impflow_b200: quadratic=True (Glow invertible 1x1 layers between the blocks) is '
                                      'outside the hot-path scope')
        if fc_nblocks is None:
            fc_nblocks = 2 if self._implicit else 4       # implicit_flow.py:280, resflow.py:281
        domains, codomains = _parse_vnorms(vnorms)
        ks = list(map(int, kernels.split('-')))
        assert len(domains) == len(ks)
        block_kw = dict(n_power_series=n_power_series, n_dist=n_dist, n_samples=n_samples,
                        n_exact_terms=n_exact_terms, neumann_grad=neumann_grad, grad_in_forward=grad_in_forward)
        lip_kw = dict(coeff=coeff, n_iterations=n_lipschitz_iters, atol=sn_atol, rtol=sn_rtol)

        def _actnorm(size, as_fc):
            if as_fc:
                return FCWrapper(layers.ActNorm1d(size[0] * size[1] * size[2]))
            return layers.ActNorm2d(size[0])

        def conv_branch(leading_act):
            # [act] conv(c->idim) act conv(idim->idim)... act conv(idim->c)   (implicit_flow.py:359-398)
            chans = [initial_size[0]] + [idim] * (len(ks) - 1) + [initial_size[0]]
            doms, cods = domains, codomains
            if learn_p:       # learnable orders, one per layer boundary, shared by neighbours (implicit_flow.py:364-366)
                doms = [nn.Parameter(torch.tensor(0.)) for _ in range(len(ks))]
                cods = doms[1:] + [doms[0]]
            # batchnorm / dropout (implicit_flow.py:374-396; no run script sets them): mean-only MovingBatchNorm2d
            # after every conv (and in front of a pre-activation), Dropout2d in front of the last conv.  Such a branch
            # has no graph-free program: it runs through the module / autograd path over the same conv kernels.
            mods = []
            if leading_act:
                if batchnorm:
                    mods.append(layers.MovingBatchNorm2d(chans[0]))
                mods.append(ACT_FNS[activation_fn](False))
            for i, k in enumerate(ks):
                if i > 0:
                    mods.append(ACT_FNS[activation_fn](True))
                if dropout and i == len(ks) - 1:
                    mods.append(nn.Dropout2d(dropout))      # the reference's is in place; same draws, same values
                mods.append(base_layers.get_conv2d(chans[i], chans[i + 1], k, 1, k // 2, domain=doms[i],
                                                   codomain=cods[i], **lip_kw))
                if batchnorm:
                    mods.append(layers.MovingBatchNorm2d(chans[i + 1]))
            return nn.Sequential(*mods)

        def fc_net(width):
            return FCNet(input_shape=initial_size, idim=width, lipschitz_layer=base_layers.get_linear,
                         nhidden=len(ks) - 1, coeff=coeff, domains=domains, codomains=codomains,
                         n_iterations=n_lipschitz_iters, activation_fn=activation_fn, preact=preact, dropout=dropout,
                         sn_atol=sn_atol, sn_rtol=sn_rtol, learn_p=learn_p)

        def _resblock(as_fc, width=idim, first=True):
            lead = (not first) and preact
            net = (lambda: fc_net(width)) if as_fc else (lambda: conv_branch(lead))
            if self._implicit:
                return layers.imBlock(net(), net(), **block_kw)
            return layers.iResBlock(net(), **block_kw)

        chain = []
        if init_layer is not None:
            chain.append(init_layer)
        if first_resblock and actnorm:
            chain.append(_actnorm(initial_size, fc))
        if first_resblock and fc_actnorm:
            chain.append(_actnorm(initial_size, True))
        for i in range(n_blocks):
            chain.append(_resblock(fc, first=first_resblock and (i == 0)))
            if actnorm:
                chain.append(_actnorm(initial_size, fc))
            if fc_actnorm:
                chain.append(_actnorm(initial_size, True))
        if squeeze:
            chain.append(layers.SqueezeLayer(2))
        elif fc_end:
            for _ in range(fc_nblocks):
                chain.append(_resblock(True, fc_idim))
                if actnorm or fc_actnorm:
                    chain.append(_actnorm(initial_size, True))
        super(StackedImplicitBlocks, self).__init__(chain)


ImplicitFlow._stack = StackedImplicitBlocks


class FCNet(nn.Module):

    def __init__(self, input_shape, idim, lipschitz_layer, nhidden, coeff, domains, codomains, n_iterations,
                 activation_fn, preact, dropout, sn_atol, sn_rtol, learn_p, div_in=1):
        super(FCNet, self).__init__()
        if learn_p:           # implicit_flow.py:450-452
            domains = [nn.Parameter(torch.tensor(0.)) for _ in range(len(domains))]
            codomains = domains[1:] + [domains[0]]
        self.input_shape = input_shape
        c, h, w = self.input_shape
        dim = c * h * w
        widths = [dim // div_in] + [idim] * nhidden + [dim]
        mods = []
        if preact:
            mods.append(ACT_FNS[activation_fn](False))
        for i in range(nhidden + 1):
            if i > 0:
                mods.append(ACT_FNS[activation_fn](True))
            j = min(i, len(domains) - 1) if i < nhidden else -1
            if dropout and i == nhidden:
                mods.append(nn.Dropout(dropout))            # implicit_flow.py:463 (in place there)
            mods.append(lipschitz_layer(widths[i], widths[i + 1], coeff=coeff, n_iterations=n_iterations,
                                        domain=domains[j], codomain=codomains[j], atol=sn_atol, rtol=sn_rtol))
        self.nnet = nn.Sequential(*mods)

    def forward(self, x, restore=False):
        y = self.nnet(x.reshape(x.shape[0], -1))
        return y.view(y.shape[0], *self.input_shape)


class FCWrapper(nn.Module):

    def __init__(self, fc_module):
        super(FCWrapper, self).__init__()
        self.fc_module = fc_module

    def forward(self, x, logpx=None, restore=False):
        shape = x.shape
        x = x.reshape(x.shape[0], -1)
        if logpx is None:
            return self.fc_module(x).view(*shape)
        y, logpy = self.fc_module(x, logpx)
        return y.view(*shape), logpy

    def inverse(self, y, logpy=None):
        shape = y.shape
        y = y.reshape(y.shape[0], -1)
        if logpy is None:
            return self.fc_module.inverse(y).view(*shape)
        x, logpx = self.fc_module.inverse(y, logpy)
        return x.view(*shape), logpx
