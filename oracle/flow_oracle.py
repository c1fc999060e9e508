"""CPU oracle, part 2 — TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.

Assembles a whole multiscale flow (LogitTransform, ActNorm2d, imBlock, SqueezeLayer — the
chain built by lib/implicit_flow.py:411-428 for the conv configs) from a reference-keyed state
dict out of the functional pieces in impflow_oracle.py, so that the reference's training step
(train_img.py:517-549, 591-660: forward, bits/dim, backward, Adam, update_lipschitz) can be
timed on host cores as the CPU baseline of bench.py and used as the checker of smoke().
Only bench.py's cpu_baseline / --impl reference legs and tests import this file."""
import torch

from . import impflow_oracle as orc


def _branch_from_sd(sd, prefix, coeff, tol, act):
    keys = [k[len(prefix):] for k in sd if k.startswith(prefix)]
    idxs = sorted({int(k.split('.')[0]) for k in keys})
    layers, pending_act, pending_beta = [], None, None
    for i in range(max(idxs) + 1):
        p = '%s%d.' % (prefix, i)
        if p + 'weight' in sd:
            W = sd[p + 'weight'].detach().clone().float().requires_grad_(True)
            b = sd[p + 'bias'].detach().clone().float().requires_grad_(True) if p + 'bias' in sd else None
            kind = 'linear' if W.dim() == 2 else 'conv'
            spatial = [int(s) for s in sd[p + 'spatial_dims'].tolist()] if p + 'spatial_dims' in sd else None
            layers.append(orc.OracleLayer(kind=kind, weight=W, bias=b, u=sd[p + 'u'].detach().clone().float(),
                                          v=sd[p + 'v'].detach().clone().float(), coeff=coeff, pre_act=pending_act,
                                          beta=pending_beta, padding=(W.shape[-1] // 2 if kind == 'conv' else 0),
                                          n_iterations=None, atol=tol, rtol=tol, spatial=spatial))
            pending_act, pending_beta = None, None
        elif p + 'beta' in sd:
            pending_act = 'swish'
            pending_beta = sd[p + 'beta'].detach().clone().float().requires_grad_(True)
        else:
            pending_act = act
    return orc.OracleBranch(layers=layers)


class OracleFlow(object):
    """Flow = list of ('logit', alpha) | ('actnorm', w, b) | ('imblock', bx, bz) | ('squeeze',)."""

    def __init__(self, state_dict, n_blocks, cfg, coeff=0.9, tol=1e-3, act='swish', logit_alpha=0.05, actnorm=True):
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.cfg = cfg
        self.chain = []
        self.params = []
        for s, nb in enumerate(n_blocks):
            pre = 'transforms.%d.chain.' % s
            i = 0
            if s == 0 and logit_alpha is not None:
                self.chain.append(('logit', logit_alpha))
                i += 1
            if s == 0 and actnorm:
                self._add_actnorm(sd, pre, i)
                i += 1
            for _ in range(nb):
                bx = _branch_from_sd(sd, '%s%d.nnet_x.' % (pre, i), coeff, tol, act)
                bz = _branch_from_sd(sd, '%s%d.nnet_z.' % (pre, i), coeff, tol, act)
                self.chain.append(('imblock', bx, bz))
                self.params += bx.parameters() + bz.parameters()
                i += 1
                if actnorm:
                    self._add_actnorm(sd, pre, i)
                    i += 1
            if s < len(n_blocks) - 1:
                self.chain.append(('squeeze',))

    def _add_actnorm(self, sd, pre, i):
        w = sd['%s%d.weight' % (pre, i)].clone().float().requires_grad_(True)
        b = sd['%s%d.bias' % (pre, i)].clone().float().requires_grad_(True)
        self.chain.append(('actnorm', w, b))
        self.params += [w, b]

    def forward(self, x, training=True, stats=None):
        logp = torch.zeros(x.shape[0], 1)
        for item in self.chain:
            if item[0] == 'logit':
                x, logp = orc.logit_forward(x, item[1], logp)
            elif item[0] == 'actnorm':
                x, logp = orc.actnorm_forward(x, item[1], item[2], logp)
            elif item[0] == 'squeeze':
                x = orc.squeeze2(x)
            else:
                x, logp = orc.imblock_forward(item[1], item[2], x, logp, self.cfg, training, stats=stats)
        return x.reshape(x.shape[0], -1), logp

    def update_lipschitz(self):
        for item in self.chain:
            if item[0] == 'imblock':
                item[1].update_lipschitz()
                item[2].update_lipschitz()

    def train_step(self, x, optimizer, stats=None):
        """train_img.py:591-660 for the density task: bits/dim, backward, clip, Adam, update_lipschitz."""
        optimizer.zero_grad()
        z, dlogp = self.forward(x, True, stats)
        n_dims = x[0].numel()
        bpd = orc.bits_per_dim(z, dlogp, n_dims)
        bpd.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.)
        optimizer.step()
        self.update_lipschitz()
        return float(bpd)


CIFAR_CFG = dict(orc.DEFAULT_CFG, n_dist='poisson', n_exact_terms=10, neumann_grad=True, grad_in_forward=True,
                 lamb=2.0)
