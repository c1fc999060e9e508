"""Shared helpers for the parity tests: build oracle branches from golden state dicts."""
import numpy as np
import torch

from oracle import impflow_oracle as orc


def sub_sd(fix, prefix):
    """{'a.b': tensor} for every fixture key starting with prefix."""
    return {k[len(prefix):]: torch.from_numpy(np.asarray(v)) for k, v in fix.items() if k.startswith(prefix)}


def oracle_branch(sd, act, coeff, tol=None, n_iterations=None, requires_grad=True, post_act=None):
    """sd: state dict of one reference nn.Sequential branch ('0.weight', '1.beta', ...)."""
    idxs = sorted({int(k.split('.')[0]) for k in sd})
    last = max(idxs)
    layers, pending_act, pending_beta = [], None, None
    have = lambda i, name: ('%d.%s' % (i, name)) in sd
    for i in range(last + 1):
        if have(i, 'weight'):
            W = sd['%d.weight' % i].clone().requires_grad_(requires_grad)
            b = sd['%d.bias' % i].clone().requires_grad_(requires_grad) if have(i, 'bias') else None
            kind = 'linear' if W.dim() == 2 else 'conv'
            spatial = None
            if have(i, 'spatial_dims'):
                spatial = [int(s) for s in sd['%d.spatial_dims' % i].tolist()]
            layers.append(orc.OracleLayer(kind=kind, weight=W, bias=b, u=sd['%d.u' % i].clone(),
                                          v=sd['%d.v' % i].clone(), coeff=coeff, pre_act=pending_act,
                                          beta=pending_beta, padding=(W.shape[-1] // 2 if kind == 'conv' else 0),
                                          n_iterations=n_iterations, atol=tol, rtol=tol, spatial=spatial))
            pending_act, pending_beta = None, None
        elif have(i, 'beta'):
            pending_act, pending_beta = 'swish', sd['%d.beta' % i].clone().requires_grad_(requires_grad)
        else:
            pending_act = act
    br = orc.OracleBranch(layers=layers)
    if post_act is not None:
        br.post_act = post_act
    return br


def branch_param_names(sd):
    """Names (in nn.Module.parameters() order) of the trainable entries of a branch state dict."""
    names = []
    idxs = sorted({int(k.split('.')[0]) for k in sd})
    for i in idxs:
        for n in ('beta', 'weight', 'bias'):
            if '%d.%s' % (i, n) in sd:
                names.append('%d.%s' % (i, n))
    return names


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
