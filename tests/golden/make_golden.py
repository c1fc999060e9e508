#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference (`/root/reference/lib`) is imported as-is behind two import shims
(`torch._six`, `termcolor` — SURVEY.md §8c).  Nothing from the reference is copied; only
its inputs/outputs on seeded synthetic data are stored as small .npz files.
"""
import collections.abc
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

_six = types.ModuleType('torch._six')
_six.container_abcs = collections.abc
sys.modules['torch._six'] = _six
_tc = types.ModuleType('termcolor')
_tc.colored = lambda s, *a, **k: s
sys.modules['termcolor'] = _tc
sys.path.insert(0, '/root/reference')

import lib.layers as layers  # noqa: E402
import lib.layers.base as base_layers  # noqa: E402
import lib.layers.broyden as ref_broyden_mod  # noqa: E402
import lib.layers.implicit_block as ref_imblock_mod  # noqa: E402
from lib.implicit_flow import ImplicitFlow  # noqa: E402

torch.set_num_threads(4)


def sd_np(module, prefix=''):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('wrote', path, '%.1f KB' % (os.path.getsize(path) / 1024))


class SolveRecorder:
    """Wraps the reference broyden() to record nstep / lowest_step / trace of each call."""

    def __init__(self):
        self.calls = []
        self._orig = ref_broyden_mod.broyden

    def __enter__(self):
        def wrapped(g, x0, threshold, eps, ls=False, name='unknown'):
            out = self._orig(g, x0, threshold, eps, ls=ls, name=name)
            self.calls.append((name, out['nstep'], out['lowest_step'], list(out['trace'])))
            return out
        ref_imblock_mod.broyden = wrapped
        return self

    def __exit__(self, *a):
        ref_imblock_mod.broyden = self._orig

    def nsteps(self, name):
        return np.array([c[1] for c in self.calls if c[0] == name], dtype=np.int64)

    def traces(self, name):
        tr = [c[3] for c in self.calls if c[0] == name]
        n = max(len(t) for t in tr) if tr else 0
        out = np.full((len(tr), n), np.nan, dtype=np.float64)
        for i, t in enumerate(tr):
            out[i, :len(t)] = t
        return out


def replay_probes(seed, shape_x, shape_z):
    """The two Rademacher draws _logdetgrad makes right after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    B = torch.distributions.bernoulli.Bernoulli(torch.Tensor([0.5]))
    vx = B.sample(shape_x).reshape(shape_x) * 2 - 1
    vz = B.sample(shape_z).reshape(shape_z) * 2 - 1
    return vx.numpy(), vz.numpy()


# ----------------------------------------------------------------------------------------------
# 1. broyden() on analytic maps
# ----------------------------------------------------------------------------------------------

def gen_broyden():
    out = {}
    rng = np.random.RandomState(1)

    def run(tag, B, d, T, eps, scale=0.5, gain=None, zero_row=False):
        W = (rng.randn(d, d) / np.sqrt(d)).astype(np.float32)
        c = rng.randn(B, d).astype(np.float32)
        if zero_row:
            c[0] = 0.
        Wt, ct = torch.from_numpy(W), torch.from_numpy(c)
        if gain is None:
            g = lambda x: ct - scale * torch.tanh(x @ Wt) - x
        else:
            g = lambda x: ct + gain * x
        res = ref_broyden_mod.broyden(g, torch.zeros(B, d), T, eps)
        out[tag + '_W'] = W
        out[tag + '_c'] = c
        out[tag + '_meta'] = np.array([B, d, T, eps, scale, -1. if gain is None else gain], dtype=np.float64)
        out[tag + '_result'] = res['result'].numpy()
        out[tag + '_ints'] = np.array([res['nstep'], res['lowest_step'], int(res['prot_break'])], dtype=np.int64)
        out[tag + '_trace'] = np.array(res['trace'], dtype=np.float64)
        out[tag + '_diff_detail'] = res['diff_detail'].numpy()
        out[tag + '_diff'] = np.array(res['diff'], dtype=np.float64)
        print(tag, 'nstep', res['nstep'], 'lowest', res['lowest_step'], 'prot', res['prot_break'],
              'trace', ['%.3g' % t for t in res['trace']])

    run('small', 7, 5, 30, 1e-6, zero_row=True)
    run('wide', 64, 300, 30, 1e-6, scale=0.9)
    run('capped', 9, 12, 3, 1e-9, scale=0.95)
    run('long', 5, 40, 30, 1e-9, scale=2.5)
    run('protbreak', 4, 6, 30, 1e-6, gain=1e7)
    save('broyden_analytic', **out)


# ----------------------------------------------------------------------------------------------
# 2. imBlock with MLP branches (toy / tabular shapes)
# ----------------------------------------------------------------------------------------------

def build_mlp(dims, coeff, n_iterations, tol, data_dim):
    nnet = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            nnet.append(base_layers.Sin())
        nnet.append(base_layers.get_linear(a, b, coeff=coeff, n_iterations=n_iterations, atol=tol, rtol=tol,
                                           domain=2, codomain=2, zero_init=(b == data_dim)))
    return torch.nn.Sequential(*nnet)


def std_normal_logprob(z):
    return -0.5 * np.log(2 * np.pi) - z.pow(2) / 2


def run_block(tag, blk, x, seed, training, out, with_grad=True, weight_perturb=None):
    if weight_perturb:
        # move the zero-init last layers away from ~0 so that the solves do real work
        with torch.no_grad():
            for n, p in blk.named_parameters():
                if n.endswith('weight') and p.requires_grad:
                    p.mul_(weight_perturb)
        blk.nnet_x_copy.load_state_dict(blk.nnet_x.state_dict())
        blk.nnet_z_copy.load_state_dict(blk.nnet_z.state_dict())
    blk.train(training)
    for k, v in sd_np(blk).items():
        out[tag + '_sd_' + k] = v
    x = x.clone().requires_grad_(with_grad)
    np.random.seed(seed)
    torch.manual_seed(seed)
    with SolveRecorder() as rec:
        z, dlogp = blk(x, torch.zeros(x.shape[0], 1))
        out[tag + '_x'] = x.detach().numpy()
        out[tag + '_z'] = z.detach().numpy()
        out[tag + '_dlogp'] = dlogp.detach().numpy()
        if with_grad:
            logpz = std_normal_logprob(z).view(z.size(0), -1).sum(1, keepdim=True)
            loss = -(logpz - dlogp).mean()
            loss.backward()
            out[tag + '_loss'] = np.array(loss.item())
            out[tag + '_grad_x'] = x.grad.numpy()
            for n, p in blk.named_parameters():
                if p.grad is not None:
                    out[tag + '_grad_' + n] = p.grad.numpy()
    out[tag + '_fwd_nstep'] = rec.nsteps('forward')
    out[tag + '_bwd_nstep'] = rec.nsteps('backward')
    out[tag + '_fwd_trace'] = rec.traces('forward')
    out[tag + '_bwd_trace'] = rec.traces('backward')
    out[tag + '_seed'] = np.array(seed)
    out[tag + '_n_draws'] = blk.last_n_samples.numpy().copy()
    vx, vz = replay_probes(seed, tuple(x.shape), tuple(z.shape))
    out[tag + '_vareps_x'] = vx
    out[tag + '_vareps_z'] = vz
    print(tag, 'fwd', rec.nsteps('forward'), 'bwd', rec.nsteps('backward'),
          'dlogp[:2]', dlogp.detach().flatten()[:2].numpy(), 'n', blk.last_n_samples.numpy())
    # inverse reconstructs x from z
    blk.eval()
    with torch.no_grad(), SolveRecorder() as rec2:
        x_rec = blk.inverse(z.detach())
    out[tag + '_x_rec'] = x_rec.numpy()
    out[tag + '_inv_nstep'] = rec2.nsteps('forward')


def gen_imblock_mlp():
    out = {}
    # toy: d=2, brute force log-det (train_toy.py:224-242, run_toy.sh)
    torch.manual_seed(0)
    np.random.seed(0)
    dims = [2, 32, 32, 2]
    blk = layers.imBlock(build_mlp(dims, 0.99, 20, None, 2), build_mlp(dims, 0.99, 20, None, 2),
                         n_dist='geometric', brute_force=True, n_samples=1, neumann_grad=False,
                         grad_in_forward=False)
    x = torch.rand(50, 2) * 4 - 2
    run_block('toy', blk, x, 11, True, out, weight_perturb=300.)

    # tabular d=6: basic estimator in training (train_tabular.py:314-336, run_tabular.sh)
    torch.manual_seed(1)
    np.random.seed(1)
    dims = [6, 64, 64, 6]
    blk = layers.imBlock(build_mlp(dims, 0.99, None, 1e-3, 6), build_mlp(dims, 0.99, None, 1e-3, 6),
                         n_dist='geometric', n_samples=1, n_exact_terms=2, neumann_grad=False,
                         grad_in_forward=False, eps_forward=1e-5)
    x = torch.randn(40, 6)
    run_block('tab6', blk, x, 12, True, out, weight_perturb=300.)
    # same block in eval mode -> brute force because d <= 10
    out_eval = {}
    run_block('tab6eval', blk, x, 13, False, out_eval, with_grad=False)
    for k in ('tab6eval_z', 'tab6eval_dlogp', 'tab6eval_fwd_nstep'):
        out[k] = out_eval[k]

    # tabular d=43: roulette in train and (20 exact terms) in eval
    torch.manual_seed(2)
    np.random.seed(2)
    dims = [43, 64, 64, 43]
    blk = layers.imBlock(build_mlp(dims, 0.99, None, 1e-3, 43), build_mlp(dims, 0.99, None, 1e-3, 43),
                         n_dist='geometric', n_samples=1, n_exact_terms=2, neumann_grad=False,
                         grad_in_forward=False, eps_forward=1e-5)
    x = torch.randn(24, 43)
    run_block('tab43', blk, x, 14, True, out, weight_perturb=300.)
    out_eval = {}
    run_block('tab43eval', blk, x, 15, False, out_eval, with_grad=False)
    for k in ('tab43eval_z', 'tab43eval_dlogp', 'tab43eval_fwd_nstep', 'tab43eval_n_draws', 'tab43eval_seed',
              'tab43eval_vareps_x', 'tab43eval_vareps_z'):
        out[k] = out_eval[k]
    # eval-mode draw is not stored by the reference in last_n_samples: replay it
    np.random.seed(15)
    out['tab43eval_n_draws'] = np.random.geometric(0.5, 1).astype(np.float32)
    save('imblock_mlp', **out)


# ----------------------------------------------------------------------------------------------
# 3. imBlock with conv branches (CIFAR-style 3-1-3 LipSwish, classifier-style 3-3 ReLU)
# ----------------------------------------------------------------------------------------------

def build_conv_branch(c, idim, coeff, tol, leading_act):
    nnet = []
    if leading_act:
        nnet.append(base_layers.Swish())
    nnet.append(base_layers.get_conv2d(c, idim, 3, 1, 1, coeff=coeff, n_iterations=None, domain=2, codomain=2,
                                       atol=tol, rtol=tol))
    nnet.append(base_layers.Swish())
    nnet.append(base_layers.get_conv2d(idim, idim, 1, 1, 0, coeff=coeff, n_iterations=None, domain=2, codomain=2,
                                       atol=tol, rtol=tol))
    nnet.append(base_layers.Swish())
    nnet.append(base_layers.get_conv2d(idim, c, 3, 1, 1, coeff=coeff, n_iterations=None, domain=2, codomain=2,
                                       atol=tol, rtol=tol))
    return torch.nn.Sequential(*nnet)


def build_cls_branch(c, hidden, coeff, tol):
    mk = lambda a, b: base_layers.get_conv2d(a, b, kernel_size=3, stride=1, padding=1, bias=False, coeff=coeff,
                                             n_iterations=None, domain=2, codomain=2, atol=tol, rtol=tol)
    return torch.nn.Sequential(mk(c, hidden), torch.nn.ReLU(), mk(hidden, c), torch.nn.ReLU())


def gen_imblock_conv():
    out = {}
    torch.manual_seed(3)
    np.random.seed(3)
    c, idim, hw, B = 4, 32, 8, 4
    blk = layers.imBlock(build_conv_branch(c, idim, 0.9, 1e-3, True), build_conv_branch(c, idim, 0.9, 1e-3, True),
                         n_dist='poisson', n_samples=1, n_exact_terms=3, neumann_grad=True, grad_in_forward=True)
    x = torch.randn(B, c, hw, hw)
    with torch.no_grad():
        blk(x, restore=True)      # lazy u/v shaping (train_img.py:502-507)
    run_block('cifar', blk, x, 21, True, out)

    torch.manual_seed(4)
    np.random.seed(4)
    blk = layers.imBlock(build_conv_branch(c, idim, 0.9, 1e-3, False), build_conv_branch(c, idim, 0.9, 1e-3, False),
                         n_dist='poisson', n_samples=1, n_exact_terms=3, neumann_grad=False, grad_in_forward=False)
    with torch.no_grad():
        blk(x, restore=True)
    run_block('cifar_basic', blk, x, 22, True, out)

    # classifier-style block, no log-det (train_classification.py:135-188)
    torch.manual_seed(5)
    np.random.seed(5)
    c, hidden = 8, 16
    blk = layers.imBlock(build_cls_branch(c, hidden, 0.9, 1e-3), build_cls_branch(c, hidden, 0.9, 1e-3))
    x = torch.rand(B, c, hw, hw)
    with torch.no_grad():
        blk(x, restore=True)
    blk.train(True)
    for k, v in sd_np(blk).items():
        out['cls_sd_' + k] = v
    xg = x.clone().requires_grad_(True)
    with SolveRecorder() as rec:
        z = blk(xg)
        loss = (z ** 2).mean()
        loss.backward()
    out['cls_x'] = x.numpy()
    out['cls_z'] = z.detach().numpy()
    out['cls_loss'] = np.array(loss.item())
    out['cls_grad_x'] = xg.grad.numpy()
    for n, p in blk.named_parameters():
        if p.grad is not None:
            out['cls_grad_' + n] = p.grad.numpy()
    out['cls_fwd_nstep'] = rec.nsteps('forward')
    out['cls_bwd_nstep'] = rec.nsteps('backward')
    out['cls_fwd_trace'] = rec.traces('forward')
    out['cls_bwd_trace'] = rec.traces('backward')
    print('cls fwd', rec.nsteps('forward'), 'bwd', rec.nsteps('backward'))
    save('imblock_conv', **out)


# ----------------------------------------------------------------------------------------------
# 4. induced-norm layers
# ----------------------------------------------------------------------------------------------

def gen_induced_norm():
    out = {}
    torch.manual_seed(6)
    lin = base_layers.InducedNormLinear(6, 16, coeff=0.5, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    for k, v in sd_np(lin).items():
        out['lin_init_' + k] = v
    with torch.no_grad():
        lin.weight.add_(0.3 * torch.randn_like(lin.weight))
    out['lin_weight2'] = lin.weight.detach().numpy().copy()
    W = lin.compute_weight(update=True)
    out['lin_W_tol'] = W.detach().numpy()
    out['lin_u_tol'], out['lin_v_tol'] = lin.u.numpy().copy(), lin.v.numpy().copy()
    out['lin_scale_tol'] = lin.scale.numpy().copy()
    W = lin.compute_weight(update=True, n_iterations=5)
    out['lin_W_it5'] = W.detach().numpy()
    out['lin_u_it5'], out['lin_v_it5'] = lin.u.numpy().copy(), lin.v.numpy().copy()
    out['lin_scale_it5'] = lin.scale.numpy().copy()
    xin = torch.randn(5, 6)
    out['lin_x'] = xin.numpy()
    out['lin_y'] = lin(xin).detach().numpy()
    # gradient through sigma
    lin.zero_grad()
    lin(xin).pow(2).sum().backward()
    out['lin_grad_weight'] = lin.weight.grad.numpy().copy()

    torch.manual_seed(7)
    conv = base_layers.InducedNormConv2d(5, 8, 3, 1, 1, coeff=0.4, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    xin = torch.randn(2, 5, 6, 6)
    y0 = conv(xin)                                  # lazy init of u/v at spatial dims 6x6
    for k, v in sd_np(conv).items():
        out['conv_init_' + k] = v
    out['conv_x'] = xin.numpy()
    out['conv_y'] = y0.detach().numpy()
    with torch.no_grad():
        conv.weight.add_(0.2 * torch.randn_like(conv.weight))
    out['conv_weight2'] = conv.weight.detach().numpy().copy()
    W = conv.compute_weight(update=True)
    out['conv_W_tol'] = W.detach().numpy()
    out['conv_u_tol'], out['conv_v_tol'] = conv.u.numpy().copy(), conv.v.numpy().copy()
    out['conv_scale_tol'] = conv.scale.numpy().copy()
    conv.zero_grad()
    conv(xin).pow(2).sum().backward()
    out['conv_grad_weight'] = conv.weight.grad.numpy().copy()

    torch.manual_seed(8)
    c1 = base_layers.InducedNormConv2d(8, 8, 1, 1, 0, coeff=0.3, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    xin = torch.randn(2, 8, 4, 4)
    y0 = c1(xin)
    for k, v in sd_np(c1).items():
        out['c1_init_' + k] = v
    out['c1_x'] = xin.numpy()
    out['c1_y'] = y0.detach().numpy()
    with torch.no_grad():
        c1.weight.add_(0.2 * torch.randn_like(c1.weight))
    out['c1_weight2'] = c1.weight.detach().numpy().copy()
    W = c1.compute_weight(update=True)
    out['c1_W_tol'] = W.detach().numpy()
    out['c1_u_tol'], out['c1_v_tol'] = c1.u.numpy().copy(), c1.v.numpy().copy()
    out['c1_scale_tol'] = c1.scale.numpy().copy()
    save('induced_norm', **out)


def gen_mixed_norm():
    """Induced p -> q norms other than 2 -> 2 and learnable orders (mixed_lipschitz.py:414-444, learn_p):
    constructor with its random restarts, tolerance-mode update after a weight change, one-iteration estimate,
    forward, weight gradient through sigma; Linear, 1x1 (incl. the orders whose buffer is NOT written back) and 3x3."""
    out = {}
    INF = float('inf')
    lin_cases = {'l23': (2, 3), 'l33': (3, 3), 'l13': (1, 3), 'l3i': (3, INF), 'l15_25': (1.5, 2.5)}
    for i, (tag, (dom, cod)) in enumerate(lin_cases.items()):
        torch.manual_seed(40 + i)
        lin = base_layers.InducedNormLinear(6, 7, coeff=0.6, domain=dom, codomain=cod, atol=1e-3, rtol=1e-3)
        for k, v in sd_np(lin).items():
            out[tag + '_init_' + k] = v
        with torch.no_grad():
            lin.weight.add_(0.3 * torch.randn_like(lin.weight))
        out[tag + '_weight2'] = lin.weight.detach().numpy().copy()
        out[tag + '_one_iter'] = lin.compute_one_iter().detach().numpy().copy()
        W = lin.compute_weight(update=True)
        out[tag + '_W_tol'] = W.detach().numpy()
        out[tag + '_u_tol'], out[tag + '_v_tol'] = lin.u.numpy().copy(), lin.v.numpy().copy()
        out[tag + '_scale_tol'] = lin.scale.numpy().copy()
        xin = torch.randn(4, 6)
        out[tag + '_x'] = xin.numpy()
        out[tag + '_y'] = lin(xin).detach().numpy()
        lin.zero_grad()
        lin(xin).pow(2).sum().backward()
        out[tag + '_grad_weight'] = lin.weight.grad.numpy().copy()
        out[tag + '_norms'] = np.array([dom, cod], dtype=np.float64)
    conv_cases = {'c1_3i': (1, (3, INF)), 'c1_13': (1, (1, 3)), 'c1_33': (1, (3, 3)), 'c3_23': (3, (2, 3)),
                  'c3_33': (3, (3, 3)), 'c3_32': (3, (3, 2))}
    for i, (tag, (k, (dom, cod))) in enumerate(conv_cases.items()):
        torch.manual_seed(60 + i)
        conv = base_layers.InducedNormConv2d(3, 4, k, 1, k // 2, coeff=0.5, domain=dom, codomain=cod, atol=1e-3,
                                             rtol=1e-3)
        xin = torch.randn(2, 3, 5, 5)
        y0 = conv(xin)                                   # lazy init with the restarts
        for kk, v in sd_np(conv).items():
            out[tag + '_init_' + kk] = v
        out[tag + '_x'], out[tag + '_y'] = xin.numpy(), y0.detach().numpy()
        with torch.no_grad():
            conv.weight.add_(0.2 * torch.randn_like(conv.weight))
        out[tag + '_weight2'] = conv.weight.detach().numpy().copy()
        out[tag + '_one_iter'] = conv.compute_one_iter().detach().numpy().copy()
        W = conv.compute_weight(update=True)
        out[tag + '_W_tol'] = W.detach().numpy()
        out[tag + '_u_tol'], out[tag + '_v_tol'] = conv.u.numpy().copy(), conv.v.numpy().copy()
        out[tag + '_scale_tol'] = conv.scale.numpy().copy()
        out[tag + '_y2'] = conv(xin).detach().numpy()
        conv.zero_grad()
        conv(xin).pow(2).sum().backward()
        out[tag + '_grad_weight'] = conv.weight.grad.numpy().copy()
        out[tag + '_norms'] = np.array([dom, cod, k], dtype=np.float64)
    # learnable orders: two shared raw parameters, p = asym_squash(raw)
    torch.manual_seed(80)
    pd, pc = torch.nn.Parameter(torch.tensor(0.)), torch.nn.Parameter(torch.tensor(0.4))
    lin = base_layers.InducedNormLinear(5, 6, coeff=0.6, domain=pd, codomain=pc, atol=1e-3, rtol=1e-3)
    for k, v in sd_np(lin).items():
        out['lp_init_' + k] = v
    out['lp_orders'] = np.array([float(o) for o in lin.compute_domain_codomain()])
    with torch.no_grad():
        lin.weight.add_(0.3 * torch.randn_like(lin.weight))
    out['lp_weight2'] = lin.weight.detach().numpy().copy()
    W = lin.compute_weight(update=True)
    out['lp_W_tol'] = W.detach().numpy()
    out['lp_u_tol'], out['lp_v_tol'] = lin.u.numpy().copy(), lin.v.numpy().copy()
    out['lp_scale_tol'] = lin.scale.numpy().copy()
    # state-dict keys of a learn_p flow (the shared order parameters appear under every layer that holds them)
    torch.manual_seed(81)
    flow = ImplicitFlow((2, 3, 8, 8), n_blocks=[1, 1], intermediate_dim=8, factor_out=False, quadratic=False,
                        init_layer=None, actnorm=False, fc_actnorm=False, batchnorm=False, dropout=0., fc=False,
                        coeff=0.9, vnorms='2222', n_lipschitz_iters=None, sn_atol=1e-3, sn_rtol=1e-3,
                        n_power_series=None, n_dist='poisson', n_samples=1, kernels='3-1-3', activation_fn='swish',
                        fc_end=False, fc_idim=16, n_exact_terms=2, preact=True, neumann_grad=True,
                        grad_in_forward=True, first_resblock=True, learn_p=True, classification=False,
                        classification_hdim=64, n_classes=10)
    out['lp_flow_keys'] = np.array(sorted(flow.state_dict().keys()))
    out['lp_flow_nparams'] = np.array([sum(p.numel() for p in flow.parameters())])
    save('mixed_norm', **out)


# ----------------------------------------------------------------------------------------------
# 5. activations
# ----------------------------------------------------------------------------------------------

def gen_activations():
    out = {}
    x = torch.linspace(-6, 6, 241).requires_grad_(True)
    for name, mod in (('sin', base_layers.Sin()), ('swish', base_layers.Swish())):
        y = mod(x)
        (d1,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        (d2,) = torch.autograd.grad(d1.sum(), x, create_graph=True)
        (d3,) = torch.autograd.grad(d2.sum(), x)
        out[name + '_y'], out[name + '_d1'] = y.detach().numpy(), d1.detach().numpy()
        out[name + '_d2'], out[name + '_d3'] = d2.detach().numpy(), d3.detach().numpy()
    sw = base_layers.Swish()
    y = sw(x)
    w = torch.cos(x.detach())
    (gb,) = torch.autograd.grad((y * w).sum(), sw.beta)
    out['swish_grad_beta'] = gb.numpy()
    out['swish_w'] = w.numpy()
    out['x'] = x.detach().numpy()
    save('activations', **out)


# ----------------------------------------------------------------------------------------------
# 6. a small multiscale ImplicitFlow, density-training step (train_img.py:517-549)
# ----------------------------------------------------------------------------------------------

def gen_flow():
    out = {}
    torch.manual_seed(9)
    np.random.seed(9)
    B, c, hw = 4, 3, 8
    model = ImplicitFlow(
        (B, c, hw, hw), n_blocks=[1, 1], intermediate_dim=16, factor_out=False, quadratic=False,
        init_layer=layers.LogitTransform(0.05), actnorm=True, fc_actnorm=False, batchnorm=False, dropout=0.,
        fc=False, coeff=0.9, vnorms='2222', n_lipschitz_iters=None, sn_atol=1e-3, sn_rtol=1e-3,
        n_power_series=None, n_dist='poisson', n_samples=1, kernels='3-1-3', activation_fn='swish', fc_end=False,
        fc_idim=128, n_exact_terms=3, preact=True, neumann_grad=True, grad_in_forward=True, first_resblock=True,
        learn_p=False, classification=False, classification_hdim=64, n_classes=10)
    x = torch.rand(B, c, hw, hw)
    with torch.no_grad():
        model(x, restore=True)
    model.train()
    for k, v in sd_np(model).items():
        out['sd_' + k] = v
    np.random.seed(31)
    torch.manual_seed(31)
    with SolveRecorder() as rec:
        z, dlogp = model(x, 0)
        logpz = std_normal_logprob(z).view(z.size(0), -1).sum(1, keepdim=True)
        ndim = c * hw * hw
        logpx = logpz - dlogp - np.log(256) * ndim
        bpd = -torch.mean(logpx) / ndim / np.log(2)
        bpd.backward()
    out['x'], out['z'], out['dlogp'], out['bpd'] = x.numpy(), z.detach().numpy(), dlogp.detach().numpy(), np.array(bpd.item())
    for n, p in model.named_parameters():
        if p.grad is not None:
            out['grad_' + n] = p.grad.numpy()
    out['fwd_nstep'], out['bwd_nstep'] = rec.nsteps('forward'), rec.nsteps('backward')
    out['n_draws'] = np.stack([m.last_n_samples.numpy() for m in model.modules() if isinstance(m, layers.imBlock)])
    out['seed'] = np.array(31)
    print('flow bpd', bpd.item(), 'fwd', rec.nsteps('forward'), 'bwd', rec.nsteps('backward'))
    model.eval()
    with torch.no_grad():
        out['x_rec'] = model(z.detach(), inverse=True).numpy()
    save('flow_small', **out)


# ----------------------------------------------------------------------------------------------
# 7. iResBlock (lib/layers/iresblock.py): forward / log-det estimators / fixed-point inverse
# ----------------------------------------------------------------------------------------------

def run_ires(tag, blk, x, seed, out, weight_perturb=None):
    """One training step (log-det estimate + gradients), the eval-mode estimate and the fixed-point inverse."""
    if weight_perturb:
        with torch.no_grad():
            for n, p in blk.named_parameters():
                if n.endswith('weight') and p.requires_grad:
                    p.mul_(weight_perturb)
    blk.train()
    for k, v in sd_np(blk).items():
        out[tag + '_sd_' + k] = v
    xg = x.clone().requires_grad_(True)
    np.random.seed(seed)
    torch.manual_seed(seed)
    y, dlogp = blk(xg, torch.zeros(x.shape[0], 1))
    logpy = std_normal_logprob(y).view(y.size(0), -1).sum(1, keepdim=True)
    loss = -(logpy - dlogp).mean()
    loss.backward()
    out[tag + '_x'], out[tag + '_y'], out[tag + '_dlogp'] = x.numpy(), y.detach().numpy(), dlogp.detach().numpy()
    out[tag + '_loss'] = np.array(loss.item())
    out[tag + '_grad_x'] = xg.grad.numpy()
    for n, p in blk.named_parameters():
        if p.grad is not None:
            out[tag + '_grad_' + n] = p.grad.numpy()
    out[tag + '_n_draws'] = blk.last_n_samples.numpy().copy()
    torch.manual_seed(seed)
    out[tag + '_vareps'] = torch.randn_like(x).numpy()          # the Gaussian probe _logdetgrad drew (:129)
    out[tag + '_seed'] = np.array(seed)
    # eval mode: 20 exact terms, basic estimator, no graph
    blk.eval()
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    rate = torch.sigmoid(blk.geom_p).item() if blk.n_dist == 'geometric' else blk.lamb.item()
    ye, dlogpe = blk(x.clone(), torch.zeros(x.shape[0], 1))
    out[tag + 'eval_y'], out[tag + 'eval_dlogp'] = ye.detach().numpy(), dlogpe.detach().numpy()
    np.random.seed(seed + 1)
    draw = np.random.geometric(rate, blk.n_samples) if blk.n_dist == 'geometric' else np.random.poisson(rate, blk.n_samples)
    out[tag + 'eval_n_draws'] = draw.astype(np.float32)
    torch.manual_seed(seed + 1)
    out[tag + 'eval_vareps'] = torch.randn_like(x).numpy()
    with torch.no_grad():
        x_rec = blk.inverse(y.detach())
    out[tag + '_x_rec'] = x_rec.numpy()
    print(tag, 'dlogp[:2]', dlogp.detach().flatten()[:2].numpy(), 'eval', dlogpe.detach().flatten()[:2].numpy(),
          'n', blk.last_n_samples.numpy(), 'rec err', float((x_rec - x).abs().max()))


def gen_ires():
    out = {}
    torch.manual_seed(40)
    np.random.seed(40)
    # train_toy.py:205-223 --arch iresnet: d=2, closed-form 2x2 determinant
    blk = layers.iResBlock(build_mlp([2, 32, 32, 2], 0.9, 20, None, 2), n_dist='geometric', brute_force=True,
                           n_samples=1, neumann_grad=False, grad_in_forward=False)
    run_ires('mlp2', blk, torch.rand(50, 2) * 4 - 2, 41, out, weight_perturb=300.)
    # basic estimator with the double-backward training gradient, geometric roulette
    torch.manual_seed(42)
    blk = layers.iResBlock(build_mlp([6, 64, 64, 6], 0.9, None, 1e-3, 6), n_dist='geometric', n_samples=1,
                           n_exact_terms=2, neumann_grad=False, grad_in_forward=False)
    run_ires('mlp6', blk, torch.randn(40, 6), 43, out, weight_perturb=300.)
    # Neumann gradient estimator + memory-efficient backward (the variant that also returns g, :186-235)
    torch.manual_seed(44)
    blk = layers.iResBlock(build_mlp([6, 64, 64, 6], 0.9, None, 1e-3, 6), n_dist='poisson', n_samples=2,
                           n_exact_terms=3, neumann_grad=True, grad_in_forward=True)
    run_ires('mlp6n', blk, torch.randn(40, 6), 45, out, weight_perturb=300.)
    # conv branch of the image flows (resflow.py:344-392), Neumann + memory-efficient
    torch.manual_seed(46)
    c, idim, hw, B = 4, 32, 8, 4
    blk = layers.iResBlock(build_conv_branch(c, idim, 0.9, 1e-3, True), n_dist='poisson', n_samples=1,
                           n_exact_terms=3, neumann_grad=True, grad_in_forward=True)
    x = torch.randn(B, c, hw, hw)
    with torch.no_grad():
        blk(x)                      # lazy u / v shaping
    run_ires('conv', blk, x, 47, out)
    save('iresblock', **out)


# ----------------------------------------------------------------------------------------------
# 8. imBlock branches the shipped configs do not reach: Banach fallback after a protective break
#    (implicit_block.py:17-28,74-75), exact_trace (:327-343), n_samples > 1, fixed n_power_series, FCNet
# ----------------------------------------------------------------------------------------------

class Affine(torch.nn.Module):
    """y = a * x: a deterministic stand-in branch (no parameters)."""

    def __init__(self, a):
        super(Affine, self).__init__()
        self.a = a

    def forward(self, x):
        return self.a * x


class Cliff(torch.nn.Module):
    """-0.5 z on z > -1 (where the Banach iteration lives), a 1e8-steep wall below: Broyden's first step from
    zeros, z1 = -g(0) = -x_embed (quirk #1), lands behind the wall and trips the 1e6 protective break."""

    def forward(self, z):
        return torch.where(z > -1, -0.5 * z, -0.5 * z + 1e8 * (z + 1))


def gen_edge():
    out = {}
    # --- Banach fallback
    blk = layers.imBlock(Affine(0.1), Cliff())
    x = torch.rand(6, 5) * 2 + 2
    info = {}
    orig = ref_broyden_mod.broyden

    def spy(g, x0, threshold, eps, ls=False, name='unknown'):
        res = orig(g, x0, threshold, eps, ls=ls, name=name)
        info.update(nstep=res['nstep'], prot=res['prot_break'])
        return res
    ref_imblock_mod.broyden = spy
    try:
        with torch.no_grad():
            z = blk(x)
    finally:
        ref_imblock_mod.broyden = orig
    assert info['prot'], 'the case must trip the protective break'
    out['banach_x'], out['banach_z'] = x.numpy(), z.numpy()
    out['banach_ints'] = np.array([info['nstep'], int(info['prot'])], dtype=np.int64)
    print('banach: nstep', info['nstep'], 'prot', info['prot'], 'residual',
          float((z + Cliff()(z) - x - 0.1 * x).abs().max()))

    # --- exact trace (Jacobian powers), training
    torch.manual_seed(50)
    np.random.seed(50)
    blk = layers.imBlock(build_mlp([6, 32, 32, 6], 0.9, None, 1e-3, 6), build_mlp([6, 32, 32, 6], 0.9, None, 1e-3, 6),
                         n_dist='geometric', exact_trace=True, n_samples=1, n_exact_terms=2, neumann_grad=False,
                         grad_in_forward=False, eps_forward=1e-5)
    run_block('exact', blk, torch.randn(16, 6), 51, True, out, weight_perturb=300.)
    # --- three roulette samples per call
    torch.manual_seed(52)
    np.random.seed(52)
    blk = layers.imBlock(build_mlp([12, 32, 32, 12], 0.9, None, 1e-3, 12),
                         build_mlp([12, 32, 32, 12], 0.9, None, 1e-3, 12), n_dist='geometric', n_samples=3,
                         n_exact_terms=2, neumann_grad=False, grad_in_forward=False, eps_forward=1e-5)
    run_block('ns3', blk, torch.randn(16, 12), 53, True, out, weight_perturb=300.)
    # --- truncated series (biased), Neumann + memory-efficient
    torch.manual_seed(54)
    np.random.seed(54)
    blk = layers.imBlock(build_mlp([12, 32, 32, 12], 0.9, None, 1e-3, 12),
                         build_mlp([12, 32, 32, 12], 0.9, None, 1e-3, 12), n_dist='poisson', n_power_series=4,
                         neumann_grad=True, grad_in_forward=True, eps_forward=1e-5)
    run_block('nps', blk, torch.randn(16, 12), 55, True, out, weight_perturb=300.)
    # --- FC tail block (implicit_flow.py:321-356,430-433): FCNet branches on an image-shaped input
    from lib.implicit_flow import FCNet
    torch.manual_seed(56)
    np.random.seed(56)
    shape = (2, 4, 4)
    fc = lambda: FCNet(input_shape=shape, idim=24, lipschitz_layer=base_layers.get_linear, nhidden=2, coeff=0.9,
                       domains=[2., 2., 2.], codomains=[2., 2., 2.], n_iterations=None, activation_fn='swish',
                       preact=True, dropout=0, sn_atol=1e-3, sn_rtol=1e-3, learn_p=False)
    blk = layers.imBlock(fc(), fc(), n_dist='poisson', n_samples=1, n_exact_terms=3, neumann_grad=True,
                         grad_in_forward=True)
    run_block('fc', blk, torch.randn(8, *shape), 57, True, out, weight_perturb=3.)
    save('imblock_edge', **out)


# ----------------------------------------------------------------------------------------------
# 9. step tail: clip_grad_norm_ + lib/optimizers.Adam + utils.ExponentialMovingAverage (train_img.py:652-658)
# ----------------------------------------------------------------------------------------------

def gen_step_tail():
    import lib.optimizers as ref_optim
    import lib.utils as ref_utils
    out = {}
    torch.manual_seed(60)
    shapes = [(8, 3, 3, 3), (8,), (7, 5), (1,), (16, 8, 1, 1)]

    class Holder(torch.nn.Module):
        def __init__(self):
            super(Holder, self).__init__()
            self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s)) for s in shapes])
    m = Holder()
    opt = ref_optim.Adam(m.parameters(), lr=1e-2, betas=(0.9, 0.99), weight_decay=1e-3)   # decay line is a no-op
    ema = ref_utils.ExponentialMovingAverage(m, decay=0.9)
    for i, p in enumerate(m.ps):
        out['p0_%d' % i] = p.detach().numpy().copy()
    n_steps = 4
    for t in range(n_steps):
        opt.zero_grad()
        for i, p in enumerate(m.ps):
            g = torch.randn(*shapes[i]) * (3.0 if t % 2 == 0 else 0.01)          # clipped and unclipped steps
            out['g%d_%d' % (t, i)] = g.numpy().copy()
            p.grad = g.clone()
        total = torch.nn.utils.clip_grad.clip_grad_norm_(m.parameters(), 1.)
        opt.step()
        ema.apply()
        out['gnorm%d' % t] = np.array(float(total))
        for i, p in enumerate(m.ps):
            out['p%d_%d' % (t + 1, i)] = p.detach().numpy().copy()
            out['ema%d_%d' % (t + 1, i)] = ema.shadow_params['ps.%d' % i].numpy().copy()
    out['meta'] = np.array([n_steps, len(shapes), 1e-2, 0.9, 0.99, 1e-8, 1.0, 0.9], dtype=np.float64)
    save('step_tail', **out)


# ----------------------------------------------------------------------------------------------
# 10. how reproducible are the reference's own implicit-backward iteration counts?  The backward solves run
#     with eps_backward = 1e-10 * sqrt(B d), below fp32 round-off, so the count depends on the summation order
#     of the BLAS kernels.  Re-run the stored fixtures' training step with 1 / 2 / 4 / 8 intra-op threads.
# ----------------------------------------------------------------------------------------------

def gen_bwd_band():
    out = {}
    fxm, fxc, fxe = (dict(np.load(os.path.join(HERE, n + '.npz'))) for n in ('imblock_mlp', 'imblock_conv',
                                                                              'imblock_edge'))
    cases = []
    mk_mlp = lambda dims, n_it, tol, **kw: layers.imBlock(build_mlp(dims, 0.99, n_it, tol, dims[0]),
                                                          build_mlp(dims, 0.99, n_it, tol, dims[0]), **kw)
    cases.append(('toy', fxm, lambda: mk_mlp([2, 32, 32, 2], 20, None, n_dist='geometric', brute_force=True,
                                             n_samples=1, neumann_grad=False, grad_in_forward=False)))
    for tag, d in (('tab6', 6), ('tab43', 43)):
        cases.append((tag, fxm, lambda d=d: mk_mlp([d, 64, 64, d], None, 1e-3, n_dist='geometric', n_samples=1,
                                                   n_exact_terms=2, neumann_grad=False, grad_in_forward=False,
                                                   eps_forward=1e-5)))
    cases.append(('cifar', fxc, lambda: layers.imBlock(
        build_conv_branch(4, 32, 0.9, 1e-3, True), build_conv_branch(4, 32, 0.9, 1e-3, True), n_dist='poisson',
        n_samples=1, n_exact_terms=3, neumann_grad=True, grad_in_forward=True)))
    cases.append(('cifar_basic', fxc, lambda: layers.imBlock(
        build_conv_branch(4, 32, 0.9, 1e-3, False), build_conv_branch(4, 32, 0.9, 1e-3, False), n_dist='poisson',
        n_samples=1, n_exact_terms=3, neumann_grad=False, grad_in_forward=False)))
    old_threads = torch.get_num_threads()
    # variants: intra-op thread counts, and mathematically equivalent re-orderings of the batch (per-sample
    # arithmetic unchanged; only the order of the batch-global sums and the GEMM row blocking move)
    for tag, fx, make in cases:
        B = fx[tag + '_x'].shape[0]
        rs = np.random.RandomState(7)
        variants = [('threads1', 1, np.arange(B)), ('threads2', 2, np.arange(B)), ('threads8', 8, np.arange(B)),
                    ('reversed', 4, np.arange(B)[::-1].copy()), ('perm_a', 4, rs.permutation(B)),
                    ('perm_b', 4, rs.permutation(B))]
        if tag in ('toy', 'tab6', 'tab43'):
            # the same function with the hidden units of both branches re-numbered: every per-sample sum runs in a
            # different order (what any re-implementation - another BLAS, FMA contraction, a GPU - also does)
            variants += [('hidden_perm_' + ch, 4, np.arange(B)) for ch in 'abcdefghijkl']
        counts, last = [], []
        for name, nt, perm in variants:
            torch.set_num_threads(nt)
            blk = make()
            x = torch.from_numpy(fx[tag + '_x'])
            with torch.no_grad():
                blk(x, restore=True)
            sd = {k[len(tag + '_sd_'):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith(tag + '_sd_')}
            if name.startswith('hidden_perm'):
                hrs = np.random.RandomState(ord(name[-1]))
                for net in ('nnet_x', 'nnet_z', 'nnet_x_copy', 'nnet_z_copy'):
                    h = sd[net + '.0.weight'].shape[0]
                    hrs2 = np.random.RandomState(hrs.randint(1 << 30) if net in ('nnet_x', 'nnet_z') else 0)
                    if net.endswith('_copy'):      # the frozen twins follow their live nets
                        p1, p2 = perms[net[:-5]]
                    else:
                        p1, p2 = torch.from_numpy(hrs2.permutation(h)), torch.from_numpy(hrs2.permutation(h))
                        perms = dict(locals().get('perms', {}), **{net: (p1, p2)})
                    sd[net + '.0.weight'] = sd[net + '.0.weight'][p1]
                    sd[net + '.0.bias'] = sd[net + '.0.bias'][p1]
                    sd[net + '.0.u'] = sd[net + '.0.u'][p1]
                    sd[net + '.2.weight'] = sd[net + '.2.weight'][p2][:, p1]
                    sd[net + '.2.bias'] = sd[net + '.2.bias'][p2]
                    sd[net + '.2.u'] = sd[net + '.2.u'][p2]
                    sd[net + '.2.v'] = sd[net + '.2.v'][p1]
                    sd[net + '.4.weight'] = sd[net + '.4.weight'][:, p2]
                    sd[net + '.4.v'] = sd[net + '.4.v'][p2]
            blk.load_state_dict(sd, strict=True)
            blk.train()
            seed = int(fx[tag + '_seed'])
            np.random.seed(seed)            # same roulette draw
            vx, vz = torch.from_numpy(fx[tag + '_vareps_x'])[perm], torch.from_numpy(fx[tag + '_vareps_z'])[perm]
            probes = iter([vx, vz])

            class _Replay(object):          # the block's two Bernoulli draws, re-ordered with the batch
                def __init__(self, *a, **k):
                    pass

                def sample(self, shape):
                    return (next(probes) + 1) / 2
            orig_b = torch.distributions.bernoulli.Bernoulli
            torch.distributions.bernoulli.Bernoulli = _Replay
            try:
                xg = x[perm].clone().requires_grad_(True)
                with SolveRecorder() as rec:
                    z, dlogp = blk(xg, torch.zeros(x.shape[0], 1))
                    loss = -(std_normal_logprob(z).view(z.size(0), -1).sum(1, keepdim=True) - dlogp).mean()
                    loss.backward()
            finally:
                torch.distributions.bernoulli.Bernoulli = orig_b
            counts.append([int(rec.nsteps('forward')[0]), int(rec.nsteps('backward')[0])])
            tr = rec.traces('backward')[0]
            last.append(float(tr[~np.isnan(tr)][-1]))
        out[tag + '_variants'] = np.array([v[0] for v in variants])
        out[tag + '_fwd_bwd_nstep'] = np.array(counts, dtype=np.int64)
        out[tag + '_bwd_last_residual'] = np.array(last)
        print(tag, 'fwd/bwd nstep by variant', counts, 'last backward residual', ['%.3g' % v for v in last])
    torch.set_num_threads(old_threads)
    save('bwd_band', **out)


# ----------------------------------------------------------------------------------------------
# 12. closed-form operator-norm layers (lipschitz.py:274-366) and a block under vnorms '122f'
# ----------------------------------------------------------------------------------------------
def gen_lop():
    """LopLinear / LopConv2d for the five (domain, codomain) pairs with a closed-form operator norm, local and global
    constraint: effective weight, scale buffer, forward, gradients w.r.t. the weight and the input; then a training
    step of an imBlock whose conv branches are built by the factories under vnorms '122f' (first layer 1 -> 2 and
    last layer 2 -> inf are Lop layers, the middle one an induced 2 -> 2 layer)."""
    out = {}
    INF = float('inf')
    pairs = {'11': (1, 1), '12': (1, 2), '1i': (1, INF), '2i': (2, INF), 'ii': (INF, INF)}
    i = 0
    for ptag, (dom, cod) in pairs.items():
        for local in (True, False):
            for kind in ('lin', 'c1', 'c3'):
                tag = 'lop_%s_%s_%s' % (kind, ptag, 'loc' if local else 'glob')
                torch.manual_seed(300 + i)
                i += 1
                if kind == 'lin':
                    m = base_layers.get_linear(6, 7, coeff=0.3, domain=dom, codomain=cod, local_constraint=local)
                    xin = torch.randn(5, 6)
                else:
                    k = 1 if kind == 'c1' else 3
                    m = base_layers.get_conv2d(3, 4, k, 1, k // 2, coeff=0.3, domain=dom, codomain=cod,
                                               local_constraint=local)
                    xin = torch.randn(2, 3, 5, 5)
                assert type(m).__name__ in ('LopLinear', 'LopConv2d')
                xin.requires_grad_(True)
                y = m(xin)
                y.pow(2).sum().backward()
                for k_, v in sd_np(m).items():
                    out[tag + '_sd_' + k_] = v
                out[tag + '_x'], out[tag + '_y'] = xin.detach().numpy(), y.detach().numpy()
                out[tag + '_W'] = m.compute_weight().detach().numpy()
                out[tag + '_grad_weight'] = m.weight.grad.numpy().copy()
                out[tag + '_grad_bias'] = m.bias.grad.numpy().copy()
                out[tag + '_grad_x'] = xin.grad.numpy().copy()
                out[tag + '_norms'] = np.array([dom, cod, 1.0 if local else 0.0])
    out['lop_tags'] = np.array(sorted({k.rsplit('_sd_', 1)[0] for k in out if '_sd_' in k}))

    def branch(c, idim, coeff, tol):
        doms, cods = [1, 2, 2], [2, 2, INF]           # vnorms '122f' (implicit_flow.py:361-363)
        mk = lambda a, b, k, j: base_layers.get_conv2d(a, b, k, 1, k // 2, coeff=coeff, n_iterations=None,
                                                       domain=doms[j], codomain=cods[j], atol=tol, rtol=tol)
        return torch.nn.Sequential(base_layers.Swish(), mk(c, idim, 3, 0), base_layers.Swish(), mk(idim, idim, 1, 1),
                                   base_layers.Swish(), mk(idim, c, 3, 2))
    torch.manual_seed(13)
    np.random.seed(13)
    c, idim, hw, B = 4, 16, 6, 3
    blk = layers.imBlock(branch(c, idim, 0.9, 1e-3), branch(c, idim, 0.9, 1e-3), n_dist='poisson', n_samples=1,
                         n_exact_terms=3, neumann_grad=True, grad_in_forward=True)
    x = torch.randn(B, c, hw, hw)
    with torch.no_grad():
        blk(x, restore=True)
    run_block('lopblk', blk, x, 23, True, out, weight_perturb=3.0)

    # the stack builders' optional layers (implicit_flow.py:374-396, 463): MovingBatchNorm2d after every conv, Dropout
    # in front of the last layer, under vnorms '122f' with the FC tail.  One training-mode forward moves the running
    # means; the fixture is the state dict after it and the eval-mode latent (dropout off, running means frozen).
    torch.manual_seed(14)
    np.random.seed(14)
    flow = ImplicitFlow((2, 3, 8, 8), n_blocks=[1, 1], intermediate_dim=8, factor_out=False, quadratic=False,
                        init_layer=layers.LogitTransform(0.05), actnorm=True, fc_actnorm=False, batchnorm=True,
                        dropout=0.2, fc=False, coeff=0.9, vnorms='122f', n_lipschitz_iters=None, sn_atol=1e-3,
                        sn_rtol=1e-3, n_power_series=None, n_dist='poisson', n_samples=1, kernels='3-1-3',
                        activation_fn='swish', fc_end=True, fc_idim=16, n_exact_terms=2, preact=True,
                        neumann_grad=True, grad_in_forward=True, first_resblock=True, learn_p=False,
                        classification=False, classification_hdim=64, n_classes=10)
    x = torch.rand(2, 3, 8, 8)
    with torch.no_grad():
        flow(x, restore=True)
    flow.train()
    with torch.no_grad():
        for n_, p_ in flow.named_parameters():
            if n_.endswith('weight') and p_.dim() > 1:
                p_.mul_(2.0)
            if n_.endswith('bias') and 'nnet' in n_:
                p_.add_(0.1 * torch.randn_like(p_))
    flow(x, 0)[0].sum().backward()        # one training evaluation: running means move, gradients exist
    flow.eval()
    with torch.no_grad():
        z = flow(x)
    for k, v in sd_np(flow).items():
        out['fopt_sd_' + k] = v
    out['fopt_x'], out['fopt_z'] = x.numpy(), z.numpy()
    out['fopt_modules'] = np.array([type(m).__name__ for m in flow.modules()])
    bn = [v for k, v in sd_np(flow).items() if k.endswith('running_mean') and '_copy' not in k]
    print('fopt: |running_mean| max', max(float(np.abs(v).max()) for v in bn), 'z[:4]', z.flatten()[:4].numpy())
    save('lop', **out)


if __name__ == '__main__':
    which = sys.argv[1:] or ['broyden', 'mlp', 'conv', 'norm', 'mixed', 'act', 'flow', 'ires', 'edge', 'tail', 'band', 'lop']
    if 'broyden' in which: gen_broyden()
    if 'mlp' in which: gen_imblock_mlp()
    if 'conv' in which: gen_imblock_conv()
    if 'norm' in which: gen_induced_norm()
    if 'mixed' in which: gen_mixed_norm()
    if 'act' in which: gen_activations()
    if 'flow' in which: gen_flow()
    if 'ires' in which: gen_ires()
    if 'edge' in which: gen_edge()
    if 'tail' in which: gen_step_tail()
    if 'band' in which: gen_bwd_band()
    if 'lop' in which: gen_lop()
