"""Pins the CPU oracle (oracle/impflow_oracle.py) against fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import impflow_oracle as orc
from tests.helpers import oracle_branch, sub_sd, rel_err, branch_param_names


@pytest.mark.parametrize('tag', ['small', 'wide', 'capped', 'long', 'protbreak'])
def test_broyden_analytic(golden, tag):
    fx = golden('broyden_analytic')
    B, d, T, eps, scale, gain = fx[tag + '_meta']
    W, c = torch.from_numpy(fx[tag + '_W']), torch.from_numpy(fx[tag + '_c'])
    if gain < 0:
        g = lambda x: c - float(scale) * torch.tanh(x @ W) - x
    else:
        g = lambda x: c + float(gain) * x
    res = orc.broyden_solve(g, torch.zeros(int(B), int(d)), int(T), float(eps))
    nstep, lowest_step, prot = fx[tag + '_ints']
    assert res['nstep'] == nstep
    assert res['lowest_step'] == lowest_step
    assert int(res['prot_break']) == prot
    np.testing.assert_array_equal(res['result'].numpy(), fx[tag + '_result'])   # same ATen ops -> bit equal
    np.testing.assert_allclose(np.array(res['trace']), fx[tag + '_trace'], rtol=0, atol=0)
    np.testing.assert_array_equal(res['diff_detail'].numpy(), fx[tag + '_diff_detail'])


def _std_normal_logprob(z):
    return -0.5 * np.log(2 * np.pi) - z.pow(2) / 2


def _run_oracle_block(fx, tag, act, coeff, cfg, training, tol=1e-3, n_iterations=None, with_grad=True):
    sd_x, sd_z = sub_sd(fx, tag + '_sd_nnet_x.'), sub_sd(fx, tag + '_sd_nnet_z.')
    bx = oracle_branch(sd_x, act, coeff, tol, n_iterations)
    bz = oracle_branch(sd_z, act, coeff, tol, n_iterations)
    x = torch.from_numpy(fx[tag + '_x']).clone().requires_grad_(with_grad)
    stats = {}
    n_draws = fx.get(tag + '_n_draws')
    probes = (torch.from_numpy(fx[tag + '_vareps_x']), torch.from_numpy(fx[tag + '_vareps_z']))
    z, dlogp = orc.imblock_forward(bx, bz, x, torch.zeros(x.shape[0], 1), cfg, training,
                                   n_draws=None if n_draws is None else n_draws.astype(np.int64), probes=probes,
                                   stats=stats)
    grads = {}
    if with_grad:
        logpz = _std_normal_logprob(z).view(z.size(0), -1).sum(1, keepdim=True)
        loss = -(logpz - dlogp).mean()
        loss.backward()
        grads['x'] = x.grad
        for name, p in zip(branch_param_names(sd_x), bx.parameters()):
            grads['nnet_x.' + name] = p.grad
        for name, p in zip(branch_param_names(sd_z), bz.parameters()):
            grads['nnet_z.' + name] = p.grad
    return z, dlogp, stats, grads, (bx, bz)


MLP_CASES = {
    'toy': dict(act='sin', coeff=0.99, n_iterations=20, tol=None,
                cfg=dict(orc.DEFAULT_CFG, brute_force=True, neumann_grad=False, grad_in_forward=False)),
    'tab6': dict(act='sin', coeff=0.99, n_iterations=None, tol=1e-3,
                 cfg=dict(orc.DEFAULT_CFG, neumann_grad=False, grad_in_forward=False, eps_forward=1e-5)),
    'tab43': dict(act='sin', coeff=0.99, n_iterations=None, tol=1e-3,
                  cfg=dict(orc.DEFAULT_CFG, neumann_grad=False, grad_in_forward=False, eps_forward=1e-5)),
}


@pytest.mark.parametrize('tag', list(MLP_CASES))
def test_imblock_mlp_train(golden, tag):
    fx = golden('imblock_mlp')
    case = MLP_CASES[tag]
    z, dlogp, stats, grads, _ = _run_oracle_block(fx, tag, case['act'], case['coeff'], case['cfg'], True,
                                                  case['tol'], case['n_iterations'])
    assert stats['fwd_nstep'] == fx[tag + '_fwd_nstep'].tolist()
    assert stats['bwd_nstep'] == fx[tag + '_bwd_nstep'].tolist()
    assert rel_err(z.detach(), fx[tag + '_z']) < 1e-6
    assert rel_err(dlogp.detach(), fx[tag + '_dlogp']) < 1e-5
    assert rel_err(grads['x'], fx[tag + '_grad_x']) < 1e-4
    for k, g in grads.items():
        if k == 'x' or g is None:
            continue
        ref = fx[tag + '_grad_' + k]
        assert rel_err(g, ref) < 2e-4, k


@pytest.mark.parametrize('tag', ['tab6', 'tab43'])
def test_imblock_mlp_eval(golden, tag):
    fx = golden('imblock_mlp')
    case = MLP_CASES[tag]
    fx2 = dict(fx)
    for k in ('n_draws', 'vareps_x', 'vareps_z'):
        if tag + 'eval_' + k in fx:
            fx2[tag + '_' + k] = fx[tag + 'eval_' + k]
    z, dlogp, stats, _, _ = _run_oracle_block(fx2, tag, case['act'], case['coeff'], case['cfg'], False,
                                              case['tol'], case['n_iterations'], with_grad=False)
    assert stats['fwd_nstep'] == fx[tag + 'eval_fwd_nstep'].tolist()
    assert rel_err(z.detach(), fx[tag + 'eval_z']) < 1e-6
    assert rel_err(dlogp.detach(), fx[tag + 'eval_dlogp']) < 1e-5


def test_imblock_mlp_inverse(golden):
    fx = golden('imblock_mlp')
    case = MLP_CASES['tab6']
    sd_x, sd_z = sub_sd(fx, 'tab6_sd_nnet_x.'), sub_sd(fx, 'tab6_sd_nnet_z.')
    bx = oracle_branch(sd_x, 'sin', 0.99, 1e-3, requires_grad=False)
    bz = oracle_branch(sd_z, 'sin', 0.99, 1e-3, requires_grad=False)
    stats = {}
    x_rec = orc.imblock_inverse(bx, bz, torch.from_numpy(fx['tab6_z']), case['cfg'], stats)
    assert stats['inv_nstep'] == fx['tab6_inv_nstep'].tolist()
    assert rel_err(x_rec, fx['tab6_x_rec']) < 1e-6
    assert rel_err(x_rec, fx['tab6_x']) < 1e-4


CONV_CASES = {
    'cifar': dict(cfg=dict(orc.DEFAULT_CFG, n_dist='poisson', n_exact_terms=3, neumann_grad=True,
                           grad_in_forward=True)),
    'cifar_basic': dict(cfg=dict(orc.DEFAULT_CFG, n_dist='poisson', n_exact_terms=3, neumann_grad=False,
                                 grad_in_forward=False)),
}


@pytest.mark.parametrize('tag', list(CONV_CASES))
def test_imblock_conv_train(golden, tag):
    fx = golden('imblock_conv')
    z, dlogp, stats, grads, _ = _run_oracle_block(fx, tag, 'swish', 0.9, CONV_CASES[tag]['cfg'], True)
    assert stats['fwd_nstep'] == fx[tag + '_fwd_nstep'].tolist()
    assert stats['bwd_nstep'] == fx[tag + '_bwd_nstep'].tolist()
    assert rel_err(z.detach(), fx[tag + '_z']) < 1e-6
    assert rel_err(dlogp.detach(), fx[tag + '_dlogp']) < 1e-5
    assert rel_err(grads['x'], fx[tag + '_grad_x']) < 1e-4
    for k, g in grads.items():
        if k == 'x' or g is None:
            continue
        assert rel_err(g, fx[tag + '_grad_' + k]) < 5e-4, k


def test_imblock_classifier_block(golden):
    fx = golden('imblock_conv')
    sd_x, sd_z = sub_sd(fx, 'cls_sd_nnet_x.'), sub_sd(fx, 'cls_sd_nnet_z.')
    bx = oracle_branch(sd_x, 'relu', 0.9, 1e-3, post_act='relu')
    bz = oracle_branch(sd_z, 'relu', 0.9, 1e-3, post_act='relu')
    x = torch.from_numpy(fx['cls_x']).clone().requires_grad_(True)
    stats = {}
    z = orc.imblock_forward(bx, bz, x, None, dict(orc.DEFAULT_CFG), True, stats=stats)
    (z ** 2).mean().backward()
    assert stats['fwd_nstep'] == fx['cls_fwd_nstep'].tolist()
    assert stats['bwd_nstep'] == fx['cls_bwd_nstep'].tolist()
    assert rel_err(z.detach(), fx['cls_z']) < 1e-6
    assert rel_err(x.grad, fx['cls_grad_x']) < 1e-4


def test_induced_norm_linear(golden):
    fx = golden('induced_norm')
    W = torch.from_numpy(fx['lin_weight2'])
    u0, v0 = torch.from_numpy(fx['lin_init_u']), torch.from_numpy(fx['lin_init_v'])
    u, v, _ = orc.power_iterate_matrix(W, u0, v0, None, 1e-3, 1e-3)
    np.testing.assert_allclose(u.numpy(), fx['lin_u_tol'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(v.numpy(), fx['lin_v_tol'], rtol=1e-6, atol=1e-7)
    sigma = orc.sigma_matrix(W, u, v)
    np.testing.assert_allclose(sigma.numpy(), fx['lin_scale_tol'], rtol=1e-6)
    np.testing.assert_allclose(orc.soft_rescale(W, sigma, 0.5).numpy(), fx['lin_W_tol'], rtol=1e-6, atol=1e-7)
    u, v, used = orc.power_iterate_matrix(W, u, v, 5, 1e-3, 1e-3)
    assert used == 5
    np.testing.assert_allclose(u.numpy(), fx['lin_u_it5'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(orc.sigma_matrix(W, u, v).numpy(), fx['lin_scale_it5'], rtol=1e-6)


def test_induced_norm_conv(golden):
    fx = golden('induced_norm')
    W = torch.from_numpy(fx['conv_weight2'])
    u0, v0 = torch.from_numpy(fx['conv_init_u']), torch.from_numpy(fx['conv_init_v'])
    u, v, _ = orc.power_iterate_conv(W, u0, v0, (5, 6, 6), 1, 1, None, 1e-3, 1e-3)
    np.testing.assert_allclose(u.numpy(), fx['conv_u_tol'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(v.numpy(), fx['conv_v_tol'], rtol=1e-5, atol=1e-7)
    sigma = orc.sigma_conv(W, u, v, (5, 6, 6), 1, 1)
    np.testing.assert_allclose(sigma.numpy(), fx['conv_scale_tol'], rtol=1e-6)
    np.testing.assert_allclose(orc.soft_rescale(W, sigma, 0.4).numpy(), fx['conv_W_tol'], rtol=1e-6, atol=1e-7)
    # 1x1
    W1 = torch.from_numpy(fx['c1_weight2'])
    u0, v0 = torch.from_numpy(fx['c1_init_u']), torch.from_numpy(fx['c1_init_v'])
    u, v, _ = orc.power_iterate_matrix(W1.view(8, 8), u0, v0, None, 1e-3, 1e-3)
    np.testing.assert_allclose(u.numpy(), fx['c1_u_tol'], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(orc.sigma_matrix(W1.view(8, 8), u, v).numpy(), fx['c1_scale_tol'], rtol=1e-6)


def test_activations(golden):
    fx = golden('activations')
    x = torch.from_numpy(fx['x']).requires_grad_(True)
    beta = torch.tensor([0.5], requires_grad=True)
    for name, fn in (('sin', orc.sin_act), ('swish', lambda t: orc.lipswish(t, beta))):
        y = fn(x)
        (d1,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        (d2,) = torch.autograd.grad(d1.sum(), x, create_graph=True)
        np.testing.assert_allclose(y.detach().numpy(), fx[name + '_y'], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(d1.detach().numpy(), fx[name + '_d1'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(d2.detach().numpy(), fx[name + '_d2'], rtol=1e-5, atol=1e-6)


def test_roulette_coefficients():
    # geometric: P(N >= k) = (1-p)^(k-1); with draw n=2 and 2 exact terms: n_ps = 4
    n_ps, coeffs = orc.roulette_coefficients([2], 2, 'geometric', 0.5)
    assert n_ps == 4
    np.testing.assert_allclose(coeffs, [1., 1., 1., 2.])
    n_ps, coeffs = orc.roulette_coefficients([0], 3, 'poisson', lamb=2.0)
    assert n_ps == 3 and coeffs == [1., 1., 1.]
    n_ps, coeffs = orc.roulette_coefficients([2], 3, 'poisson', lamb=2.0)
    assert n_ps == 5
    np.testing.assert_allclose(coeffs[3], 1 / (1 - np.exp(-2.0)))
    np.testing.assert_allclose(coeffs[4], 1 / (1 - np.exp(-2.0) * (1 + 2.0)))


# ---- round 2: branches of the block the shipped configs do not reach, iResBlock, the step tail -------------------

class _Cliff(object):
    """nnet_z of the Banach fixture (tests/golden/make_golden.py: Cliff)."""

    def __call__(self, z):
        return torch.where(z > -1, -0.5 * z, -0.5 * z + 1e8 * (z + 1))


def test_banach_fallback_after_protective_break(golden):
    fx = golden('imblock_edge')
    x = torch.from_numpy(fx['banach_x'])
    _, info = orc.root_find(_Cliff(), lambda t: 0.1 * t, x.clone(), x, 1e-6, 30)
    assert [info['nstep'], int(info['prot_break'])] == fx['banach_ints'].tolist()
    with torch.no_grad():       # solve + re-attach, as imBlock.forward does (implicit_block.py:224-227)
        z = orc.imblock_forward(lambda t: 0.1 * t, _Cliff(), x, None, dict(orc.DEFAULT_CFG), True)
    np.testing.assert_array_equal(z.numpy(), fx['banach_z'])


EDGE_CASES = {
    'exact': dict(act='sin', cfg=dict(orc.DEFAULT_CFG, exact_trace=True, neumann_grad=False, grad_in_forward=False,
                                      eps_forward=1e-5)),
    'ns3': dict(act='sin', cfg=dict(orc.DEFAULT_CFG, n_samples=3, neumann_grad=False, grad_in_forward=False,
                                    eps_forward=1e-5)),
    'nps': dict(act='sin', cfg=dict(orc.DEFAULT_CFG, n_dist='poisson', n_power_series=4, neumann_grad=True,
                                    grad_in_forward=True, eps_forward=1e-5)),
}


@pytest.mark.parametrize('tag', list(EDGE_CASES))
def test_imblock_edge_train(golden, tag):
    fx = golden('imblock_edge')
    case = EDGE_CASES[tag]
    z, dlogp, stats, grads, _ = _run_oracle_block(fx, tag, case['act'], 0.9, case['cfg'], True, 1e-3, None)
    assert stats['fwd_nstep'] == fx[tag + '_fwd_nstep'].tolist()
    assert stats['bwd_nstep'] == fx[tag + '_bwd_nstep'].tolist()
    assert rel_err(z.detach(), fx[tag + '_z']) < 1e-6
    assert rel_err(dlogp.detach(), fx[tag + '_dlogp']) < 1e-5
    assert rel_err(grads['x'], fx[tag + '_grad_x']) < 1e-4
    for k, g in grads.items():
        if k != 'x' and g is not None:
            assert rel_err(g, fx[tag + '_grad_' + k]) < 5e-4, k


class _FC(object):
    """FCNet wrapper of an oracle branch (implicit_flow.py:437-474): flatten, MLP, reshape."""

    def __init__(self, branch, shape):
        self.branch, self.shape = branch, shape

    def __call__(self, x):
        return self.branch(x.reshape(x.shape[0], -1)).view(x.shape[0], *self.shape)

    def parameters(self):
        return self.branch.parameters()


def test_imblock_fc_tail_block(golden):
    fx = golden('imblock_edge')
    sd_x, sd_z = sub_sd(fx, 'fc_sd_nnet_x.nnet.'), sub_sd(fx, 'fc_sd_nnet_z.nnet.')
    bx = _FC(oracle_branch(sd_x, 'swish', 0.9, 1e-3), (2, 4, 4))
    bz = _FC(oracle_branch(sd_z, 'swish', 0.9, 1e-3), (2, 4, 4))
    cfg = dict(orc.DEFAULT_CFG, n_dist='poisson', n_exact_terms=3, neumann_grad=True, grad_in_forward=True)
    x = torch.from_numpy(fx['fc_x']).clone().requires_grad_(True)
    stats = {}
    z, dlogp = orc.imblock_forward(bx, bz, x, torch.zeros(x.shape[0], 1), cfg, True,
                                   n_draws=fx['fc_n_draws'].astype(np.int64),
                                   probes=(torch.from_numpy(fx['fc_vareps_x']), torch.from_numpy(fx['fc_vareps_z'])),
                                   stats=stats)
    loss = -(_std_normal_logprob(z).view(z.size(0), -1).sum(1, keepdim=True) - dlogp).mean()
    loss.backward()
    assert stats['fwd_nstep'] == fx['fc_fwd_nstep'].tolist()
    assert rel_err(z.detach(), fx['fc_z']) < 1e-6
    assert rel_err(dlogp.detach(), fx['fc_dlogp']) < 1e-5
    assert rel_err(x.grad, fx['fc_grad_x']) < 1e-4


IRES_CASES = {
    'mlp2': dict(act='sin', n_it=20, tol=None, cfg=dict(orc.DEFAULT_CFG, brute_force=True, neumann_grad=False,
                                                        grad_in_forward=False)),
    'mlp6': dict(act='sin', n_it=None, tol=1e-3, cfg=dict(orc.DEFAULT_CFG, neumann_grad=False, grad_in_forward=False)),
    'mlp6n': dict(act='sin', n_it=None, tol=1e-3, cfg=dict(orc.DEFAULT_CFG, n_dist='poisson', n_samples=2,
                                                           n_exact_terms=3, neumann_grad=True, grad_in_forward=True)),
    'conv': dict(act='swish', n_it=None, tol=1e-3, cfg=dict(orc.DEFAULT_CFG, n_dist='poisson', n_exact_terms=3,
                                                            neumann_grad=True, grad_in_forward=True)),
}


@pytest.mark.parametrize('tag', list(IRES_CASES))
def test_iresblock(golden, tag):
    fx = golden('iresblock')
    case = IRES_CASES[tag]
    sd = sub_sd(fx, tag + '_sd_nnet.')
    net = oracle_branch(sd, case['act'], 0.9, case['tol'], case['n_it'])
    x = torch.from_numpy(fx[tag + '_x']).clone().requires_grad_(True)
    y, dlogp = orc.ires_forward(net, x, torch.zeros(x.shape[0], 1), case['cfg'], True,
                                n_draws=fx[tag + '_n_draws'].astype(np.int64), probe=torch.from_numpy(fx[tag + '_vareps']))
    loss = -(_std_normal_logprob(y).view(y.size(0), -1).sum(1, keepdim=True) - dlogp).mean()
    loss.backward()
    assert rel_err(y.detach(), fx[tag + '_y']) < 1e-6
    assert rel_err(dlogp.detach(), fx[tag + '_dlogp']) < 1e-5
    np.testing.assert_allclose(loss.item(), fx[tag + '_loss'], rtol=1e-6)
    assert rel_err(x.grad, fx[tag + '_grad_x']) < 1e-4
    for name, p in zip(branch_param_names(sd), net.parameters()):
        if tag + '_grad_nnet.' + name in fx and p.grad is not None:
            assert rel_err(p.grad, fx[tag + '_grad_nnet.' + name]) < 5e-4, name
    # eval mode (20 exact terms / closed form) and the fixed-point inverse
    ye, dle = orc.ires_forward(net, torch.from_numpy(fx[tag + '_x']).clone(), torch.zeros(x.shape[0], 1), case['cfg'],
                               False, n_draws=fx[tag + 'eval_n_draws'].astype(np.int64),
                               probe=torch.from_numpy(fx[tag + 'eval_vareps']))
    assert rel_err(ye.detach(), fx[tag + 'eval_y']) < 1e-6
    assert rel_err(dle.detach(), fx[tag + 'eval_dlogp']) < 1e-5
    with torch.no_grad():
        x_rec, _ = orc.ires_inverse(net, torch.from_numpy(fx[tag + '_y']))
    assert rel_err(x_rec, fx[tag + '_x_rec']) < 1e-6
    assert rel_err(x_rec, fx[tag + '_x']) < 1e-3


def test_step_tail_matches_reference_adam_and_ema(golden):
    """clip_grad_norm_ + lib/optimizers.Adam + utils.ExponentialMovingAverage over four steps (fixture produced by
    the reference's own classes): parameters and EMA shadow after every step, incl. the copy-only first apply()."""
    fx = golden('step_tail')
    n_steps, n_p, lr, b1, b2, eps, max_norm, decay = fx['meta']
    n_steps, n_p = int(n_steps), int(n_p)
    ps = [torch.from_numpy(fx['p0_%d' % i]).clone() for i in range(n_p)]
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    ema = [torch.zeros_like(p) for p in ps]
    for t in range(n_steps):
        gs = [torch.from_numpy(fx['g%d_%d' % (t, i)]).clone() for i in range(n_p)]
        orc.clip_adam_ema_step(ps, gs, ms, vs, t + 1, lr, (b1, b2), eps, max_norm, ema, decay)
        for i in range(n_p):
            np.testing.assert_allclose(ps[i].numpy(), fx['p%d_%d' % (t + 1, i)], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(ema[i].numpy(), fx['ema%d_%d' % (t + 1, i)], rtol=2e-6, atol=1e-7)
