"""CPU tests of the product's HOST logic (autograd wiring of the kernel primitives, conv-as-GEMM
weight layouts, Broyden host loop, RNG / roulette handling, state-dict compatibility) with the
C ABI replaced by the numpy emulator in tests/cabi_emulator.py.  The same cases run against the
real kernels in test_gpu_imblock.py."""
import numpy as np
import pytest
import torch

from tests import cabi_emulator
from tests import imblock_cases as cases
from tests.helpers import rel_err


@pytest.fixture(autouse=True)
def _emulated(monkeypatch):
    cases.DEV['device'] = 'cpu'
    lib = cabi_emulator.install(monkeypatch)
    yield lib
    cases.DEV['device'] = 'cuda'


@pytest.mark.parametrize('tag', list(cases.MLP))
def test_imblock_mlp_train(golden, tag):
    cases.case_imblock_mlp_train(golden, tag)


@pytest.mark.parametrize('tag', ['tab6'])
def test_imblock_mlp_train_reference_rng(golden, tag):
    cases.case_imblock_mlp_train_reference_rng(golden, tag)


@pytest.mark.parametrize('tag', ['tab6', 'tab43'])
def test_imblock_mlp_eval_and_inverse(golden, tag):
    cases.case_imblock_mlp_eval_and_inverse(golden, tag)


@pytest.mark.parametrize('tag', list(cases.CONV))
@pytest.mark.parametrize('backend', ['simt', 'tc'])
def test_imblock_conv_train(golden, tag, backend):
    cases.case_imblock_conv_train(golden, tag, backend)


def test_imblock_classifier_block(golden):
    cases.case_imblock_classifier_block(golden)


def test_imblock_conv_train_unfused(golden):
    """Same parity through the module / autograd path (graph-free branch programs switched off)."""
    from impflow_b200.layers import implicit_block
    implicit_block.FUSED['on'] = False
    try:
        cases.case_imblock_conv_train(golden, 'cifar', 'simt')
        cases.case_imblock_classifier_block(golden)
    finally:
        implicit_block.FUSED['on'] = True


def test_implicit_flow_density_step(golden):
    cases.case_implicit_flow_density_step(golden)


@pytest.mark.parametrize('tag', ['small', 'wide', 'capped', 'protbreak'])
def test_broyden_host_loop(golden, tag):
    import impflow_b200
    fx = golden('broyden_analytic')
    B, d, T, eps, scale, gain = fx[tag + '_meta']
    W, c = torch.from_numpy(fx[tag + '_W']), torch.from_numpy(fx[tag + '_c'])
    g = (lambda x: c - float(scale) * torch.tanh(x @ W) - x) if gain < 0 else (lambda x: c + float(gain) * x)
    res = impflow_b200.layers.broyden.broyden(g, torch.zeros(int(B), int(d)), int(T), float(eps))
    nstep, lowest_step, prot = fx[tag + '_ints']
    assert (res['nstep'], res['lowest_step'], int(res['prot_break'])) == (nstep, lowest_step, prot)
    assert rel_err(res['result'], fx[tag + '_result']) < 1e-5
    assert len(res['trace']) == nstep + 1


def test_state_dict_keys_match_reference(golden):
    """Key-for-key state-dict compatibility with the reference modules (SURVEY.md §5 checkpoint row)."""
    import impflow_b200
    layers = impflow_b200.layers
    fx = golden('imblock_mlp')
    blk = cases.make_mlp_block('tab6')
    ref_keys = sorted(k[len('tab6_sd_'):] for k in fx if k.startswith('tab6_sd_'))
    assert sorted(blk.state_dict().keys()) == ref_keys
    fx = golden('imblock_conv')
    blk = layers.imBlock(cases.build_conv_branch(layers, 4, 32, 0.9, 1e-3, True),
                         cases.build_conv_branch(layers, 4, 32, 0.9, 1e-3, True))
    ref_keys = sorted(k[len('cifar_sd_'):] for k in fx if k.startswith('cifar_sd_'))
    assert sorted(blk.state_dict().keys()) == ref_keys
    assert 'geom_p' not in blk.state_dict() and 'lamb' in blk.state_dict()       # quirk #12


def test_compat_install_registers_reference_import_names():
    import impflow_b200
    impflow_b200.compat.install()
    import lib.layers as L
    import lib.layers.base as BL
    from lib.implicit_flow import ImplicitFlow
    from lib.layers.broyden import broyden
    assert L.imBlock is impflow_b200.layers.imBlock and BL.get_conv2d is impflow_b200.layers.base.get_conv2d
    assert ImplicitFlow is impflow_b200.ImplicitFlow and callable(broyden)
    import sys
    for k in [k for k in sys.modules if k == 'lib' or k.startswith('lib.')]:
        del sys.modules[k]


def test_induced_norm_layers_vs_golden(golden):
    import impflow_b200
    BL = impflow_b200.layers.base
    fx = golden('induced_norm')
    lin = BL.InducedNormLinear(6, 16, coeff=0.5, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    lin.load_state_dict({k[len('lin_init_'):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith('lin_init_')})
    with torch.no_grad():
        lin.weight.copy_(torch.from_numpy(fx['lin_weight2']))
    W = lin.compute_weight(update=True)
    np.testing.assert_allclose(W.detach().numpy(), fx['lin_W_tol'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lin.scale.numpy(), fx['lin_scale_tol'], rtol=1e-5)
    W = lin.compute_weight(update=True, n_iterations=5)
    np.testing.assert_allclose(W.detach().numpy(), fx['lin_W_it5'], rtol=1e-5, atol=1e-6)
    y = lin(torch.from_numpy(fx['lin_x']))
    np.testing.assert_allclose(y.detach().numpy(), fx['lin_y'], rtol=1e-4, atol=1e-5)
    lin.zero_grad()
    lin(torch.from_numpy(fx['lin_x'])).pow(2).sum().backward()
    np.testing.assert_allclose(lin.weight.grad.numpy(), fx['lin_grad_weight'], rtol=1e-3, atol=1e-5)

    conv = BL.InducedNormConv2d(5, 8, 3, 1, 1, coeff=0.4, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    x = torch.from_numpy(fx['conv_x'])
    with torch.no_grad():
        conv(x)
    conv.load_state_dict({k[len('conv_init_'):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith('conv_init_')})
    np.testing.assert_allclose(conv(x).detach().numpy(), fx['conv_y'], rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        conv.weight.copy_(torch.from_numpy(fx['conv_weight2']))
    W = conv.compute_weight(update=True)
    np.testing.assert_allclose(conv.u.numpy(), fx['conv_u_tol'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(W.detach().numpy(), fx['conv_W_tol'], rtol=1e-5, atol=1e-6)
    conv.zero_grad()
    conv(x).pow(2).sum().backward()
    np.testing.assert_allclose(conv.weight.grad.numpy(), fx['conv_grad_weight'], rtol=2e-3, atol=1e-4)

    c1 = BL.InducedNormConv2d(8, 8, 1, 1, 0, coeff=0.3, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    x = torch.from_numpy(fx['c1_x'])
    with torch.no_grad():
        c1(x)
    c1.load_state_dict({k[len('c1_init_'):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith('c1_init_')})
    np.testing.assert_allclose(c1(x).detach().numpy(), fx['c1_y'], rtol=1e-4, atol=1e-5)
    with torch.no_grad():
        c1.weight.copy_(torch.from_numpy(fx['c1_weight2']))
    W = c1.compute_weight(update=True)
    np.testing.assert_allclose(W.detach().numpy(), fx['c1_W_tol'], rtol=1e-5, atol=1e-6)


def _branch_cases():
    import impflow_b200
    L = impflow_b200.layers
    torch.manual_seed(0)
    mk = lambda a, b, k, bias=True: L.base.get_conv2d(a, b, k, 1, k // 2, bias=bias, coeff=0.9, n_iterations=None,
                                                      domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    lin = lambda a, b: L.base.get_linear(a, b, coeff=0.9, n_iterations=None, atol=1e-3, rtol=1e-3, domain=2, codomain=2)
    return {
        'cifar_lead': (torch.nn.Sequential(L.base.Swish(), mk(4, 32, 3), L.base.Swish(), mk(32, 32, 1), L.base.Swish(),
                                           mk(32, 4, 3)), (2, 4, 8, 8)),
        'cifar_nolead': (torch.nn.Sequential(mk(3, 32, 3), L.base.Swish(), mk(32, 32, 1), L.base.Swish(), mk(32, 3, 3)),
                         (2, 3, 8, 8)),
        'cls_relu': (torch.nn.Sequential(mk(8, 16, 3, False), torch.nn.ReLU(), mk(16, 8, 3, False), torch.nn.ReLU()),
                     (2, 8, 8, 8)),
        'mlp_sin': (torch.nn.Sequential(lin(6, 64), L.base.Sin(), lin(64, 64), L.base.Sin(), lin(64, 6)), (9, 6)),
        'wide_both': (torch.nn.Sequential(mk(32, 64, 3), L.base.Swish(), mk(64, 32, 3)), (1, 32, 4, 4)),
    }


@pytest.mark.parametrize('name', ['cifar_lead', 'cifar_nolead', 'cls_relu', 'mlp_sin', 'wide_both'])
@pytest.mark.parametrize('backend', ['simt', 'tc'])
def test_branch_program_matches_module_autograd(name, backend):
    """The fused graph-free forward / vjp equals the module's autograd forward / vjp."""
    import impflow_b200
    from impflow_b200.branch_program import compile_branch
    impflow_b200.ops.set_gemm_backend(backend)
    try:
        net, shape = _branch_cases()[name]
        x = torch.randn(*shape)
        with torch.no_grad():
            net(x)                                  # lazy u/v shaping
            for p in net.parameters():
                if p.dim() > 1:
                    p.mul_(3.0)                     # make the spectral rescale active
        prog = compile_branch(net)
        assert prog is not None
        xr = x.clone().requires_grad_(True)
        y_ref = net(xr)
        v = torch.randn_like(y_ref)
        (vjp_ref,) = torch.autograd.grad(y_ref, xr, v)
        with torch.no_grad():
            y = prog.forward(x)
            y2 = prog.forward(x, save=True)
            vjp = prog.vjp(v)
            vjp_again = prog.vjp(v)
        assert rel_err(y, y_ref.detach()) < 2e-6
        assert rel_err(y2, y_ref.detach()) < 2e-6
        assert rel_err(vjp, vjp_ref) < 5e-6
        assert rel_err(vjp_again, vjp_ref) < 5e-6
    finally:
        impflow_b200.ops.set_gemm_backend('auto')


def test_branch_program_rejects_unknown_modules():
    from impflow_b200.branch_program import compile_branch
    assert compile_branch(torch.nn.Sequential(torch.nn.Linear(3, 3))) is None
    assert compile_branch(torch.nn.Tanh()) is None


@pytest.mark.parametrize('name', ['cifar_lead', 'cifar_nolead', 'cls_relu', 'mlp_sin', 'wide_both'])
@pytest.mark.parametrize('backend', ['simt', 'tc'])
def test_branch_program_gradients_match_autograd(name, backend):
    """backward_full (first order) and neumann (hand-derived double backward) against autograd through
    the differentiable kernel primitives."""
    import impflow_b200
    from impflow_b200.branch_program import compile_branch
    impflow_b200.ops.set_gemm_backend(backend)
    try:
        net, shape = _branch_cases()[name]
        x = torch.randn(*shape)
        with torch.no_grad():
            net(x)
            for p in net.parameters():
                if p.dim() > 1:
                    p.mul_(3.0)
        prog = compile_branch(net)
        params = list(net.parameters())
        # ---- first-order backward
        xr = x.clone().requires_grad_(True)
        y = net(xr)
        gout = torch.randn_like(y)
        ref = torch.autograd.grad(y, [xr] + params, gout, allow_unused=True)
        with torch.no_grad():
            _, saved = prog.forward_saved(x)
            gx, pg = prog.backward_full(saved, gout)
        assert rel_err(gx, ref[0]) < 1e-5
        for p, g, r in zip(params, pg, ref[1:]):
            assert (g is None) == (r is None)
            if r is not None:
                assert rel_err(g, r) < 2e-5, tuple(p.shape)
        # ---- Neumann estimator: S = <w^T J, v>, dS/dx, dS/dtheta
        w, v = torch.randn_like(y), torch.randn_like(x)
        xr = x.clone().requires_grad_(True)
        y = net(xr)
        (wJ,) = torch.autograd.grad(y, xr, w, create_graph=True)
        S_ref = (wJ.reshape(x.shape[0], -1) * v.reshape(x.shape[0], -1)).sum(1)
        ref = torch.autograd.grad(S_ref.sum(), [xr] + params, allow_unused=True)
        with torch.no_grad():
            _, saved = prog.forward_saved(x)
            S, gx, pg = prog.neumann(saved, w, v)
        assert rel_err(S, S_ref.detach()) < 1e-5
        if ref[0] is not None and float(ref[0].norm()) > 0:
            assert rel_err(gx, ref[0]) < 2e-5
        else:
            assert float(gx.norm()) < 1e-6
        for p, g, r in zip(params, pg, ref[1:]):
            if r is None or float(r.norm()) == 0:
                assert g is None or float(g.norm()) < 1e-6
            else:
                assert rel_err(g, r) < 5e-5, tuple(p.shape)
    finally:
        impflow_b200.ops.set_gemm_backend('auto')


def test_fused_paths_are_taken(golden, monkeypatch):
    """Guards against silently falling back to the module/autograd path: one CIFAR-style training
    forward+backward must run the Neumann gradient, the attach backward and all vjps on the
    graph-free branch programs and never build an autograd graph through the branch."""
    import impflow_b200
    from impflow_b200 import branch_program as bp
    from impflow_b200.layers import implicit_block as ib
    counts = {'neumann': 0, 'backward_full': 0, 'vjp': 0, 'autograd_estimator': 0, 'module_calls': 0}

    def wrap(cls, name, key):
        orig = getattr(cls, name)

        def f(self, *a, **k):
            counts[key] += 1
            return orig(self, *a, **k)
        monkeypatch.setattr(cls, name, f)
    wrap(bp.BranchProgram, 'neumann', 'neumann')
    wrap(bp.BranchProgram, 'backward_full', 'backward_full')
    wrap(bp.BranchProgram, 'vjp', 'vjp')
    orig_est = ib.neumann_logdet_estimator
    monkeypatch.setattr(ib, 'neumann_logdet_estimator',
                        lambda *a, **k: (counts.__setitem__('autograd_estimator', counts['autograd_estimator'] + 1),
                                         orig_est(*a, **k))[1])
    fx = golden('imblock_conv')
    layers = impflow_b200.layers
    blk = layers.imBlock(cases.build_conv_branch(layers, 4, 32, 0.9, 1e-3, True),
                         cases.build_conv_branch(layers, 4, 32, 0.9, 1e-3, True), **cases.CONV['cifar']['kw'])
    blk = cases.load_block(blk, fx, 'cifar', torch.from_numpy(fx['cifar_x']))
    hook = lambda m, i: counts.__setitem__('module_calls', counts['module_calls'] + 1)
    hs = [blk.nnet_x.register_forward_pre_hook(hook), blk.nnet_z.register_forward_pre_hook(hook)]
    cases.run_train(blk, fx, 'cifar')
    for h in hs:
        h.remove()
    assert counts['neumann'] == 2 and counts['backward_full'] == 2
    assert counts['autograd_estimator'] == 0 and counts['module_calls'] == 0
    assert counts['vjp'] > 2 * 3
