"""Device-parameterised checks of the graph-free branch programs (branch_program.py) against the module's
own autograd path.  Run on CPU under the C-ABI emulator (tests/test_host_logic.py) and on the GPU with
the real kernels (tests/test_gpu_branch.py)."""
import torch

from tests.helpers import rel_err

NAMES = ['cifar_lead', 'cifar_nolead', 'cls_relu', 'mlp_sin', 'wide_both', 'fused3_lead', 'fused3_nolead',
         'fused3_512']


def branch_cases():
    import impflow_b200
    L = impflow_b200.layers
    torch.manual_seed(0)
    mk = lambda a, b, k, bias=True: L.base.get_conv2d(a, b, k, 1, k // 2, bias=bias, coeff=0.9, n_iterations=None,
                                                      domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    lin = lambda a, b: L.base.get_linear(a, b, coeff=0.9, n_iterations=None, atol=1e-3, rtol=1e-3, domain=2, codomain=2)
    sw = L.base.Swish
    return {
        'cifar_lead': (torch.nn.Sequential(sw(), mk(4, 32, 3), sw(), mk(32, 32, 1), sw(), mk(32, 4, 3)), (2, 4, 8, 8)),
        'cifar_nolead': (torch.nn.Sequential(mk(3, 32, 3), sw(), mk(32, 32, 1), sw(), mk(32, 3, 3)), (2, 3, 8, 8)),
        'cls_relu': (torch.nn.Sequential(mk(8, 16, 3, False), torch.nn.ReLU(), mk(16, 8, 3, False), torch.nn.ReLU()),
                     (2, 8, 8, 8)),
        'mlp_sin': (torch.nn.Sequential(lin(6, 64), L.base.Sin(), lin(64, 64), L.base.Sin(), lin(64, 6)), (9, 6)),
        'wide_both': (torch.nn.Sequential(mk(32, 64, 3), sw(), mk(64, 32, 3)), (1, 32, 4, 4)),
        # shapes the one-launch tile kernel (csrc/branch_fused.cu) takes: 9c <= 32, width % 256 == 0
        'fused3_lead': (torch.nn.Sequential(sw(), mk(3, 256, 3), sw(), mk(256, 256, 1), sw(), mk(256, 3, 3)),
                        (2, 3, 8, 8)),
        'fused3_nolead': (torch.nn.Sequential(mk(3, 256, 3), sw(), mk(256, 256, 1), sw(), mk(256, 3, 3)),
                          (3, 3, 8, 8)),
        'fused3_512': (torch.nn.Sequential(sw(), mk(2, 512, 3), sw(), mk(512, 512, 1, False), sw(), mk(512, 2, 3)),
                       (5, 2, 8, 8)),
    }


def _setup(name, device):
    net, shape = branch_cases()[name]
    x = torch.randn(*shape)
    net = net.to(device)
    x = x.to(device)
    with torch.no_grad():
        net(x)                                  # lazy u/v shaping
        for p in net.parameters():
            if p.dim() > 1:
                p.mul_(3.0)                     # make the spectral rescale active
    return net, x


def case_matches_module_autograd(name, backend, device='cpu'):
    """The graph-free forward / vjp equals the module's autograd forward / vjp."""
    import impflow_b200
    from impflow_b200.branch_program import compile_branch
    impflow_b200.ops.set_gemm_backend(backend)
    try:
        net, x = _setup(name, device)
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
        # three chained 3xTF32 GEMMs against the module path (exact fp32 on the CUDA cores at these sizes)
        assert rel_err(y.cpu(), y_ref.detach().cpu()) < 1e-5
        assert rel_err(y2.cpu(), y_ref.detach().cpu()) < 1e-5
        assert rel_err(vjp.cpu(), vjp_ref.cpu()) < 1e-5
        assert rel_err(vjp_again.cpu(), vjp_ref.cpu()) < 1e-5
        return prog
    finally:
        impflow_b200.ops.set_gemm_backend('auto')


def case_gradients_match_autograd(name, backend, device='cpu'):
    """backward_full (first order) and neumann (hand-derived double backward) against autograd through
    the differentiable kernel primitives."""
    import impflow_b200
    from impflow_b200.branch_program import compile_branch
    impflow_b200.ops.set_gemm_backend(backend)
    try:
        net, x = _setup(name, device)
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
        assert rel_err(gx.cpu(), ref[0].cpu()) < 1e-5
        for p, g, r in zip(params, pg, ref[1:]):
            assert (g is None) == (r is None)
            if r is not None:
                assert rel_err(g.cpu(), r.cpu()) < 2e-5, tuple(p.shape)
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
        assert rel_err(S.cpu(), S_ref.detach().cpu()) < 1e-5
        if ref[0] is not None and float(ref[0].norm()) > 0:
            assert rel_err(gx.cpu(), ref[0].cpu()) < 2e-5
        else:
            assert float(gx.norm()) < 1e-6
        for p, g, r in zip(params, pg, ref[1:]):
            if r is None or float(r.norm()) == 0:
                assert g is None or float(g.norm()) < 1e-6
            else:
                assert rel_err(g.cpu(), r.cpu()) < 5e-5, tuple(p.shape)
    finally:
        impflow_b200.ops.set_gemm_backend('auto')


def case_fused3_is_taken(name, device='cpu'):
    """The qualifying conv branches really run on the one-launch tile kernel (and not otherwise)."""
    from impflow_b200 import ops
    from impflow_b200.branch_program import compile_branch
    net, x = _setup(name, device)
    prog = compile_branch(net)
    ops.GEMM_PROFILE['on'], ops.GEMM_PROFILE['shapes'] = True, {}
    try:
        with torch.no_grad():
            prog.forward(x, save=True)
            prog.vjp(torch.randn_like(x))
        kinds = [k[0] for k in ops.GEMM_PROFILE['shapes'] if isinstance(k[0], str)]
    finally:
        ops.GEMM_PROFILE['on'], ops.GEMM_PROFILE['shapes'] = False, {}
    return kinds.count('branch3')
