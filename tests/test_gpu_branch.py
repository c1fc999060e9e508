"""GPU checks of the graph-free branch programs with the real kernels (incl. the one-launch tile kernel of
csrc/branch_fused.cu) against the module's autograd path.  Cases: tests/branch_cases.py."""
import pytest

from tests import branch_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', branch_cases.NAMES)
@pytest.mark.parametrize('backend', ['simt', 'auto'])
def test_branch_program_matches_module_autograd(name, backend):
    branch_cases.case_matches_module_autograd(name, backend, 'cuda')


@pytest.mark.parametrize('name', branch_cases.NAMES)
def test_branch_program_gradients_match_autograd(name):
    branch_cases.case_gradients_match_autograd(name, 'auto', 'cuda')


def test_fused3_tile_kernel_is_taken():
    assert branch_cases.case_fused3_is_taken('fused3_lead', 'cuda') == 2
    assert branch_cases.case_fused3_is_taken('fused3_512', 'cuda') == 2
    assert branch_cases.case_fused3_is_taken('cifar_lead', 'cuda') == 0


def test_fused3_matches_unfused_program():
    """Same branch, same inputs: tile kernel vs the three-GEMM program (forward, saved pre-activations, vjp)."""
    import torch
    from impflow_b200 import branch_program
    from impflow_b200.branch_program import compile_branch
    from tests.helpers import rel_err
    net, x = branch_cases._setup('fused3_512', 'cuda')
    v = torch.randn_like(x)
    outs = []
    for on in (True, False):
        branch_program.FUSED3['on'] = on
        try:
            prog = compile_branch(net)
            with torch.no_grad():
                y, saved = prog.forward_saved(x)
                outs.append((y, saved.pres[1], saved.pres[2], prog.vjp(v, saved)))
        finally:
            branch_program.FUSED3['on'] = True
    for a, b in zip(*outs):
        assert rel_err(a.cpu(), b.cpu()) < 4e-6


@pytest.mark.parametrize('c,hw,width,lead', [(3, 8, 512, True), (3, 12, 256, False), (12, 8, 512, True),
                                             (48, 4, 512, True), (5, 6, 128, False)])
@pytest.mark.parametrize('n', [1, 2, 5])
def test_power_series_chain_fused_epilogue(c, hw, width, lead, n):
    """impflow_conv3_power_series with the col2im epilogue of term k also writing the im2col rows (and clearing the
    tap accumulator) of term k + 1 against a k_conv3_in launch per term: Neumann sum and Hutchinson dots bit-identical
    (tile kernel: c = 3, 5; layer-1 GEMM + k_chain23: c = 12, 48), with and without a leading activation."""
    import torch
    import impflow_b200 as pkg
    from impflow_b200.branch_program import compile_branch
    from tests.imblock_cases import build_conv_branch
    torch.manual_seed(c * 100 + hw + n)
    net = build_conv_branch(pkg.layers, c, width, 0.9, 1e-3, lead).cuda()
    x = torch.randn(4, c, hw, hw, device='cuda')
    with torch.no_grad():
        net(x)
    prog = compile_branch(net)
    lib = pkg._cabi.load()
    v = torch.randn_like(x)
    coeffs = [(-1) ** k / (k + 1.0) for k in range(n)]
    outs = []
    for on in (1, 0):
        was = lib.impflow_conv3_set_chain_fuse(on)
        try:
            with torch.no_grad():
                _, saved = prog.forward_saved(x)
                assert prog.native_plan(x) is not None
                w = prog.neumann_chain(saved, v, coeffs)
                dots = prog.hutchinson_series(saved, v, coeffs)
            outs.append((w.clone(), dots.clone()))
        finally:
            lib.impflow_conv3_set_chain_fuse(was)
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    # and against the chain of single vjp evaluations
    with torch.no_grad():
        _, saved = prog.forward_saved(x)
        cur, ref = v, v.clone()
        for k in range(n):
            cur = prog.vjp(cur, saved)
            ref = ref + coeffs[k] * cur
    assert float((outs[0][0] - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


@pytest.mark.parametrize('B,d,hidden,nh,act,n', [(1000, 6, 128, 4, 'sin', 4), (37, 43, 64, 2, 'swish', 3),
                                                 (5, 2, 16, 1, 'relu', 1), (130, 63, 128, 4, 'sin', 7)])
def test_mlp_series_one_launch(B, d, hidden, nh, act, n):
    """k_mlp_series (left / right vectors, their combinations and the estimate of the basic power series in one
    launch) against the host-driven chains of fused vjp / tangent evaluations it replaces."""
    import torch
    import impflow_b200 as pkg
    from impflow_b200.branch_program import compile_branch
    L = pkg.layers
    torch.manual_seed(B + d + n)
    acts = {'sin': L.base.Sin, 'swish': L.base.Swish, 'relu': L.base.ReLU}
    dims = [d] + [hidden] * nh + [d]
    mods = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            mods.append(acts[act]())
        mods.append(L.base.get_linear(a, b, coeff=0.9, n_iterations=None, atol=1e-3, rtol=1e-3, domain=2, codomain=2))
    net = torch.nn.Sequential(*mods)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() > 1:
                p.mul_(3.0)
    net = net.cuda()
    prog = compile_branch(net)
    x = torch.randn(B, d, device='cuda')
    v = torch.randn(B, d, device='cuda')
    coeffs = [(-1) ** k / (k + 1.0) * (1.0 + 0.1 * k) for k in range(n)]
    with torch.no_grad():
        _, saved = prog.forward_saved(x)
        spec = prog.mlp_series_spec(saved)
        assert spec is not None
        S, Ls, Rs, Wm = pkg.ops.mlp_series(spec, v, coeffs)
        ls, rs = [v], [v]
        S_ref = torch.zeros(B, device='cuda')
        for k in range(n):
            ls.append(prog.vjp(ls[-1], saved))
            S_ref += coeffs[k] * (ls[-1] * v).sum(1)
        for _ in range(n - 1):
            rs.append(prog.tangent(saved, rs[-1]))
        close = lambda a, b: float((a - b).abs().max()) <= 2e-5 * max(float(b.abs().max()), 1e-6)
        for k in range(n + 1):
            assert close(Ls[k], ls[k]), k
        for m in range(n):
            assert close(Rs[m], rs[m]), m
            w = sum(coeffs[a + m] * ls[a] for a in range(n - m))
            assert close(Wm[m], w), m
        assert close(S, S_ref)
