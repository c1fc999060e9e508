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
