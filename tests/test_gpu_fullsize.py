"""Size-independent properties at BASELINE.json's FULL sizes (real kernels, through the C ABI): the oracle cannot run
these shapes in seconds, so parity at full size is the round trip data -> latent -> data through every block's forward
and inverse Broyden solve, the fixed-point residual of every block and run-to-run determinism.  The same case runs at a
reduced size on the C-ABI emulator in tests/test_host_logic.py."""
import pytest

from tests import imblock_cases as cases

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _on_gpu():
    cases.DEV['device'] = 'cuda'
    yield


# workload, round-trip bound (relative L2; measured 5e-7 ... 2e-6, profiles/r02_fullsize_properties.txt),
# residual bound relative to the block input (measured 4e-8 ... 6e-8; classifier 2e-6)
@pytest.mark.parametrize('workload,rt_tol,res_tol', [('cifar', 1e-4, 1e-5), ('tabular-power', 1e-4, 1e-5),
                                                     ('tabular-bsds300', 1e-4, 1e-5), ('toy', 1e-4, 1e-5)])
def test_full_size_round_trip(workload, rt_tol, res_tol):
    r = cases.workload_round_trip(workload)
    print(workload, r)
    assert r['round_trip'] < rt_tol, r
    assert r['residual'] < res_tol, r
    assert r['deterministic'], r
    assert len(r['fwd_nstep']) == r['blocks'] and len(r['inv_nstep']) == r['blocks'], r


def test_full_size_classifier_blocks():
    """ImplicitResNet18 at B = 128 (d up to 65 536 per sample): every imBlock's output solves its fixed-point equation."""
    r = cases.workload_round_trip('classifier')
    print('classifier', r)
    assert r['residual'] < 5e-5 and r['deterministic'] and len(r['fwd_nstep']) == r['blocks'] == 4, r
