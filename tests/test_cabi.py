"""CPU-side checks of the C-ABI library: it loads and exports every symbol the header declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, 'include', 'impflow_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(impflow_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    import impflow_b200
    lib = ctypes.CDLL(impflow_b200._cabi.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), 'missing export ' + n
    # every declared function has a typed binding and vice versa
    assert set(impflow_b200._cabi.SIGNATURES) == set(names)
    assert impflow_b200._cabi.load().impflow_version() == 1


def test_state_struct_layout_matches_host_dtype():
    import impflow_b200
    from impflow_b200.layers import broyden as b
    assert int(impflow_b200._cabi.load().impflow_broyden_state_bytes()) == b._STATE_DTYPE.itemsize


def test_product_refuses_cpu_tensors():
    import pytest
    import torch
    import impflow_b200
    with pytest.raises(RuntimeError, match='CUDA'):
        impflow_b200.layers.broyden.broyden(lambda x: x, torch.zeros(2, 3), 5, 1e-3)
    with pytest.raises(RuntimeError, match='CUDA'):
        impflow_b200.ops.act_mul(torch.zeros(4), None, 1, 0)
    lin = impflow_b200.layers.base.get_linear(4, 3, coeff=0.9, domain=2, codomain=2, atol=1e-3, rtol=1e-3)
    with pytest.raises(RuntimeError, match='CUDA'):
        lin(torch.zeros(2, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'implicit-normalizing-flows_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('no oracle', ''), os.path.join(dirpath, f)
