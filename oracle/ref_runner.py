"""Loader / harness for the UNMODIFIED reference in oracle/_ref — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`load()` imports the reference's `lib` package from oracle/_ref (bytecode compiled by oracle/build_ref.py) behind the two
import shims it needs on a current stack and returns a namespace with the same attribute names bench.py uses for
this repo's package (`layers`, `ImplicitFlow`, `optim`, `utils`), so one model builder / one training-step
function drives both arms.  The reference's modules are taken out of `sys.modules` again afterwards, so they
never collide with `impflow_b200.compat.install()`'s `lib` overlay."""
import collections.abc
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
ARCHIVE = os.path.join(REF_DIR, 'reference_lib.bin')      # zip container of sourceless bytecode (build_ref.py)


def available():
    return os.path.isfile(ARCHIVE)


_ns = None


def load():
    global _ns
    if _ns is not None:
        return _ns
    if not available():
        raise ImportError('oracle/_ref is not built: run python oracle/build_ref.py in the build container')
    saved = {k: v for k, v in sys.modules.items() if k == 'lib' or k.startswith('lib.')}
    for k in saved:
        del sys.modules[k]
    had_six, had_tc = sys.modules.get('torch._six'), sys.modules.get('termcolor')
    six = types.ModuleType('torch._six')
    six.container_abcs = collections.abc                      # lib/layers/base/utils.py:1, mixed_lipschitz.py:1
    sys.modules['torch._six'] = six
    if had_tc is None:
        tc = types.ModuleType('termcolor')
        tc.colored = lambda s, *a, **k: s                     # lib/layers/broyden.py:11
        sys.modules['termcolor'] = tc
    sys.path.insert(0, ARCHIVE)
    try:
        import lib.layers as layers
        import lib.layers.base as base_layers
        import lib.layers.broyden as broyden_mod
        import lib.layers.implicit_block as imblock_mod
        import lib.optimizers as optim
        import lib.utils as utils
        from lib.implicit_flow import ImplicitFlow
    finally:
        sys.path.remove(ARCHIVE)
        for k in [k for k in sys.modules if k == 'lib' or k.startswith('lib.')]:
            del sys.modules[k]
        sys.modules.update(saved)
        if had_six is None:
            sys.modules.pop('torch._six', None)
    ns = types.SimpleNamespace(layers=layers, base_layers=base_layers, ImplicitFlow=ImplicitFlow, optim=optim,
                               utils=utils, broyden_mod=broyden_mod, imblock_mod=imblock_mod, kind='reference')
    _ns = ns
    return ns


class SolveCounter(object):
    """Records (name, nstep) of every reference broyden() call made inside the `with` block."""

    def __init__(self, ns):
        self.ns, self.calls = ns, []

    def __enter__(self):
        self._orig = self.ns.broyden_mod.broyden

        def wrapped(g, x0, threshold, eps, ls=False, name='unknown'):
            out = self._orig(g, x0, threshold, eps, ls=ls, name=name)
            self.calls.append((name, int(out['nstep'])))
            return out
        self.ns.imblock_mod.broyden = wrapped
        return self

    def __exit__(self, *a):
        self.ns.imblock_mod.broyden = self._orig

    def nsteps(self, name):
        return [n for k, n in self.calls if k == name]


def update_lipschitz(ns, model, n_iterations=None):
    """train_img.py:786-792 / train_toy.py:174-179: compute_weight(update=True) on every induced-norm module
    (the frozen *_copy nets included, as the reference does)."""
    import torch
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (ns.base_layers.InducedNormConv2d, ns.base_layers.InducedNormLinear)):
                if n_iterations is None:
                    m.compute_weight(update=True)
                else:
                    m.compute_weight(update=True, n_iterations=n_iterations)
