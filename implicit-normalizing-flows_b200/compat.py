"""Drop-in registration under the reference's import names.

    import impflow_b200; impflow_b200.compat.install()        # before the script's own `import lib....`
    import lib.layers as layers; import lib.layers.base as base_layers
    from lib.implicit_flow import ImplicitFlow

after which train_toy.py / train_tabular.py / train_img.py / train_classification.py of the reference resolve
their MODEL code (`lib.layers*`, `lib.implicit_flow`, `lib.resflow`) to this package, while everything else they
import from `lib` (`lib.utils`, `lib.optimizers`, `lib.toy_data`, `lib.tabular`, `lib.datasets`,
`lib.lr_scheduler`, `lib.visualize_flow` — logging, data and plotting, none of it on the hot path) keeps coming
from the reference checkout the scripts live in.  tests/test_train_scripts.py runs the
reference's train_toy.py and train_tabular.py this way."""
import os
import sys
import types

_HOT_PATH = ('lib.layers', 'lib.implicit_flow', 'lib.resflow')


def _find_reference_lib():
    """The `lib/` directory of the reference checkout the running script belongs to: first sys.path entry
    (the script directory comes first) that holds lib/utils.py and lib/layers/."""
    for base in [os.getcwd()] + list(sys.path):
        cand = os.path.join(base or '.', 'lib')
        if os.path.isfile(os.path.join(cand, 'utils.py')) and os.path.isdir(os.path.join(cand, 'layers')):
            return os.path.abspath(cand)
    return None


def install(reference_root=None):
    """Register the package under the reference's module names.  `reference_root` = the reference checkout
    (directory holding `lib/`); default: found on sys.path / cwd, else the non-model `lib.*` modules are
    simply not importable (the model API is complete without them)."""
    from . import implicit_flow, layers, resflow
    from .layers import base, broyden, container, extras, glue, implicit_block, iresblock
    from .layers.base import activations, lipschitz, mixed_lipschitz
    lib_dir = os.path.join(reference_root, 'lib') if reference_root else _find_reference_lib()
    for k in [k for k in sys.modules if k == 'lib' or k.startswith('lib.')]:
        if k == 'lib' or k.startswith(_HOT_PATH):
            del sys.modules[k]
    lib = types.ModuleType('lib')
    lib.__path__ = [lib_dir] if lib_dir else []       # lib.utils, lib.optimizers, ... resolve to the reference's files
    own = {
        'lib.layers': layers, 'lib.implicit_flow': implicit_flow, 'lib.resflow': resflow,
        'lib.layers.base': base, 'lib.layers.base.activations': activations, 'lib.layers.base.lipschitz': lipschitz,
        'lib.layers.base.mixed_lipschitz': mixed_lipschitz, 'lib.layers.broyden': broyden,
        'lib.layers.implicit_block': implicit_block, 'lib.layers.iresblock': iresblock,
        'lib.layers.container': container, 'lib.layers.act_norm': glue, 'lib.layers.squeeze': glue,
        'lib.layers.elemwise': glue, 'lib.layers.normalization': extras, 'lib.layers.coupling': extras,
        'lib.layers.glow': extras,
    }
    sys.modules['lib'] = lib
    for name, mod in own.items():
        sys.modules[name] = mod
    lib.layers, lib.implicit_flow, lib.resflow = layers, implicit_flow, resflow
    return lib


def uninstall():
    for k in [k for k in sys.modules if k == 'lib' or k.startswith('lib.')]:
        del sys.modules[k]
