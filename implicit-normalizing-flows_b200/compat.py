"""Drop-in registration under the reference's import names.

    import impflow_b200; impflow_b200.compat.install()
    import lib.layers as layers; import lib.layers.base as base_layers
    from lib.implicit_flow import ImplicitFlow

after which train_toy.py / train_tabular.py / train_img.py / train_classification.py of the
reference resolve their model code to this package."""
import sys
import types


def install():
    from . import implicit_flow, layers
    from .layers import base, broyden, implicit_block, iresblock, container
    lib = types.ModuleType('lib')
    lib.__path__ = []
    lib.layers = layers
    lib.implicit_flow = implicit_flow
    sys.modules['lib'] = lib
    sys.modules['lib.layers'] = layers
    sys.modules['lib.layers.base'] = base
    sys.modules['lib.layers.broyden'] = broyden
    sys.modules['lib.layers.implicit_block'] = implicit_block
    sys.modules['lib.layers.iresblock'] = iresblock
    sys.modules['lib.layers.container'] = container
    sys.modules['lib.implicit_flow'] = implicit_flow
    return lib
