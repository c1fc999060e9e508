"""impflow_b200 — B200-native (sm_100a) hot path of Implicit Normalizing Flows.

The package mirrors the reference's Python module API for the path (`layers.imBlock`,
`layers.iResBlock`, `layers.broyden.broyden`, `layers.base.get_linear/get_conv2d`,
`implicit_flow.ImplicitFlow`, `layers.SequentialFlow`) and executes it on hand-written CUDA
kernels reached through the C ABI in include/impflow_b200.h.  `compat.install()` registers the
package under the reference's import names (`lib.layers`, `lib.layers.base`, `lib.implicit_flow`)
so the reference train scripts run unchanged."""
from . import _cabi  # noqa: F401
from . import ops  # noqa: F401
from . import branch_program  # noqa: F401
from . import layers  # noqa: F401
from . import implicit_flow  # noqa: F401
from . import resflow  # noqa: F401
from . import parallel  # noqa: F401
from . import optim  # noqa: F401
from . import compat  # noqa: F401
from .implicit_flow import ImplicitFlow  # noqa: F401
from .resflow import ResidualFlow  # noqa: F401

__version__ = '0.1.0'
