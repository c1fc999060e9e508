from .activations import *  # noqa: F401,F403
from .lipschitz import *  # noqa: F401,F403
from .mixed_lipschitz import *  # noqa: F401,F403
