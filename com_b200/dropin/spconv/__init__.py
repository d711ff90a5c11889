"""Drop-in `spconv` package backed by com_b200 (put com_b200/dropin on PYTHONPATH or call
com_b200.install_dropins()).  Only the surface the COM hot path touches is provided."""
from . import pytorch, utils  # noqa: F401
from .pytorch import *  # noqa: F401,F403  (spconv 1.x style `import spconv` users)

__version__ = "2.3.6+comb200"
