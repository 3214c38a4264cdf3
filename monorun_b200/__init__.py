"""monorun_b200 -- B200-native uncertainty-PnP hot path of MonoRUn behind the reference's plugin surface.

Importing the package registers the drop-in classes (``PnPUncert`` in ``PNP``; ``UncertPropPnPOptimizer``,
``FCNNOCDecoder``, ``UncertProjectionHead``, ``MonoRUnRoIHead`` in ``HEADS``; the coders), the way
``import monorun`` does for the reference (monorun/__init__.py:1-5)."""
from . import registry  # noqa: F401
from . import coders  # noqa: F401
from . import pnp  # noqa: F401
from . import heads  # noqa: F401
from .pnp import PnPUncert, pnp_uncert, solve_batched, solve_host, u2d_pnp_cpu  # noqa: F401
from .registry import PNP, HEADS, build_pnp, build_head  # noqa: F401

__all__ = ['u2d_pnp_cpu', 'build_pnp', 'PnPUncert', 'pnp_uncert',   # monorun/ops/least_squares/__init__.py:1-5
           'solve_batched', 'solve_host', 'PNP', 'HEADS', 'build_head']
