"""Model nodes: distributions, potentials, deterministics and transforms."""
from .continuous import *  # noqa: F401,F403
from .discrete import *  # noqa: F401,F403
from .multivariate import *  # noqa: F401,F403
from .distribution import Potential, Deterministic, Distribution  # noqa: F401
from . import transforms  # noqa: F401
from .state_functions import *  # noqa: F401,F403
from . import state_functions
from . import continuous, discrete, multivariate

__all__ = list(continuous.__all__) + list(discrete.__all__) + list(multivariate.__all__) + list(state_functions.__all__) + ["Potential", "Deterministic", "Distribution", "transforms"]
