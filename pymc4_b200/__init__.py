"""pymc4_b200 -- B200-native NUTS/HMC backend behind the PyMC4 ``pm.sample`` API."""
from . import utils
from .coroutine_model import Model, model
from .scopes import name_scope, variable_name
from . import coroutine_model
from . import symbolic as math
from . import distributions
from . import flow
from .flow import evaluate_model_transformed, evaluate_model, evaluate_meta_model
from .distributions import *  # noqa: F401,F403
from . import gp
from . import mcmc
from . import inference
from .mcmc.utils import initialize_sampling_state, trace_to_arviz
from .mcmc.samplers import *  # noqa: F401,F403
from .inference.sampling import sample
from .forward_sampling import sample_posterior_predictive, sample_prior_predictive

__version__ = "0.1.0"
