"""Inference entry points."""
from . import sampling  # noqa: F401
from ..mcmc import utils  # noqa: F401  (reference spelling: pm.inference.utils.trace_to_arviz)
