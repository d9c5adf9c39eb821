"""MCMC samplers (NUTS / HMC on the GPU) and trace utilities."""
from . import kernels, samplers, utils  # noqa: F401
from .samplers import HMC, NUTS, reg_samplers  # noqa: F401
