"""MCMC samplers (NUTS / HMC on the GPU) and trace utilities."""
from . import kernels, samplers, utils  # noqa: F401
from .samplers import HMC, HMCSimple, NUTS, NUTSSimple, RandomWalkM, reg_samplers  # noqa: F401
