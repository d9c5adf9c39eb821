"""Discrete distributions usable as observed likelihoods on the sampler's hot path.

Closed forms are the TFP ones the reference delegates to
(pymc4/distributions/discrete.py:75-78 ``tfd.Bernoulli(probs=)``, :500-503
``tfd.Poisson(rate=)``); values are float-typed integers as in the reference
(discrete.py:496).  Free (unobserved) discrete variables cannot be sampled by
the gradient-based kernels (reference samplers.py:62-66 raises ValueError).
"""
import math

import numpy as np

from .distribution import BoundedDiscreteDistribution, PositiveDiscreteDistribution

__all__ = ["Bernoulli", "Poisson"]

_lgamma = np.vectorize(math.lgamma, otypes=[float])


def _xlogy(x, y):
    with np.errstate(all="ignore"):
        return np.where(x == 0, 0.0, x * np.log(y))


def _xlog1py(x, y):
    with np.errstate(all="ignore"):
        return np.where(x == 0, 0.0, x * np.log1p(y))


class Bernoulli(BoundedDiscreteDistribution):
    """Bernoulli with success probability ``probs``."""

    _grad_support = False
    _kind = "bernoulli"
    _param_names = ("probs",)

    def __init__(self, name, probs, **kwargs):
        super().__init__(name, probs=probs, **kwargs)

    def lower_limit(self):
        return 0.0

    def upper_limit(self):
        return 1.0

    def _log_prob_numpy(self, value, probs):
        return _xlogy(value, probs) + _xlog1py(1.0 - value, -probs)

    def _sample_numpy(self, rng, shape, probs):
        return (rng.random(shape) < probs).astype(np.float64)


class Poisson(PositiveDiscreteDistribution):
    """Poisson with mean ``rate``."""

    _kind = "poisson"
    _param_names = ("rate",)

    def __init__(self, name, rate, **kwargs):
        super().__init__(name, rate=rate, **kwargs)

    def _log_prob_numpy(self, value, rate):
        return _xlogy(value, rate) - rate - _lgamma(value + 1.0)

    def _sample_numpy(self, rng, shape, rate):
        return rng.poisson(np.broadcast_to(rate, shape)).astype(np.float64)
