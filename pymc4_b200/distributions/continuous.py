"""Continuous distributions on the sampler's hot path.

Closed forms are the TFP ``log_prob`` forms the reference delegates to
(pymc4/distributions/continuous.py:108-112 Normal, :234-238 HalfNormal,
:659-663 HalfCauchy -> ``tfd.Normal/HalfNormal/HalfCauchy``); see SURVEY App. B.
"""
import math

import numpy as np

from .distribution import ContinuousDistribution, PositiveContinuousDistribution

__all__ = ["Normal", "HalfNormal", "HalfCauchy"]

_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
_HALF_LOG_HALF_PI = 0.5 * math.log(math.pi / 2.0)
_LOG_2_OVER_PI = math.log(2.0 / math.pi)


class Normal(ContinuousDistribution):
    """Univariate normal, parameterised by ``loc`` and ``scale`` (standard deviation)."""

    _kind = "normal"
    _param_names = ("loc", "scale")

    def __init__(self, name, loc, scale, **kwargs):
        super().__init__(name, loc=loc, scale=scale, **kwargs)

    def _log_prob_numpy(self, value, loc, scale):
        # tfd.Normal divides before subtracting
        return -0.5 * np.square(value / scale - loc / scale) - _HALF_LOG_2PI - np.log(scale)

    def _sample_numpy(self, rng, shape, loc, scale):
        return loc + scale * rng.standard_normal(shape)


class HalfNormal(PositiveContinuousDistribution):
    """Half-normal on [0, inf) with ``scale``; sampled through the Log transform."""

    _kind = "halfnormal"
    _param_names = ("scale",)

    def __init__(self, name, scale, **kwargs):
        super().__init__(name, scale=scale, **kwargs)

    def _log_prob_numpy(self, value, scale):
        lp = -0.5 * np.square(value / scale) - _HALF_LOG_HALF_PI - np.log(scale)
        return np.where(value < 0, -np.inf, lp)

    def _sample_numpy(self, rng, shape, scale):
        return np.abs(scale * rng.standard_normal(shape))


class HalfCauchy(PositiveContinuousDistribution):
    """Half-Cauchy on [0, inf) with ``scale`` (location fixed at 0, as in the reference)."""

    _kind = "halfcauchy"
    _param_names = ("scale",)

    def __init__(self, name, scale, **kwargs):
        super().__init__(name, scale=scale, **kwargs)

    def _log_prob_numpy(self, value, scale):
        lp = _LOG_2_OVER_PI - np.log(scale) - np.log1p(np.square(value / scale))
        return np.where(value < 0, -np.inf, lp)

    def _sample_numpy(self, rng, shape, scale):
        return np.abs(scale * rng.standard_cauchy(shape))
