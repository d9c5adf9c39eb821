"""Discrete distributions: observed likelihoods on the sampler's hot path and free (latent) variables of the
compound step.

Closed forms are the TFP ones the reference delegates to
(pymc4/distributions/discrete.py:75-78 ``tfd.Bernoulli(probs=)``, :500-503
``tfd.Poisson(rate=)``, :310-320 ``tfd.FiniteDiscrete(range(K), probs=)`` for ``Categorical``); values are
float-typed integers as in the reference (discrete.py:496).  Free (unobserved) discrete variables cannot be sampled
by the gradient-based kernels (reference samplers.py:62-66 raises ValueError): ``pm.sample`` gives them to the
compound step, which moves them with the random-walk proposals of ``state_functions`` (Bernoulli and Categorical
carry their default proposal like the reference, discrete.py:73, :315).
"""
import math

import numpy as np

from .. import symbolic as sym
from .distribution import BoundedDiscreteDistribution, PositiveDiscreteDistribution
from .state_functions import bernoulli_fn, categorical_uniform_fn

__all__ = ["Bernoulli", "Poisson", "Categorical"]

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
        self._default_new_state_part = bernoulli_fn()

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
        return np.asarray(rng.poisson(np.broadcast_to(rate, shape)), dtype=np.float64)


class Categorical(BoundedDiscreteDistribution):
    """Categorical over ``range(K)`` with constant class probabilities ``probs[..., K]``.

    On the device the log-mass of a free value is the interpolating polynomial of ``log probs`` through the integer
    nodes ``0 .. K-1`` (exact at every class, built from the element-wise ops of the model tape), lowered as an
    additive term of the joint density."""

    _grad_support = False
    _kind = "categorical"
    _param_names = ("probs",)

    def __init__(self, name, probs, **kwargs):
        super().__init__(name, probs=probs, **kwargs)
        self.classes = int(self.params["probs"].shape[-1])
        self._default_new_state_part = categorical_uniform_fn(classes=self.classes)

    @property
    def _param_batch_shape(self):
        return tuple(self.params["probs"].shape[:-1])

    def lower_limit(self):
        return 0.0

    def upper_limit(self):
        return float(self.classes)

    def _probs(self):
        try:
            return np.asarray(sym.as_numpy(self.params["probs"]), dtype=np.float64)
        except TypeError as e:
            raise NotImplementedError("Categorical needs constant `probs` (a symbolic class-probability vector is not "
                                      "lowered to the device)") from e

    def log_prob_expr(self, value):
        """Symbolic log-mass of ``value`` (an expression or data): sum_k log p_k L_k(value), L_k the Lagrange basis."""
        probs = self._probs()
        K = probs.shape[-1]
        with np.errstate(divide="ignore"):
            logp = np.maximum(np.log(probs), -690.0)  # a class of probability zero: effectively never accepted
        x = sym.as_expr(value)
        total = None
        for k in range(K):
            denom = float(np.prod([k - m for m in range(K) if m != k])) if K > 1 else 1.0
            term = sym.as_expr(logp[..., k] / denom)
            for m in range(K):
                if m != k:
                    term = term * (x - float(m))
            total = term if total is None else total + term
        return total

    def _log_prob_numpy(self, value, probs):
        idx = np.clip(np.asarray(value, dtype=np.int64), 0, probs.shape[-1] - 1)
        with np.errstate(divide="ignore"):
            lp = np.log(np.take_along_axis(np.broadcast_to(probs, idx.shape + probs.shape[-1:]), idx[..., None], -1)[..., 0])
        return np.where((value < 0) | (value >= probs.shape[-1]), -np.inf, lp)

    def _sample_numpy(self, rng, shape, probs):
        flat = np.broadcast_to(probs, tuple(shape) + probs.shape[-1:]).reshape(-1, probs.shape[-1])
        u = rng.random((flat.shape[0], 1))
        return (np.cumsum(flat, axis=1) < u).sum(axis=1).clip(0, probs.shape[-1] - 1).reshape(shape).astype(np.float64)
