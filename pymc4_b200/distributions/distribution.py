"""Distribution / Potential / Deterministic model nodes.

Mirrors the reference's node contract (pymc4/distributions/distribution.py:29-358):
constructor kwargs ``transform / observed / batch_stack / event_stack /
conditionally_independent / reinterpreted_batch_ndims``, test values, the
class hierarchy that fixes the default unconstraining transform -- but the
parameters are symbolic expressions (``pymc4_b200.symbolic``) instead of a
wrapped ``tfd.Distribution``, because the joint log-density is lowered to a flat
IR and evaluated by CUDA device functions, not by TensorFlow-Probability.

Each concrete distribution names its IR term kind (``_kind``) and its ordered
parameter list (``_param_names``); the log-density closed forms live in
``csrc/pm4_dists.cuh`` (device) and ``oracle/pm4_oracle.c`` (CPU checker).  The
NumPy ``log_prob`` here follows the same closed forms and is
host-side model introspection only -- the sampler never calls it.
"""
import abc
import copy
import warnings
from typing import Any, Optional, Tuple, Union

import numpy as np

from .. import symbolic as sym
from ..coroutine_model import Model, unpack
from . import transforms

NameType = Union[str, int]

__all__ = (
    "Distribution",
    "Potential",
    "Deterministic",
    "ContinuousDistribution",
    "DiscreteDistribution",
    "PositiveContinuousDistribution",
    "PositiveDiscreteDistribution",
    "BoundedDistribution",
    "BoundedDiscreteDistribution",
    "BoundedContinuousDistribution",
    "UnitContinuousDistribution",
)


def _as_shape(v) -> Tuple[int, ...]:
    if v is None:
        return ()
    if isinstance(v, (int, np.integer)):
        return (int(v),)
    return tuple(int(s) for s in np.atleast_1d(np.asarray(v)).tolist())


class Distribution(Model):
    """Statistical distribution node of a coroutine model."""

    _grad_support: bool = True
    _test_value: Any = 0.0
    _base_parameters = ["dtype", "validate_args", "allow_nan_stats"]
    _kind: Optional[str] = None  # IR term kind (ir.TERM_KINDS)
    _param_names: Tuple[str, ...] = ()

    def __init__(
        self,
        name: Optional[NameType],
        *,
        transform=None,
        observed=None,
        batch_stack=None,
        event_stack=None,
        conditionally_independent=False,
        reinterpreted_batch_ndims=0,
        dtype=None,
        validate_args=False,
        allow_nan_stats=False,
        **kwargs,
    ):
        self.conditions, self.base_parameters = self.unpack_conditions(
            dtype=dtype, validate_args=validate_args, allow_nan_stats=allow_nan_stats, **kwargs
        )
        # parameters become expressions (constants fold to Const nodes)
        self.params = {k: sym.as_expr(v) for k, v in self.conditions.items()}
        self._default_new_state_part = None
        super().__init__(self.unpack_distribution, name=name, keep_return=True, keep_auxiliary=False)
        if name is None and observed is not None:
            raise ValueError(
                "Observed variables are not allowed for anonymous (with name=None) Distributions"
            )
        self.model_info.update(observed=observed)
        self.transform = self._init_transform(transform)
        self.batch_stack = batch_stack
        self.event_stack = event_stack
        self.conditionally_independent = conditionally_independent
        self.reinterpreted_batch_ndims = int(reinterpreted_batch_ndims or 0)

    # -- construction helpers ----------------------------------------------
    def _init_transform(self, transform):
        return transform

    def unpack_distribution(self):
        return unpack(self)

    @classmethod
    def unpack_conditions(cls, **kwargs) -> Tuple[dict, dict]:
        base = {k: v for k, v in kwargs.items() if k in cls._base_parameters}
        if base.get("dtype") is None:
            base.pop("dtype", None)
        conditions = {k: v for k, v in kwargs.items() if k not in cls._base_parameters}
        return conditions, base

    @classmethod
    def dist(cls, *args, **kwargs):
        """Anonymous distribution; give it a name later with :meth:`prior`."""
        return cls(None, *args, **kwargs)

    def prior(self, name: NameType, *, transform=None, observed=None):
        if not self.is_anonymous:
            raise TypeError("Distribution is already not anonymous and cant define a new prior")
        if name is None:
            raise ValueError("Can't create a prior Distribution without a name")
        clone = copy.copy(self)
        clone.model_info = dict(self.model_info, observed=observed)
        clone.conditions = dict(self.conditions)
        clone.params = dict(self.params)
        if transform is not None:
            clone.transform = transform
        clone.name = clone.validate_name(name)
        clone.model_info["name"] = clone.name
        from ..scopes import name_scope

        clone.model_info["scope"] = name_scope(clone.name)
        clone.genfn = clone.unpack_distribution
        return clone

    # -- shapes ------------------------------------------------------------
    @property
    def _param_batch_shape(self) -> Tuple[int, ...]:
        shapes = [p.shape for p in self.params.values()]
        return tuple(np.broadcast_shapes(*shapes)) if shapes else ()

    @property
    def batch_shape(self) -> Tuple[int, ...]:
        pb = self._param_batch_shape
        k = self.reinterpreted_batch_ndims
        base = pb[: len(pb) - k] if k else pb
        return _as_shape(self.batch_stack) + base

    @property
    def event_shape(self) -> Tuple[int, ...]:
        pb = self._param_batch_shape
        k = self.reinterpreted_batch_ndims
        ev = pb[len(pb) - k :] if k else ()
        return _as_shape(self.event_stack) + ev

    @property
    def dist_shape(self) -> Tuple[int, ...]:
        return self.batch_shape + self.event_shape

    @property
    def dtype(self):
        return np.dtype("float32")

    @property
    def validate_args(self):
        return self.base_parameters.get("validate_args", False)

    @property
    def allow_nan_stats(self):
        return self.base_parameters.get("allow_nan_stats", False)

    # -- values ------------------------------------------------------------
    @property
    def test_value(self) -> np.ndarray:
        return np.broadcast_to(np.asarray(self._test_value, dtype=np.float64), self.dist_shape).copy()

    def get_test_sample(self, sample_shape=(), seed=None) -> np.ndarray:
        return np.broadcast_to(self.test_value, _as_shape(sample_shape) + self.dist_shape).copy()

    def _concrete_params(self):
        try:
            return {k: sym.as_numpy(v) for k, v in self.params.items()}
        except TypeError as e:
            raise TypeError(
                "{} {!r} has symbolic parameters; this operation needs concrete values".format(
                    type(self).__name__, self.name
                )
            ) from e

    def _sample_numpy(self, rng: np.random.Generator, shape, **params) -> np.ndarray:
        raise NotImplementedError("{} has no forward sampler".format(type(self).__name__))

    def sample(self, sample_shape=(), seed=None) -> np.ndarray:
        """Host-side forward draw (NumPy); used by the untransformed executor passes only."""
        rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        shape = _as_shape(sample_shape) + self.dist_shape
        return np.asarray(self._sample_numpy(rng, shape, **self._concrete_params()), dtype=np.float64)

    sample_numpy = sample

    def _log_prob_numpy(self, value: np.ndarray, **params) -> np.ndarray:
        raise NotImplementedError

    def log_prob(self, value):
        """Element-wise log-density (NumPy, float64) for concrete ``value`` and parameters."""
        value = np.asarray(sym.as_numpy(value), dtype=np.float64)
        with np.errstate(all="ignore"):
            return np.asarray(self._log_prob_numpy(value, **self._concrete_params()))

    log_prob_numpy = log_prob

    # -- flags -------------------------------------------------------------
    @property
    def is_anonymous(self) -> bool:
        return self.name is None

    @property
    def is_observed(self) -> bool:
        return self.model_info["observed"] is not None

    @property
    def is_root(self) -> bool:
        return self.conditionally_independent


class Potential:
    """An arbitrary additive term of the joint log-density."""

    __slots__ = ("_value", "_coef")

    def __init__(self, value, coef=1.0):
        self._value = value
        self._coef = coef

    @property
    def value(self):
        v = self._value() if callable(self._value) else self._value
        return v * self._coef

    @property
    def value_numpy(self):
        return sym.as_numpy(self.value)


class Deterministic(Model):
    """A named value with no log-density."""

    __slots__ = ("value",)

    def __init__(self, name: Optional[NameType], value: Any):
        self.value = value
        super().__init__(self.get_value, name=name, keep_return=True, keep_auxiliary=False)

    def get_value(self):
        return self.value

    @property
    def value_numpy(self):
        return sym.as_numpy(self.value)

    @property
    def is_anonymous(self):
        return self.name is None


# ---------------------------------------------------------------------------
# class hierarchy fixing default transform + test value (reference :267-358)
# ---------------------------------------------------------------------------
class ContinuousDistribution(Distribution):
    _test_value = 0.0

    @classmethod
    def unpack_conditions(cls, **kwargs):
        conditions, base = super().unpack_conditions(**kwargs)
        dtype = base.pop("dtype", None)
        if dtype is not None:
            warnings.warn(
                "Continuous distributions compute in the sampler's dtype; the supplied "
                "dtype={} will be ignored.".format(dtype)
            )
        return conditions, base


class DiscreteDistribution(Distribution):
    _test_value = 0


class BoundedDistribution(Distribution):
    @abc.abstractmethod
    def lower_limit(self):
        raise NotImplementedError

    @abc.abstractmethod
    def upper_limit(self):
        raise NotImplementedError


class BoundedDiscreteDistribution(DiscreteDistribution, BoundedDistribution):
    @property
    def _test_value(self):
        return np.round(0.5 * (self.upper_limit() + self.lower_limit()))


class BoundedContinuousDistribution(ContinuousDistribution, BoundedDistribution):
    def _init_transform(self, transform):
        if transform is None:
            return transforms.Interval(self.lower_limit(), self.upper_limit())
        return transform

    @property
    def _test_value(self):
        return 0.5 * (self.upper_limit() + self.lower_limit())


class UnitContinuousDistribution(BoundedContinuousDistribution):
    def _init_transform(self, transform):
        return transforms.Sigmoid() if transform is None else transform

    def lower_limit(self):
        return 0.0

    def upper_limit(self):
        return 1.0


class PositiveContinuousDistribution(BoundedContinuousDistribution):
    _test_value = 1.0

    def _init_transform(self, transform):
        return transforms.Log() if transform is None else transform

    def lower_limit(self):
        return 0.0

    def upper_limit(self):
        return float("inf")


class PositiveDiscreteDistribution(BoundedDiscreteDistribution):
    _test_value = 1

    def lower_limit(self):
        return 0

    def upper_limit(self):
        return float("inf")
