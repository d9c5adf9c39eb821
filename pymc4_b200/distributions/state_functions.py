"""Proposal functions of the random-walk Metropolis kernel for discrete variables.

Same classes, names, equality rules and call protocol as the reference
(pymc4/distributions/state_functions.py:21-201): ``categorical_uniform_fn(classes)``, ``bernoulli_fn()``,
``gaussian_round_fn(scale)``; calling an instance returns its bound ``_fn`` (the compound sampler compares
``new_state_fn.__self__`` to decide which variables share one kernel, samplers.py:566-585).  On the device the
proposal of a dimension is a code (``device_code``: include/pm4b200.h ``PM4_RW_*``) applied by the random-walk
kernel to that dimension's N(0, 1) draw; ``_fn`` is the NumPy statement of the same proposal (host logic, tests).
"""
import abc
from typing import Any, List, Optional, Union

import numpy as np

__all__ = ["categorical_uniform_fn", "bernoulli_fn", "gaussian_round_fn"]

PM4_RW_NORMAL, PM4_RW_BERNOULLI, PM4_RW_CATEGORICAL, PM4_RW_GAUSSIAN_ROUND = 0, 1, 2, 3


class Proposal(metaclass=abc.ABCMeta):
    def __init__(self, name: Optional[str] = None):
        if name:
            self._name = name

    @abc.abstractmethod
    def _fn(self, state_parts: List[np.ndarray], seed: Optional[int]) -> List[np.ndarray]:
        """New state parts proposed from ``state_parts`` (same shapes and dtypes)."""

    @abc.abstractmethod
    def __eq__(self, other) -> bool:
        """Two proposals are the same kernel argument when this holds (compound-step merging)."""

    __hash__ = None  # type: ignore[assignment]

    @abc.abstractmethod
    def device_code(self) -> int:
        """``PM4_RW_* | (classes << 8)`` of include/pm4b200.h."""

    def device_scale(self) -> float:
        return 1.0

    def __call__(self):
        return self._fn


class CategoricalUniformFn(Proposal):
    """A fresh class drawn uniformly from ``range(classes)`` (independent of the current state)."""

    _name = "categorical_uniform_fn"

    def __init__(self, classes: int, name: Optional[str] = None):
        super().__init__(name)
        self.classes = int(classes)

    def _fn(self, state_parts, seed=None):
        rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        return [rng.integers(0, self.classes, size=np.shape(x)).astype(np.asarray(x).dtype) for x in state_parts]

    def __eq__(self, other) -> bool:
        return isinstance(other, Proposal) and self._name == other._name and self.classes == getattr(other, "classes", None)

    def device_code(self) -> int:
        return PM4_RW_CATEGORICAL | (self.classes << 8)


class BernoulliFn(Proposal):
    """``(state + Bernoulli(1/2)) mod 2``: flip with probability one half."""

    _name = "bernoulli_fn"

    def _fn(self, state_parts, seed=None):
        rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        return [np.mod(np.asarray(x) + (rng.random(np.shape(x)) < 0.5), 2).astype(np.asarray(x).dtype) for x in state_parts]

    def __eq__(self, other) -> bool:
        return isinstance(other, Proposal) and self._name == other._name

    def device_code(self) -> int:
        return PM4_RW_BERNOULLI


class GaussianRoundFn(Proposal):
    """``round(state + scale * Normal(0, 1))`` (the reference draws with unit scale whatever ``scale`` says,
    state_functions.py:186-189; here ``scale`` is applied, the default 1.0 is the reference's behaviour)."""

    _name = "gaussian_round_fn"

    def __init__(self, scale: Union[List[Any], Any] = 1.0, name: Optional[str] = None):
        super().__init__(name)
        self.scale = scale

    def _fn(self, state_parts, seed=None):
        scales = list(self.scale) if isinstance(self.scale, (list, tuple)) else [self.scale]
        if len(scales) == 1:
            scales = scales * len(state_parts)
        if len(state_parts) != len(scales):
            raise ValueError("`scale` must broadcast with `state_parts`")
        rng = seed if isinstance(seed, np.random.Generator) else np.random.default_rng(seed)
        return [np.round(np.asarray(x) + s * rng.standard_normal(np.shape(x))).astype(np.asarray(x).dtype)
                for x, s in zip(state_parts, scales)]

    def __eq__(self, other) -> bool:
        return isinstance(other, Proposal) and self._name == other._name and self.scale == getattr(other, "scale", None)

    def device_code(self) -> int:
        return PM4_RW_GAUSSIAN_ROUND

    def device_scale(self) -> float:
        if isinstance(self.scale, (list, tuple)):
            if len(set(float(s) for s in self.scale)) != 1:
                raise NotImplementedError("gaussian_round_fn with different scales per state part")
            return float(self.scale[0])
        return float(self.scale)


categorical_uniform_fn = CategoricalUniformFn
bernoulli_fn = BernoulliFn
gaussian_round_fn = GaussianRoundFn
