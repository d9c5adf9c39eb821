"""Unconstraining transforms for bounded distributions.

Follows the reference's contract (pymc4/distributions/transforms.py:13-181):
``forward`` maps the constrained value x to the unconstrained z, ``inverse``
maps z back, and ``forward_log_det_jacobian(x)`` is log|dz/dx|.  The transformed
executor adds ``-forward_log_det_jacobian(x)`` as a potential
(pymc4/flow/transformed_executor.py:90-101; the reference's ``BackwardTransform``
never overrides ``jacobian_preference`` -- transforms.py:119 assigns a
differently-named attribute -- so the Forward branch is the one always taken).

All supported transforms are of the form ``x = shift + scale * f(z)`` with
``f`` in {exp, sigmoid}; the sampler kernels apply exactly that map (and its
derivative for the chain rule), so each transform carries ``ir_code`` /
``ir_shift`` / ``ir_scale``.
"""
import enum
from typing import Optional

import numpy as np

from .. import symbolic as sym

__all__ = ["Log", "Sigmoid", "LowerBound", "UpperBound", "Interval"]

# transform kinds understood by the kernels (include/pm4b200.h: PM4_TR_*)
TR_IDENTITY, TR_EXP, TR_SIGMOID = 0, 1, 2


class JacobianPreference(enum.Enum):
    Forward = "Forward"
    Backward = "Backward"


class Transform:
    name: Optional[str] = None
    jacobian_preference = JacobianPreference.Forward
    ir_code = TR_IDENTITY
    ir_shift = 0.0
    ir_scale = 1.0

    def forward(self, x):
        raise NotImplementedError

    def inverse(self, z):
        raise NotImplementedError

    def forward_log_det_jacobian(self, x):
        raise NotImplementedError

    def inverse_log_det_jacobian(self, z):
        return -self.forward_log_det_jacobian(self.inverse(z))

    # ------------------------------------------------------------------
    def _native_inverse(self, z):
        """If ``z`` is the whole unconstrained value of a sampled variable, return the
        matching constrained view: the kernels compute x = shift + scale*f(z) themselves."""
        if isinstance(z, sym.Expr) and z.op == "view" and z.space == "z":
            return sym.Expr("view", shape=z.shape, rv=z.rv, space="x", index=z.index, rv_size=z.rv_size)
        return None


class _AffineOf(Transform):
    """x = shift + scale * f(z)."""

    def _f(self, z):
        raise NotImplementedError

    def _finv(self, u):
        raise NotImplementedError

    def inverse(self, z):
        native = self._native_inverse(z)
        if native is not None:
            return native
        return self.ir_shift + self.ir_scale * self._f(z)

    def forward(self, x):
        return self._finv((x - self.ir_shift) / self.ir_scale)


class Log(_AffineOf):
    """Positive support: z = log x (reference transforms.py:137-143)."""

    name = "log"
    ir_code = TR_EXP

    def _f(self, z):
        return sym.exp(z)

    def _finv(self, u):
        return sym.log(u)

    def forward_log_det_jacobian(self, x):
        # tfb.Exp().inverse_log_det_jacobian(x) = -log x
        return -sym.log(x)


class Sigmoid(_AffineOf):
    """Unit-interval support: z = logit x (reference transforms.py:146-151)."""

    name = "sigmoid"
    ir_code = TR_SIGMOID

    def _f(self, z):
        return sym.sigmoid(z)

    def _finv(self, u):
        return sym.log(u) - sym.log1p(-u)

    def forward_log_det_jacobian(self, x):
        # tfb.Sigmoid().inverse_log_det_jacobian(x) = -log x - log1p(-x)
        return -sym.log(x) - sym.log1p(-x)


class LowerBound(Log):
    """Support [lower_limit, inf): x = lower + exp z (reference transforms.py:154-161)."""

    name = "lowerbound"

    def __init__(self, lower_limit):
        self.ir_shift = float(np.asarray(sym.as_numpy(lower_limit)))

    def forward_log_det_jacobian(self, x):
        return -sym.log(x - self.ir_shift)


class UpperBound(Log):
    """Support (-inf, upper_limit]: x = upper - exp z (reference transforms.py:164-171)."""

    name = "upperbound"
    ir_scale = -1.0

    def __init__(self, upper_limit):
        self.ir_shift = float(np.asarray(sym.as_numpy(upper_limit)))

    def forward_log_det_jacobian(self, x):
        return -sym.log(self.ir_shift - x)


class Interval(Sigmoid):
    """Support [lower, upper]: x = lower + (upper-lower) sigmoid z (reference transforms.py:174-181)."""

    name = "interval"

    def __init__(self, lower_limit, upper_limit):
        lo = float(np.asarray(sym.as_numpy(lower_limit)))
        hi = float(np.asarray(sym.as_numpy(upper_limit)))
        self.ir_shift = lo
        self.ir_scale = hi - lo

    def forward_log_det_jacobian(self, x):
        lo, width = self.ir_shift, self.ir_scale
        return -sym.log(x - lo) - sym.log((lo + width) - x) + float(np.log(width))
