"""Covariance functions (/root/reference/pymc4/gp/cov.py).

``ExpQuad`` mirrors the reference class (cov.py:399-475): ``k(x, x') = amplitude^2 exp(-|x - x'|^2 / (2 l^2))``
with ``feature_ndims = 1`` and optional ``active_dims`` / ``scale_diag``.  With concrete parameters a call
returns the NumPy matrix; with free random variables as parameters it returns an expression node that
``pm.MvNormal`` lowers to the dense GP term of the device (csrc/pm4_gp.cu).
"""
from numbers import Number
from typing import Iterable, Optional, Union

import numpy as np

from .. import symbolic as sym

__all__ = ["Covariance", "Stationary", "ExpQuad"]


class Covariance:
    """Base class of the covariance functions (cov.py:84-298)."""

    def __init__(self, feature_ndims=1, active_dims=None, scale_diag=None, **kwargs):
        if not isinstance(feature_ndims, (int, np.integer)) or feature_ndims <= 0:
            raise ValueError(
                "expected 'feature_ndims' to be an integer greater or equal to 1"
                " but found {}.".format(feature_ndims)
            )
        if feature_ndims != 1:
            raise NotImplementedError("feature_ndims > 1 is not built; flatten the feature axes")
        kwargs.pop("ARD", None)
        if kwargs:
            raise TypeError("unexpected keyword arguments {}".format(sorted(kwargs)))
        self._feature_ndims = int(feature_ndims)
        # reference semantics (cov.py:115-146): an int k (or [k]) keeps the first k features, a nested list
        # [[i, j, ...]] keeps those feature columns
        self._columns = None
        if active_dims is not None:
            if isinstance(active_dims, (int, np.integer)):
                active_dims = (int(active_dims),)
            else:
                active_dims = tuple(active_dims)
            if len(active_dims) > self._feature_ndims:
                raise ValueError(
                    "'active_dims' contain more entries than number of feature dimensions."
                    " Consider increasing the 'feature_ndims' or decreasing the entries"
                    " in 'active_dims'. expected len(active_dims) < feature_ndims but got"
                    " {} > {}".format(len(active_dims), self._feature_ndims)
                )
            dim = active_dims[0]
            if isinstance(dim, slice):
                self._columns = dim
            elif isinstance(dim, (int, np.integer)):
                self._columns = slice(0, int(dim))
            else:
                self._columns = [int(i) for i in dim]
        self._active_dims = active_dims
        self._scale_diag = scale_diag
        self._ard = True

    @property
    def feature_ndims(self) -> int:
        return self._feature_ndims

    @property
    def active_dims(self):
        return self._active_dims

    @property
    def scale_diag(self):
        return self._scale_diag

    @property
    def ard(self) -> bool:
        return self._ard

    def _slice(self, X1, X2):
        """Keep the active feature columns (cov.py:164-180) and apply the ARD scaling."""
        X1, X2 = np.asarray(sym.as_numpy(X1), dtype=np.float64), np.asarray(sym.as_numpy(X2), dtype=np.float64)
        if X1.ndim == 1:
            X1 = X1[:, None]
        if X2.ndim == 1:
            X2 = X2[:, None]
        if X1.ndim != 2 or X2.ndim != 2 or X1.shape[1] != X2.shape[1]:
            raise ValueError("expected points of shape [n, d], got {} and {}".format(X1.shape, X2.shape))
        if self._columns is not None:
            X1, X2 = X1[:, self._columns], X2[:, self._columns]
        if self._scale_diag is not None:
            sd = np.asarray(sym.as_numpy(self._scale_diag), dtype=np.float64)
            X1, X2 = X1 / sd, X2 / sd
        return X1, X2

    def _matrix(self, X1, X2):
        raise NotImplementedError("Your Covariance class should override this method.")

    def evaluate_kernel(self, X1, X2):
        """Point-wise evaluation ``k(X1[i], X2[i])`` (cov.py:189-205)."""
        X1, X2 = self._slice(X1, X2)
        return self._pointwise(X1, X2)

    def __call__(self, X1, X2, diag=False, to_dense=True):
        """Covariance matrix between the points ``X1`` and ``X2`` (cov.py:227-272)."""
        X1, X2 = self._slice(X1, X2)
        cov = self._matrix(X1, X2)
        if diag and not to_dense:  # the reference's _diag (cov.py:182-187) returns the full matrix when to_dense
            if isinstance(cov, sym.Expr):
                raise NotImplementedError("diag=True needs concrete kernel parameters")
            return np.diag(cov).copy()
        return cov


class Stationary(Covariance):
    """Base class of the stationary covariance functions (cov.py:376-396)."""

    @property
    def length_scale(self):
        return self._length_scale


class ExpQuad(Stationary):
    """Exponentiated quadratic covariance function (cov.py:399-475)."""

    def __init__(self, length_scale, amplitude: Union[float, "sym.Expr"] = 1.0, feature_ndims: int = 1,
                 active_dims: Optional[Union[int, Iterable]] = None, scale_diag: Optional[Union[Number, Iterable]] = None,
                 **kwargs):
        self._amplitude = amplitude
        self._length_scale = length_scale
        super().__init__(feature_ndims=feature_ndims, active_dims=active_dims, scale_diag=scale_diag, **kwargs)

    @property
    def amplitude(self):
        return self._amplitude

    def _concrete(self):
        return not (sym.is_symbolic(self._length_scale) or sym.is_symbolic(self._amplitude))

    def _matrix(self, X1, X2):
        if self._concrete():
            return sym.expquad_numpy(X1, X2, sym.as_numpy(self._length_scale), sym.as_numpy(self._amplitude))
        return sym.expquad(X1, X2, self._length_scale, self._amplitude)

    def _pointwise(self, X1, X2):
        if not self._concrete():
            raise NotImplementedError("point-wise evaluation needs concrete kernel parameters")
        d2 = np.sum(np.square(X1 - X2), axis=-1)
        ls, amp = sym.as_numpy(self._length_scale), sym.as_numpy(self._amplitude)
        return np.square(amp) * np.exp(-d2 / (2.0 * np.square(ls)))
