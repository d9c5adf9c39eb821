"""Gaussian-process building blocks on the sampler's hot path (/root/reference/pymc4/gp/__init__.py).

Only what the marginal GP regression of BASELINE configs[3] needs is built: the ExpQuad covariance
function and ``stabilize``; the model is spelled with ``pm.MvNormal(..., covariance_matrix=k(X, X) + noise,
observed=y)`` (the reference has no ``marginal_likelihood``, /root/reference/pymc4/gp/gp.py:51-60).
"""
from . import cov, util
from .cov import ExpQuad

__all__ = ["cov", "util", "ExpQuad"]
