"""GP helpers (/root/reference/pymc4/gp/util.py:15-21)."""
import numpy as np

from .. import symbolic as sym

__all__ = ["stabilize"]


def stabilize(K, shift=None):
    """Add a diagonal shift to a covariance matrix (reference default: 1e-4 in float32, 1e-6 in float64; the
    device factorises in float64 whatever the sampling dtype, the float32 default is kept)."""
    if shift is None:
        shift = 1e-4
    K = sym.as_expr(K)
    if K.ndim != 2 or K.shape[0] != K.shape[1]:
        raise ValueError("stabilize expects a square matrix, got shape {}".format(K.shape))
    out = K + shift * np.eye(K.shape[0])
    return out.data if isinstance(out, sym.Expr) and out.op == "const" else out
