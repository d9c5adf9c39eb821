"""Symbolic tensor expressions: the glue between coroutine models and the flat model IR.

The reference lets model code mix random variables with TensorFlow ops
(``tf.gather(a, idx) + tf.gather(b, idx) * x``, ``mu + tau * eta`` ...;
notebooks/radon_hierarchical.ipynb:72-95, tests/test_8schools.py:12-21) and
relies on ``tf.function`` tracing + autodiff.  There is no TensorFlow here, so
random variables are ``Expr`` nodes with static shapes; operator overloading and
the functions in this module (exported as ``pm.math``) build a small DAG that
``pymc4_b200.ir`` lowers to the element-wise tape the CUDA code generator and
the CPU oracle consume.  ``Expr.eval`` is a NumPy interpreter used only for
host logic (test values, shapes, trace post-processing checks) -- never for
sampling.
"""
from __future__ import annotations

import math as _math
from typing import Any, Dict, Optional, Sequence, Tuple, Union

import numpy as np

ArrayLike = Union["Expr", np.ndarray, float, int, Sequence]

_UNARY_NP = {
    "neg": np.negative,
    "exp": np.exp,
    "log": np.log,
    "log1p": np.log1p,
    "sqrt": np.sqrt,
    "square": np.square,
    "sigmoid": lambda v: 1.0 / (1.0 + np.exp(-v)),
    "softplus": lambda v: np.logaddexp(0.0, v),
    "tanh": np.tanh,
    "abs": np.abs,
    "recip": lambda v: 1.0 / v,
    "lgamma": np.vectorize(_math.lgamma, otypes=[float]),
}
_BINARY_NP = {
    "add": np.add,
    "sub": np.subtract,
    "mul": np.multiply,
    "div": np.divide,
    "pow": np.power,
}
UNARY_OPS = tuple(_UNARY_NP)
BINARY_OPS = tuple(_BINARY_NP)
REDUCE_OPS = ("reduce_sum", "reduce_mean")


class Expr:
    """A node of the expression DAG.

    ``op`` is one of
      * ``"const"``  -- ``data`` holds a float64 ndarray;
      * ``"view"``   -- a gather out of a free random variable: ``rv`` names it
        (full scoped name of the *sampled* variable, e.g. ``m/__log_tau``),
        ``space`` is ``"x"`` (constrained value) or ``"z"`` (unconstrained),
        ``index`` is an int32 ndarray of flat positions inside the variable,
        shaped like the view;
      * a unary / binary / reduce op name with ``args``.
    """

    __slots__ = ("op", "args", "shape", "data", "rv", "space", "index", "rv_size", "_hash")
    __array_ufunc__ = None  # make ndarray <op> Expr defer to Expr.__r<op>__
    __array_priority__ = 1000

    def __init__(self, op, args=(), shape=(), data=None, rv=None, space=None, index=None, rv_size=None):
        self.op = op
        self.args = tuple(args)
        self.shape = tuple(int(s) for s in shape)
        self.data = data
        self.rv = rv
        self.space = space
        self.index = index
        self.rv_size = rv_size
        self._hash = None

    # -- basic properties --------------------------------------------------
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def dtype(self):
        return np.dtype("float32")

    def __repr__(self) -> str:
        if self.op == "const":
            return "Const(shape={})".format(self.shape)
        if self.op == "view":
            return "View({}:{}, shape={})".format(self.space, self.rv, self.shape)
        return "Expr({}, shape={})".format(self.op, self.shape)

    def __bool__(self):
        raise TypeError(
            "The truth value of a symbolic random variable is undefined; "
            "data-dependent Python control flow cannot be lowered to the model IR."
        )

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of a 0-d symbolic value")
        return self.shape[0]

    def __iter__(self):
        if not self.shape:
            raise TypeError("iteration over a 0-d symbolic value")
        return (self[i] for i in range(self.shape[0]))

    # -- operators -----------------------------------------------------------
    def __add__(self, o):
        return _binary("add", self, o)

    def __radd__(self, o):
        return _binary("add", o, self)

    def __sub__(self, o):
        return _binary("sub", self, o)

    def __rsub__(self, o):
        return _binary("sub", o, self)

    def __mul__(self, o):
        return _binary("mul", self, o)

    def __rmul__(self, o):
        return _binary("mul", o, self)

    def __truediv__(self, o):
        return _binary("div", self, o)

    def __rtruediv__(self, o):
        return _binary("div", o, self)

    def __pow__(self, o):
        if isinstance(o, (int, float)) and float(o) == 2.0:
            return _unary("square", self)
        return _binary("pow", self, o)

    def __rpow__(self, o):
        return _binary("pow", o, self)

    def __neg__(self):
        return _unary("neg", self)

    def __pos__(self):
        return self

    def __matmul__(self, o):
        return matvec(self, o) if as_expr(o).ndim == 1 else NotImplemented

    def __rmatmul__(self, o):
        return matvec(o, self)

    def __getitem__(self, key):
        return _getitem(self, key)

    # -- NumPy interpreter (host logic only) ------------------------------
    def eval(self, env: Dict[str, np.ndarray], _memo: Optional[dict] = None) -> np.ndarray:
        """Evaluate with ``env`` mapping ``"<space>:<rv name>"`` to arrays."""
        memo = {} if _memo is None else _memo
        key = id(self)
        if key in memo:
            return memo[key]
        if self.op == "const":
            out = self.data
        elif self.op == "view":
            base = np.asarray(env[self.space + ":" + self.rv], dtype=np.float64).reshape(-1)
            out = base[self.index]
        elif self.op in _UNARY_NP:
            out = _UNARY_NP[self.op](self.args[0].eval(env, memo))
        elif self.op in _BINARY_NP:
            out = _BINARY_NP[self.op](self.args[0].eval(env, memo), self.args[1].eval(env, memo))
        elif self.op == "reduce_sum":
            out = np.sum(self.args[0].eval(env, memo))
        elif self.op == "reduce_mean":
            out = np.mean(self.args[0].eval(env, memo))
        elif self.op == "broadcast":
            out = np.broadcast_to(self.args[0].eval(env, memo), self.shape)
        elif self.op == "reshape":
            out = np.reshape(self.args[0].eval(env, memo), self.shape)
        elif self.op == "take":
            out = np.reshape(self.args[0].eval(env, memo), -1)[self.index]
        elif self.op == "matvec":
            out = self.args[0].eval(env, memo) @ self.args[1].eval(env, memo)
        elif self.op == "expquad":
            out = expquad_numpy(self.data[0], self.data[1], self.args[0].eval(env, memo), self.args[1].eval(env, memo))
        else:  # pragma: no cover - guarded by constructors
            raise NotImplementedError(self.op)
        out = np.asarray(out, dtype=np.float64)
        memo[key] = out
        return out

    def free_variables(self) -> set:
        seen, out, stack = set(), set(), [self]
        while stack:
            node = stack.pop()
            if id(node) in seen:
                continue
            seen.add(id(node))
            if node.op == "view":
                out.add(node.rv)
            stack.extend(node.args)
        return out


# ---------------------------------------------------------------------------
# constructors
# ---------------------------------------------------------------------------
def const(value) -> Expr:
    arr = np.asarray(value, dtype=np.float64)
    return Expr("const", shape=arr.shape, data=arr)


def free_rv(name: str, shape: Tuple[int, ...], space: str = "x") -> Expr:
    """The whole value of a sampled variable, as a view with the identity index."""
    size = int(np.prod(shape, dtype=np.int64)) if shape else 1
    idx = np.arange(size, dtype=np.int32).reshape(shape)
    return Expr("view", shape=shape, rv=name, space=space, index=idx, rv_size=size)


def is_symbolic(v: Any) -> bool:
    return isinstance(v, Expr) and v.op != "const"


def as_expr(v: ArrayLike) -> Expr:
    if isinstance(v, Expr):
        return v
    if hasattr(v, "detach") and hasattr(v, "cpu"):  # torch tensor
        v = v.detach().cpu().numpy()
    if hasattr(v, "values") and not isinstance(v, np.ndarray) and hasattr(v, "dtype"):
        v = np.asarray(v.values)  # pandas Series
    return const(v)


def as_numpy(v: ArrayLike) -> np.ndarray:
    """Concrete value of a constant expression / array-like."""
    e = as_expr(v)
    if e.op != "const":
        raise TypeError("expected a constant, got a symbolic expression {!r}".format(e))
    return e.data


def _maybe_fold(e: Expr) -> Expr:
    """Constant-fold nodes whose arguments are all constants."""
    if e.args and all(a.op == "const" for a in e.args):
        return const(e.eval({}))
    return e


def _unary(op: str, a: ArrayLike):
    if not isinstance(a, Expr):
        # plain numbers / arrays stay plain (transforms and model glue also run on concrete values)
        with np.errstate(all="ignore"):
            return _UNARY_NP[op](np.asarray(as_numpy(a), dtype=np.float64))
    return _maybe_fold(Expr(op, (a,), a.shape))


def _binary(op: str, a: ArrayLike, b: ArrayLike):
    try:
        a, b = as_expr(a), as_expr(b)
    except (TypeError, ValueError):
        return NotImplemented
    shape = np.broadcast_shapes(a.shape, b.shape)
    return _maybe_fold(Expr(op, (a, b), shape))


def _getitem(e: Expr, key) -> Expr:
    if e.op == "const":
        return const(e.data[key])
    if e.op == "view":
        idx = e.index[key]
        return Expr("view", shape=np.shape(idx), rv=e.rv, space=e.space,
                    index=np.asarray(idx, dtype=np.int32), rv_size=e.rv_size)
    # general expression: static take through a flat position map
    pos = np.arange(e.size, dtype=np.int32).reshape(e.shape)[key]
    return Expr("take", (e,), np.shape(pos), index=np.asarray(pos, dtype=np.int32))


# ---------------------------------------------------------------------------
# the ``pm.math`` namespace
# ---------------------------------------------------------------------------
def exp(x):
    return _unary("exp", x)


def log(x):
    return _unary("log", x)


def log1p(x):
    return _unary("log1p", x)


def sqrt(x):
    return _unary("sqrt", x)


def square(x):
    return _unary("square", x)


def sigmoid(x):
    return _unary("sigmoid", x)


def softplus(x):
    return _unary("softplus", x)


def tanh(x):
    return _unary("tanh", x)


def abs(x):  # noqa: A001 - mirrors tf.math.abs
    return _unary("abs", x)


def negative(x):
    return _unary("neg", x)


def reciprocal(x):
    return _unary("recip", x)


def lgamma(x):
    return _unary("lgamma", x)


def add(a, b):
    return _binary("add", a, b)


def subtract(a, b):
    return _binary("sub", a, b)


def multiply(a, b):
    return _binary("mul", a, b)


def divide(a, b):
    return _binary("div", a, b)


def pow(a, b):  # noqa: A001 - mirrors tf.math.pow
    return as_expr(a) ** b


def gather(params, indices, axis: int = 0) -> Expr:
    """``tf.gather(params, indices, axis)`` with constant integer ``indices``."""
    p = as_expr(params)
    idx = np.asarray(as_numpy(indices) if isinstance(indices, Expr) else indices)
    if not np.issubdtype(idx.dtype, np.integer):
        if np.any(idx != np.round(idx)):
            raise TypeError("gather indices must be integers")
        idx = idx.astype(np.int64)
    ax = axis % max(p.ndim, 1)
    key = (slice(None),) * ax + (idx,)
    return _getitem(p, key)


def broadcast_to(x, shape) -> Expr:
    x = as_expr(x)
    shape = tuple(int(s) for s in shape)
    if x.shape == shape:
        return x
    if x.op == "const":
        return const(np.broadcast_to(x.data, shape))
    if x.op == "view":
        return Expr("view", shape=shape, rv=x.rv, space=x.space,
                    index=np.ascontiguousarray(np.broadcast_to(x.index, shape)), rv_size=x.rv_size)
    np.broadcast_shapes(x.shape, shape)
    return Expr("broadcast", (x,), shape)


def reshape(x, shape) -> Expr:
    x = as_expr(x)
    if x.op == "const":
        return const(np.reshape(x.data, shape))
    if x.op == "view":
        idx = np.reshape(x.index, shape)
        return Expr("view", shape=idx.shape, rv=x.rv, space=x.space, index=idx, rv_size=x.rv_size)
    shape = np.reshape(np.empty(x.shape, dtype=np.int8), shape).shape
    return Expr("reshape", (x,), shape)


def expand_dims(x, axis: int) -> Expr:
    x = as_expr(x)
    shape = list(x.shape)
    ax = axis if axis >= 0 else axis + x.ndim + 1
    shape.insert(ax, 1)
    return reshape(x, shape)


def reduce_sum(x) -> Expr:
    """Sum over all elements (partial-axis reductions are not part of the IR)."""
    x = as_expr(x)
    return _maybe_fold(Expr("reduce_sum", (x,), ()))


def reduce_mean(x) -> Expr:
    x = as_expr(x)
    return _maybe_fold(Expr("reduce_mean", (x,), ()))


def matvec(a, b) -> Expr:
    """``tf.linalg.matvec`` with a constant matrix ``a`` [n, k] and a vector ``b`` [k]."""
    a, b = as_expr(a), as_expr(b)
    if a.ndim != 2 or b.ndim != 1 or a.shape[1] != b.shape[0]:
        raise ValueError("matvec expects [n, k] @ [k], got {} @ {}".format(a.shape, b.shape))
    if a.op != "const":
        raise NotImplementedError("matvec with a symbolic matrix is not part of the model IR")
    return _maybe_fold(Expr("matvec", (a, b), (a.shape[0],)))


def expquad_numpy(X1: np.ndarray, X2: np.ndarray, length_scale, amplitude) -> np.ndarray:
    """``amplitude^2 exp(-|x - x'|^2 / (2 length_scale^2))`` between the rows of X1 [n1, d] and X2 [n2, d]
    (tfp ExponentiatedQuadratic.matrix, which /root/reference/pymc4/gp/cov.py:463-469 wraps)."""
    X1, X2 = np.asarray(X1, dtype=np.float64), np.asarray(X2, dtype=np.float64)
    d2 = np.sum(np.square(X1[:, None, :] - X2[None, :, :]), axis=-1)
    return np.square(amplitude) * np.exp(-d2 / (2.0 * np.square(length_scale)))


def expquad(X1, X2, length_scale, amplitude) -> Expr:
    """Covariance-matrix node: constant points, scalar (possibly free) length scale and amplitude."""
    X1, X2 = as_numpy(X1), as_numpy(X2)
    ls, amp = as_expr(length_scale), as_expr(amplitude)
    if ls.size != 1 or amp.size != 1:
        raise NotImplementedError("ExpQuad on the device takes a scalar length_scale and amplitude")
    return _maybe_fold(Expr("expquad", (reshape(ls, ()), reshape(amp, ())), (X1.shape[0], X2.shape[0]), data=(X1, X2)))


def zeros(shape):
    return np.zeros(shape, dtype=np.float64)


def ones(shape):
    return np.ones(shape, dtype=np.float64)


def shape_of(v) -> Tuple[int, ...]:
    return as_expr(v).shape
