"""Name scoping for model variables.

Behavioural contract follows the reference (pymc4/scopes.py:6-222): a
thread-local stack of scopes, each of which may carry a ``name``; a variable's
full name is the "/"-joined chain of non-None scope names plus the leaf name
(pymc4/scopes.py:143-178) and an auto-transformed variable is spelled
``__<transform>_<name>`` (pymc4/scopes.py:180-212).  The trace keys the sampler
returns (``model/__log_tau`` ...) depend on this contract.
"""
import threading
from typing import Any, Callable, Iterator, List, Optional

_NO_LEAF = object()
_tls = threading.local()


def _stack() -> List["Scope"]:
    st = getattr(_tls, "stack", None)
    if st is None:
        st = _tls.stack = []
    return st


class Scope:
    """A context manager carrying arbitrary keyword attributes.

    Unknown attributes read as ``None`` so callers can probe any key.
    """

    def __init__(self, **attrs: Any):
        self.__dict__.update(attrs)

    def __getattr__(self, key: str) -> Any:  # only called when normal lookup fails
        return None

    def __enter__(self) -> "Scope":
        _stack().append(self)
        return self

    def __exit__(self, *exc) -> None:
        _stack().pop()

    def __repr__(self) -> str:
        return "Scope({})".format(self.__dict__)

    # -- queries over the active stack -------------------------------------
    @classmethod
    def get_contexts(cls) -> List["Scope"]:
        return _stack()

    @classmethod
    def chain(
        cls,
        attr: str,
        *,
        predicate: Callable[["Scope"], bool] = lambda _s: True,
        drop_none: bool = False,
        leaf: Any = _NO_LEAF,
    ) -> Iterator[Any]:
        """Values of ``attr`` from the outermost to the innermost scope, then ``leaf``."""
        for scope in _stack():
            if not predicate(scope):
                continue
            val = getattr(scope, attr)
            if val is None and drop_none:
                continue
            yield val
        if leaf is not _NO_LEAF and not (drop_none and leaf is None):
            yield leaf

    @classmethod
    def variable_name(cls, name: Optional[str]) -> Optional[str]:
        parts = [str(p) for p in cls.chain("name", leaf=name, drop_none=True)]
        joined = "/".join(parts)
        return joined or None

    @classmethod
    def transformed_variable_name(cls, transform_name: str, name: str) -> Optional[str]:
        return cls.variable_name("__{}_{}".format(transform_name, name))


def name_scope(name: Optional[str]) -> Scope:
    return Scope(name=name)


variable_name = Scope.variable_name
transformed_variable_name = Scope.transformed_variable_name
