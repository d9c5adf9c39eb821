"""Small host-side helpers: the variable-name grammar and the optional-kwargs decorator.

Name grammar follows the reference (pymc4/utils.py:44-212, regex at :59):
``path/to/__<transform>_<name>`` or ``path/to/<name>``; a plain name may not
start with an underscore, and a transform tag contains no underscore.
"""
import functools
import re
from typing import Callable, Optional, Sequence, Tuple

_LEAF_RE = re.compile(r"^(?:__(?P<transform>[^_]+)_)?(?P<name>[^_].*)$")


class NameParts:
    """A scoped variable name split into (path, transform tag, plain name)."""

    NAME_ERROR_MESSAGE = (
        "Invalid name: `{}`, the correct one should look like: `__transform_name` or `name`, "
        "note only one underscore between the transform and actual name"
    )
    UNTRANSFORMED_NAME_ERROR_MESSAGE = (
        "Invalid name: `{}`, the correct one should look like: `name` without leading underscore"
    )
    __slots__ = ("path", "transform_name", "untransformed_name")

    def __init__(self, path: Sequence[str], transform_name: Optional[str], untransformed_name: str):
        self.path: Tuple[str, ...] = tuple(path)
        self.transform_name = transform_name
        self.untransformed_name = untransformed_name

    # -- validation --------------------------------------------------------
    @staticmethod
    def is_valid_name(name: str) -> bool:
        return _LEAF_RE.match(name) is not None

    @staticmethod
    def is_valid_untransformed_name(name: str) -> bool:
        m = _LEAF_RE.match(name)
        return m is not None and m.group("transform") is None

    # -- construction ------------------------------------------------------
    @classmethod
    def from_name(cls, name: str) -> "NameParts":
        *path, leaf = name.split("/")
        if not cls.is_valid_name(name):
            raise ValueError(cls.NAME_ERROR_MESSAGE.format(name))
        m = _LEAF_RE.match(leaf)
        if m is None:
            raise ValueError(cls.NAME_ERROR_MESSAGE.format(name))
        return cls(path, m.group("transform"), m.group("name"))

    def replace_transform(self, transform_name: Optional[str]) -> "NameParts":
        return type(self)(self.path, transform_name, self.untransformed_name)

    # -- views -------------------------------------------------------------
    @property
    def is_transformed(self) -> bool:
        return self.transform_name is not None

    @property
    def original_name(self) -> str:
        if self.is_transformed:
            return "__{}_{}".format(self.transform_name, self.untransformed_name)
        return self.untransformed_name

    @property
    def full_original_name(self) -> str:
        return "/".join(self.path + (self.original_name,))

    @property
    def full_untransformed_name(self) -> str:
        return "/".join(self.path + (self.untransformed_name,))

    def __repr__(self) -> str:
        return "<NameParts of {}>".format(self.full_original_name)


def biwrap(wrapper: Callable) -> Callable:
    """Let a decorator be used both bare (``@deco``) and with kwargs (``@deco(k=v)``).

    Same contract as the reference helper (pymc4/utils.py:10-41), including the
    bound-method case where the first positional argument is ``self``.
    """

    @functools.wraps(wrapper)
    def flexible(*args, **kwargs):
        n_self = 1 if (args and hasattr(args[0], wrapper.__name__)) else 0
        if len(args) > n_self:
            return wrapper(*args, **kwargs)
        return functools.partial(wrapper, *args, **kwargs)

    return flexible
