"""Coroutine (generator) models: ``@pm.model`` and the ``Model`` object it produces.

Keeps the user-facing contract of the reference (pymc4/coroutine_model.py:14-217):
a decorated generator function becomes a template; calling the template binds
the conditioners and yields a ``Model`` whose ``control_flow()`` re-creates the
generator; ``model_info`` carries ``keep_auxiliary``, ``keep_return``, ``scope``
and ``name`` for the executors.
"""
import functools
import types
from typing import Optional, Union

from .scopes import name_scope
from .utils import NameParts, biwrap

# distinguishes "no name argument given" from an explicit ``name=None``
_UNSET = object()


def _resolve_name(template_default, fn, call_name):
    if call_name is not _UNSET:
        return call_name
    if template_default is not _UNSET:
        return template_default
    for attr in ("name", "__name__"):
        if hasattr(fn, attr):
            return getattr(fn, attr)
    return call_name


class Model:
    """A conditioned generative process; iterate it through ``control_flow()``."""

    default_model_info = types.MappingProxyType(
        dict(keep_auxiliary=True, keep_return=False, scope=name_scope(None), name=None)
    )

    @staticmethod
    def validate_name(name: Optional[Union[int, str]]) -> Optional[str]:
        if name is None:
            return None
        if not isinstance(name, (int, str)):
            raise ValueError("name should be either `str` or `int`, got type {}".format(type(name)))
        return str(name)

    def __init__(self, genfn, *, name=None, keep_auxiliary=True, keep_return=True):
        self.genfn = genfn
        self.name = self.validate_name(name)
        self.model_info = dict(
            keep_auxiliary=keep_auxiliary,
            keep_return=keep_return,
            scope=name_scope(self.name),
            name=self.name,
        )

    def control_flow(self):
        return (yield from self.genfn())


class ModelTemplate:
    """A generator function plus default metadata; call it to obtain a ``Model``."""

    def __init__(self, template, *, name=None, keep_auxiliary=True, keep_return=True):
        self.template = template
        self.name = name
        self.keep_auxiliary = keep_auxiliary
        self.keep_return = keep_return

    def __call__(self, *args, name=_UNSET, keep_auxiliary=None, keep_return=None, **kwargs):
        bound = functools.partial(self.template, *args, **kwargs)
        final_name = _resolve_name(self.name, self.template, name)
        if final_name is not None and not NameParts.is_valid_untransformed_name(str(final_name)):
            raise ValueError(NameParts.UNTRANSFORMED_NAME_ERROR_MESSAGE.format(final_name))
        return Model(
            bound,
            name=final_name,
            keep_auxiliary=self.keep_auxiliary if keep_auxiliary is None else keep_auxiliary,
            keep_return=self.keep_return if keep_return is None else keep_return,
        )


@biwrap
def model(genfn, *, name=_UNSET, keep_auxiliary=True, keep_return=True, method=False):
    """Turn a generator function into a model template (usable bare or with kwargs)."""
    template = ModelTemplate(genfn, name=name, keep_auxiliary=keep_auxiliary, keep_return=keep_return)
    if not method:
        return template

    @functools.wraps(genfn)
    def as_method(*args, **kwargs):
        return template(*args, **kwargs)

    return as_method


def unpack(arg):
    """Generator helper: yield model-like objects through, return plain values as is."""
    if isinstance(arg, (Model, types.GeneratorType)):
        return (yield arg)
    return arg
