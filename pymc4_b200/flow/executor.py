"""Generator-driving model interpreters and the ``SamplingState`` they fill.

Behavioural contract follows the reference interpreter
(pymc4/flow/executor.py:93-327 ``SamplingState``, :355-560 ``evaluate_model``,
:581-677 ``proceed_distribution`` / ``proceed_deterministic``, :679-717 control
flow set-up/tear-down, :720-819 observed handling and shape checks).  Values
flowing through a model are either concrete NumPy arrays or symbolic
expressions (``pymc4_b200.symbolic``); the samplers run the transformed executor
once with symbolic values and lower the resulting state to the flat model IR.
"""
import itertools
import types
from collections import ChainMap
from typing import Any, Dict, List, Mapping, Optional, Set, Tuple, Union

import numpy as np

from .. import coroutine_model, scopes, symbolic as sym, utils
from ..distributions import distribution

ModelType = Union[types.GeneratorType, coroutine_model.Model]
MODEL_TYPES = (types.GeneratorType, coroutine_model.Model)
MODEL_POTENTIAL_AND_DETERMINISTIC_TYPES = MODEL_TYPES + (
    distribution.Potential,
    distribution.Deterministic,
)


class EvaluationError(RuntimeError):
    OBSERVED_VARIABLE_IS_NOT_SUPPRESSED_BUT_ADDITIONAL_VALUE_PASSED = (
        "Attempting to evaluate a model with both observed and unobserved values provided "
        "what requires to choose what value to actually yield. To remove this error you "
        "either need to add `observed={{{0!r}: None}}` or remove {0!r} from untransformed values."
    )
    OBSERVED_VARIABLE_IS_NOT_SUPPRESSED_BUT_ADDITIONAL_TRANSFORMED_VALUE_PASSED = (
        "Attempting to evaluate a model with both observed and unobserved values provided "
        "what requires to choose what value to actually yield. To remove this error you "
        "either need to add `observed={{{0!r}: None}}` or remove {1!r} from transformed values."
    )
    INCOMPATIBLE_VALUE_AND_DISTRIBUTION_SHAPE = (
        "The values supplied to the distribution {0!r} are not consistent with the "
        "distribution's shape (dist_shape).\ndist_shape = {1}\nSupplied values shape = {2}.\n"
        "A values array is considered to have a consistent shape with the distribution if "
        "its rightmost axes match the distribution's shape (batch_shape + event_shape) and "
        "it has at least as many axes as the event_shape."
    )


class StopExecution(StopIteration):
    NOT_HELD_ERROR_MESSAGE = (
        "the model's control flow swallowed an error injected at a `yield`; evaluation cannot "
        "continue in a well-defined state"
    )


class EarlyReturn(StopIteration):
    ...


def _ordered_chainmap(*maps) -> ChainMap:
    return ChainMap(*maps)


def _cm_iter(self):
    # iterate a ChainMap in insertion order of the *last* map first (py3.6 ordering quirk the
    # reference relies on, executor.py:24-35): untransformed values come before transformed.
    merged = {}
    for mapping in reversed(self.maps):
        merged.update(mapping)
    return iter(merged)


class _OrderedChainMap(ChainMap):
    __iter__ = _cm_iter


class SamplingState:
    """Everything an executor learns (or is told) about one pass through a model."""

    __slots__ = (
        "transformed_values",
        "untransformed_values",
        "observed_values",
        "posterior_predictives",
        "all_values",
        "all_unobserved_values",
        "discrete_distributions",
        "continuous_distributions",
        "potentials",
        "deterministics",
        "deterministics_values",
    )

    def __init__(
        self,
        transformed_values: Dict[str, Any] = None,
        untransformed_values: Dict[str, Any] = None,
        observed_values: Dict[str, Any] = None,
        discrete_distributions: Dict[str, distribution.Distribution] = None,
        continuous_distributions: Dict[str, distribution.Distribution] = None,
        potentials: List[distribution.Potential] = None,
        deterministics: Dict[str, distribution.Deterministic] = None,
        posterior_predictives: Optional[Set[str]] = None,
        deterministics_values: Dict[str, Any] = None,
    ) -> None:
        self.transformed_values = dict(transformed_values or {})
        self.untransformed_values = dict(untransformed_values or {})
        self.observed_values = dict(observed_values or {})
        self.discrete_distributions = dict(discrete_distributions or {})
        self.continuous_distributions = dict(continuous_distributions or {})
        self.potentials = list(potentials or [])
        self.deterministics = dict(deterministics or {})
        self.posterior_predictives = set(posterior_predictives or ())
        self.deterministics_values = dict(deterministics_values or {})
        self.all_values = _OrderedChainMap(
            self.untransformed_values, self.transformed_values, self.observed_values
        )
        self.all_unobserved_values = _OrderedChainMap(
            self.transformed_values, self.untransformed_values
        )

    # -- joint density (host-side NumPy; concrete values only) -------------
    def collect_log_prob_elemwise(self):
        dists = itertools.chain(
            self.discrete_distributions.items(), self.continuous_distributions.items()
        )
        return itertools.chain(
            (dist.log_prob(self.all_values[name]) for name, dist in dists),
            (sym.as_numpy(p.value) for p in self.potentials),
        )

    def collect_log_prob(self):
        return sum(np.sum(t) for t in self.collect_log_prob_elemwise())

    def collect_unreduced_log_prob(self):
        return sum(self.collect_log_prob_elemwise())

    def __repr__(self):
        def tagged(dists):
            return ["{}:{}".format(type(d).__name__, k) for k, d in dists.items()]

        rows = [
            ("untransformed_values", list(self.untransformed_values)),
            ("transformed_values", list(self.transformed_values)),
            ("observed_values", list(self.observed_values)),
            ("discrete_distributions", tagged(self.discrete_distributions)),
            ("continuous_distributions", tagged(self.continuous_distributions)),
            ("num_potentials", len(self.potentials)),
            ("deterministics", list(self.deterministics)),
            ("deterministics_values", list(self.deterministics_values)),
            ("posterior_predictives", list(self.posterior_predictives)),
        ]
        body = "\n".join("    {}: {}".format(k, v) for k, v in rows)
        return "{}(\n{})".format(type(self).__name__, body)

    @classmethod
    def from_values(
        cls, values: Dict[str, Any] = None, observed_values: Dict[str, Any] = None
    ) -> "SamplingState":
        transformed, untransformed = {}, {}
        for fullname, value in (values or {}).items():
            if utils.NameParts.from_name(fullname).is_transformed:
                transformed[fullname] = value
            else:
                untransformed[fullname] = value
        return cls(transformed, untransformed, observed_values)

    def clone(self) -> "SamplingState":
        return type(self)(**{k: getattr(self, k) for k in (
            "transformed_values", "untransformed_values", "observed_values",
            "discrete_distributions", "continuous_distributions", "potentials",
            "deterministics", "posterior_predictives", "deterministics_values")})

    def as_sampling_state(self) -> "Tuple[SamplingState, List[str]]":
        """Keep only what MCMC carries: transformed values where a transform exists,
        untransformed values otherwise, and the observed values (reference :272-327)."""
        if not self.discrete_distributions and not self.continuous_distributions:
            raise TypeError(
                "No distributions found in the state. the model you evaluated is empty and "
                "does not yield any PyMC4 distribution"
            )
        untransformed, transformed, observed = {}, {}, {}
        back_transform_names: List[str] = []
        for name, dist in itertools.chain(
            self.discrete_distributions.items(), self.continuous_distributions.items()
        ):
            if dist.transform is not None and name not in self.observed_values:
                tparts = utils.NameParts.from_name(name).replace_transform(dist.transform.name)
                tname = tparts.full_original_name
                if tname not in self.transformed_values:
                    raise TypeError(
                        "Transformed value {!r} is not found for {} distribution with name {!r}. "
                        "You should evaluate the model using the transformed executor to get "
                        "the correct sampling state.".format(tname, dist, name)
                    )
                transformed[tname] = self.transformed_values[tname]
                back_transform_names.append(tparts.full_untransformed_name)
            elif name in self.observed_values:
                observed[name] = self.observed_values[name]
            elif name in self.untransformed_values:
                untransformed[name] = self.untransformed_values[name]
            else:
                raise TypeError(
                    "{} distribution with name {!r} does not have the corresponding value in the "
                    "state. This may happen if the current state was modified in the wrong "
                    "way.".format(dist, name)
                )
        return (
            type(self)(
                transformed_values=transformed,
                untransformed_values=untransformed,
                observed_values=observed,
            ),
            back_transform_names,
        )


def _map_structure(fn, obj):
    if isinstance(obj, (list, tuple)):
        for o in obj:
            _map_structure(fn, o)
    elif isinstance(obj, dict):
        for o in obj.values():
            _map_structure(fn, o)
    else:
        fn(obj)


class SamplingExecutor:
    """Evaluate a model in the untransformed (constrained) space."""

    def validate_return_object(self, return_object: Any):
        if isinstance(return_object, MODEL_POTENTIAL_AND_DETERMINISTIC_TYPES):
            raise EvaluationError(
                "Return values should not contain instances of `pm.coroutine_model.Model`, "
                "`types.GeneratorType` and `pm.distributions.Potential`. To fix the error you "
                "should change the return statement to something like\n    ...\n    return (yield variable)"
            )

    def validate_return_value(self, return_value: Any):
        _map_structure(self.validate_return_object, return_value)

    def new_state(self, values=None, observed=None) -> SamplingState:
        return SamplingState.from_values(values=values, observed_values=observed)

    def validate_state(self, state: SamplingState):
        if state.transformed_values:
            raise ValueError(
                "untransformed executor should not contain transformed variables but found "
                "{}".format(set(state.transformed_values))
            )

    def modify_distribution(self, dist, model_info: Mapping[str, Any], state: SamplingState):
        return dist

    # ------------------------------------------------------------------
    def evaluate_model(
        self,
        model: ModelType,
        *,
        state: SamplingState = None,
        _validate_state: bool = True,
        values: Dict[str, Any] = None,
        observed: Dict[str, Any] = None,
        sample_shape=(),
    ) -> Tuple[Any, SamplingState]:
        if state is None:
            state = self.new_state(values=values, observed=observed)
        elif values or observed:
            raise ValueError("Provided arguments along with not empty state")
        if _validate_state:
            self.validate_state(state)

        if isinstance(model, distribution.Distribution):
            # a bare distribution: wrap it so its name is not doubled ("n/n")
            bare = model
            model = (lambda: (yield bare))()
        if isinstance(model, coroutine_model.Model):
            model_info = model.model_info
            try:
                control_flow = self.prepare_model_control_flow(model, model_info, state)
            except EarlyReturn as early:
                return early.args, state
        elif isinstance(model, types.GeneratorType):
            control_flow, model_info = model, coroutine_model.Model.default_model_info
        else:
            raise StopExecution(
                "Attempting to call `evaluate_model` on a non model-like object {}. Supported "
                "types are `types.GeneratorType` and `pm.coroutine_model.Model`".format(type(model))
            )

        def fail(error: Exception):
            """Inject ``error`` at the yield; if the model swallows it, stop anyway."""
            control_flow.throw(error)
            raise StopExecution(StopExecution.NOT_HELD_ERROR_MESSAGE) from error

        send_back = None
        while True:
            try:
                with model_info["scope"]:
                    node = control_flow.send(send_back)
                    if not isinstance(node, MODEL_POTENTIAL_AND_DETERMINISTIC_TYPES):
                        fail(EvaluationError(
                            "Type {} can't be processed in evaluation".format(type(node))))
                    if isinstance(node, distribution.Distribution):
                        node = self.modify_distribution(node, model_info, state)
                    if isinstance(node, distribution.Potential):
                        state.potentials.append(node)
                        send_back = node
                    elif isinstance(node, distribution.Deterministic):
                        try:
                            send_back, state = self.proceed_deterministic(node, state)
                        except EvaluationError as error:
                            fail(error)
                    elif isinstance(node, distribution.Distribution):
                        try:
                            send_back, state = self.proceed_distribution(
                                node, state, sample_shape=sample_shape
                            )
                        except EvaluationError as error:
                            fail(error)
                    elif isinstance(node, MODEL_TYPES):
                        send_back, state = self.evaluate_model(
                            node, state=state, _validate_state=False, sample_shape=sample_shape
                        )
                    else:  # pragma: no cover - unreachable by construction
                        fail(EvaluationError(
                            "Type {} can't be processed in evaluation.".format(type(node))))
            except StopExecution:
                control_flow.close()
                raise
            except EarlyReturn as early:
                return early.args, state
            except StopIteration as stop:
                self.validate_return_value(stop.args[:1])
                return self.finalize_control_flow(stop, model_info, state)

    __call__ = evaluate_model

    # ------------------------------------------------------------------
    def _draw(self, dist, sample_shape):
        return dist.sample(sample_shape=sample_shape) if dist.is_root else dist.sample()

    def _check_fresh_name(self, scoped_name, state, what):
        if (
            scoped_name in state.discrete_distributions
            or scoped_name in state.continuous_distributions
            or scoped_name in state.deterministics_values
        ):
            raise EvaluationError(
                "Attempting to create a duplicate {} {!r}, this may happen if you forget to use "
                "`pm.name_scope()` when calling same model/function twice without providing "
                "explicit names. If you see this error message and the function being called is "
                "not wrapped with `pm.model`, you should better wrap it to provide explicit name "
                "for this model".format(what, scoped_name)
            )

    def proceed_distribution(self, dist, state: SamplingState, sample_shape=None):
        if dist.is_anonymous:
            raise EvaluationError("Attempting to create an anonymous Distribution")
        scoped_name = scopes.variable_name(dist.name)
        if scoped_name is None:
            raise EvaluationError("Attempting to create an anonymous Distribution")
        self._check_fresh_name(scoped_name, state, "variable")

        if scoped_name in state.observed_values or dist.is_observed:
            observed_variable = observed_value_in_evaluation(scoped_name, dist, state)
            if observed_variable is None:
                # observation explicitly suppressed: predictive draw or user-supplied latent
                if scoped_name not in state.untransformed_values:
                    state.untransformed_values[scoped_name] = self._draw(dist, sample_shape)
                value = state.untransformed_values[scoped_name]
                state.posterior_predictives.add(scoped_name)
                state.observed_values.pop(scoped_name)
            else:
                if scoped_name in state.untransformed_values:
                    raise EvaluationError(
                        EvaluationError.OBSERVED_VARIABLE_IS_NOT_SUPPRESSED_BUT_ADDITIONAL_VALUE_PASSED.format(
                            scoped_name
                        )
                    )
                assert_values_compatible_with_distribution(scoped_name, observed_variable, dist)
                value = state.observed_values[scoped_name] = observed_variable
        elif scoped_name in state.untransformed_values:
            value = state.untransformed_values[scoped_name]
        else:
            value = state.untransformed_values[scoped_name] = self._draw(dist, sample_shape)

        registry = state.continuous_distributions if dist._grad_support else state.discrete_distributions
        registry[scoped_name] = dist
        return value, state

    def proceed_deterministic(self, deterministic, state: SamplingState):
        if deterministic.is_anonymous:
            raise EvaluationError("Attempting to create an anonymous Deterministic")
        scoped_name = scopes.variable_name(deterministic.name)
        if scoped_name is None:
            raise EvaluationError("Attempting to create an anonymous Deterministic")
        self._check_fresh_name(scoped_name, state, "deterministic")
        value = state.deterministics_values[scoped_name] = deterministic.get_value()
        state.deterministics[scoped_name] = deterministic
        return value, state

    def prepare_model_control_flow(self, model, model_info, state: SamplingState):
        control_flow = model.control_flow()
        model_name = model_info["name"]
        if model_name is None and model_info["keep_return"]:
            error = EvaluationError(
                "Attempting to create unnamed return variable when `keep_return` is set to True"
            )
            control_flow.throw(error)
            control_flow.close()
            raise StopExecution(StopExecution.NOT_HELD_ERROR_MESSAGE) from error
        return_name = scopes.variable_name(model_name)
        if not model_info["keep_auxiliary"] and return_name in state.untransformed_values:
            raise EarlyReturn(state.untransformed_values[model_name], state)
        return control_flow

    def finalize_control_flow(self, stop_iteration, model_info, state: SamplingState):
        return_value = stop_iteration.args[0] if stop_iteration.args else None
        if return_value is not None and model_info["keep_return"]:
            return_name = scopes.variable_name(model_info["name"])
            if return_name is None:
                raise AssertionError(
                    "Attempting to create unnamed return variable *after* making a check"
                )
            state.deterministics_values[return_name] = return_value
        return return_value, state


def observed_value_in_evaluation(scoped_name: str, dist, state: SamplingState):
    return state.observed_values.get(scoped_name, dist.model_info["observed"])


def get_observed_tensor_shape(arr: Any) -> Tuple[int, ...]:
    return sym.as_expr(arr).shape


def _compatible(a: Tuple[int, ...], b: Tuple[int, ...]) -> bool:
    return len(a) == len(b) and all(x == y for x, y in zip(a, b))


def assert_values_compatible_with_distribution_shape(
    scoped_name: str, values: Any, batch_shape: Tuple[int, ...], event_shape: Tuple[int, ...]
) -> None:
    """Right-aligned shape check of a supplied value against batch+event shape (reference :762-819)."""
    value_shape = tuple(get_observed_tensor_shape(values))
    dist_shape = tuple(batch_shape) + tuple(event_shape)
    vr, dr = len(value_shape), len(dist_shape)
    bad = (
        vr < len(event_shape)
        or (dr < vr and not _compatible(dist_shape, value_shape[vr - dr :]))
        or (vr < dr and not _compatible(value_shape, dist_shape[dr - vr :]))
        or (vr == dr and not _compatible(value_shape, dist_shape))
    )
    if bad:
        raise EvaluationError(
            EvaluationError.INCOMPATIBLE_VALUE_AND_DISTRIBUTION_SHAPE.format(
                scoped_name, dist_shape, value_shape
            )
        )


def assert_values_compatible_with_distribution(scoped_name: str, values: Any, dist) -> None:
    assert_values_compatible_with_distribution_shape(
        scoped_name, values, dist.batch_shape, dist.event_shape
    )
