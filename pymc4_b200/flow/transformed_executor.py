"""Evaluate a model in the unconstrained space.

Same contract as the reference (pymc4/flow/transformed_executor.py:22-138): a
distribution with a transform that is not observed is replaced by a tiny
sub-model which (a) maps between the constrained value ``name`` and the
unconstrained value ``__<transform>_name`` and (b) adds the log-det-Jacobian as a
``Potential`` with ``coef=-1`` applied to ``forward_log_det_jacobian(x)``.
"""
import functools
from typing import Any, Mapping

from .. import scopes
from ..distributions import distribution
from ..distributions.transforms import JacobianPreference
from .executor import (
    EvaluationError,
    SamplingExecutor,
    SamplingState,
    observed_value_in_evaluation,
)


def _logdet_potential(transform, x, z):
    if transform.jacobian_preference == JacobianPreference.Forward:
        return distribution.Potential(
            functools.partial(transform.forward_log_det_jacobian, x), coef=-1.0
        )
    return distribution.Potential(
        functools.partial(transform.inverse_log_det_jacobian, z), coef=1.0
    )


def make_untransformed_model(dist, transform, state: SamplingState):
    """No unconstrained value supplied: obtain x from the executor, derive z from it."""
    dist.model_info["autotransformed"] = True
    x = yield dist
    z = transform.forward(x)
    tname = scopes.transformed_variable_name(transform.name, dist.name)
    state.transformed_values[tname] = z
    yield _logdet_potential(transform, x, z)
    return x


def make_transformed_model(dist, transform, state: SamplingState):
    """An unconstrained value z is in the state: x = inverse(z) is what the model sees."""
    name = scopes.variable_name(dist.name)
    tname = scopes.transformed_variable_name(transform.name, dist.name)
    z = state.transformed_values[tname]
    x = state.untransformed_values[name] = transform.inverse(z)
    dist.model_info["autotransformed"] = True
    yield _logdet_potential(transform, x, z)
    return (yield dist)


def transform_dist_if_necessary(dist, state: SamplingState, *, allow_transformed_and_untransformed):
    if dist.transform is None or dist.model_info.get("autotransformed", False):
        return dist
    name = scopes.variable_name(dist.name)
    transform = dist.transform
    tname = scopes.transformed_variable_name(transform.name, dist.name)
    if observed_value_in_evaluation(name, dist, state) is not None:
        # observed variables are never transformed; a competing free value is an error
        if tname in state.transformed_values:
            raise EvaluationError(
                EvaluationError.OBSERVED_VARIABLE_IS_NOT_SUPPRESSED_BUT_ADDITIONAL_TRANSFORMED_VALUE_PASSED.format(
                    name, tname
                )
            )
        if name in state.untransformed_values:
            raise EvaluationError(
                EvaluationError.OBSERVED_VARIABLE_IS_NOT_SUPPRESSED_BUT_ADDITIONAL_VALUE_PASSED.format(
                    name, name
                )
            )
        return dist
    if tname in state.transformed_values:
        if not allow_transformed_and_untransformed and name in state.untransformed_values:
            state.untransformed_values.pop(name)
        return make_transformed_model(dist, transform, state)
    return make_untransformed_model(dist, transform, state)


class TransformedSamplingExecutor(SamplingExecutor):
    """Perform inference in an unconstrained space."""

    def validate_state(self, state):
        return

    def modify_distribution(self, dist, model_info: Mapping[str, Any], state: SamplingState):
        dist = super().modify_distribution(dist, model_info, state)
        if not isinstance(dist, distribution.Distribution):
            return dist
        return transform_dist_if_necessary(dist, state, allow_transformed_and_untransformed=True)
