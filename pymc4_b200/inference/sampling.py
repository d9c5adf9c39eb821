"""``pm.sample``: the user-facing MCMC entry point.

Signature, sampler selection and the ``num_adaptation_steps = burn_in`` rule are the reference's
(pymc4/inference/sampling.py:52-184, :187-214).  Samplers: ``"nuts"`` (default for models without free
discrete variables), ``"hmc"``, ``"nuts_simple"``, ``"hmc_simple"``, ``"rwm"`` and ``"compound"`` (default when the
model has free discrete variables).
"""
from typing import Any, Dict, List, Optional

from .. import flow
from ..coroutine_model import Model
from ..mcmc.samplers import _log, reg_samplers
from ..mcmc.utils import initialize_state, scope_remove_transformed_part_if_required


def check_proposal_functions(model: Model, state: Optional[flow.SamplingState] = None,
                             observed: Optional[dict] = None) -> bool:
    """True when a free variable's distribution carries a non-default proposal function (reference :14-49)."""
    (_, state, _, _, continuous_distrs, discrete_distrs) = initialize_state(model, observed=observed, state=state)
    for key in state.all_unobserved_values:
        untrs_var, _ = scope_remove_transformed_part_if_required(key, state.transformed_values)
        distr = continuous_distrs.get(untrs_var, None)
        if distr is None:
            distr = discrete_distrs[untrs_var]
        if callable(distr._default_new_state_part):
            return True
    return False


def sample(
    model: Model,
    sampler_type: Optional[str] = None,
    num_samples: int = 1000,
    num_chains: int = 10,
    burn_in: int = 100,
    observed: Optional[Dict[str, Any]] = None,
    state: Optional[flow.SamplingState] = None,
    xla: bool = False,
    use_auto_batching: bool = True,
    sampler_methods: Optional[List] = None,
    trace_discrete: Optional[List[str]] = None,
    seed: Optional[int] = None,
    include_log_likelihood: bool = False,
    **kwargs,
):
    """Draw posterior samples with NUTS (or HMC) on the GPU.

    Parameters are those of the reference ``pm.sample``; ``**kwargs`` are routed to the kernel
    (``step_size``, ``max_tree_depth``, ``max_energy_diff``, ``num_leapfrog_steps`` ...), to the
    dual-averaging adaptation (``target_accept_prob`` ...) and to the chain driver by name.
    Backend-specific extras: ``dtype`` ("float32" | "float64"), ``device`` (CUDA ordinal),
    ``chain_offset`` (global index of the first chain when chains are sharded over processes).
    ``xla`` is accepted and ignored.

    Returns an ``InferenceData``-like object with groups ``posterior``, ``sample_stats`` and
    ``observed_data`` holding arrays laid out ``[chain, draw, *core]``.
    """
    sampler_assigned: str = auto_assign_sampler(model, sampler_type)
    try:
        sampler = reg_samplers[sampler_assigned]
    except KeyError:
        _log.warning(
            "The given sampler doesn't exist. Please choose samplers from: {}".format(list(reg_samplers.keys()))
        )
        raise
    if any(x in sampler_assigned for x in ["nuts", "hmc"]):
        kwargs["num_adaptation_steps"] = burn_in
    sampler = sampler(model, **kwargs)
    # distributions with their own proposal function (Bernoulli, Categorical) turn "rwm" into the compound step
    if sampler_assigned == "rwm":
        if check_proposal_functions(model, state=state, observed=observed):
            sampler_assigned = "compound"
            sampler = reg_samplers[sampler_assigned](model, **kwargs)
    if sampler_assigned == "compound":
        sampler._assign_default_methods(sampler_methods=sampler_methods, state=state, observed=observed)
    return sampler(
        num_samples=num_samples,
        num_chains=num_chains,
        burn_in=burn_in,
        observed=observed,
        state=state,
        use_auto_batching=use_auto_batching,
        xla=xla,
        seed=seed,
        trace_discrete=trace_discrete,
        include_log_likelihood=include_log_likelihood,
    )


def auto_assign_sampler(model: Model, sampler_type: Optional[str] = None) -> str:
    """NUTS for models whose free variables are all continuous (reference :187-214)."""
    if sampler_type:
        _log.info("Working with {} sampler".format(reg_samplers[sampler_type].__name__))
        return sampler_type
    _, _, free_disc_names, free_cont_names, _, _ = initialize_state(model)
    if not free_disc_names:
        _log.info("Auto-assigning NUTS sampler")
        return "nuts"
    _log.info("The model contains discrete distributions. " "\nCompound step is chosen.")
    return "compound"
