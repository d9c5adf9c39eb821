"""Sampler-side helpers: initial state discovery and trace assembly.

Contracts follow the reference (pymc4/mcmc/utils.py): ``initialize_sampling_state``
(:14-36) runs the meta executor and keeps the MCMC-relevant values;
``initialize_state`` (:39-92) lists free discrete / continuous variable names;
``trace_to_arviz`` (:95-144) swaps ``[draw, chain]`` to ``[chain, draw]``, drops
keys without a "/" and builds the ``posterior`` / ``sample_stats`` /
``observed_data`` (/ ``log_likelihood``) groups.  ArviZ is not part of this
image, so the groups are returned in a small ``InferenceData`` stand-in with the
same attribute access; when ``arviz`` is importable ``to_arviz()`` hands the
same dictionaries to ``az.from_dict``.
"""
import collections
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .. import flow
from ..coroutine_model import Model

KERNEL_KWARGS_SET = collections.namedtuple(
    "KERNEL_ARGS_SET", ["kernel", "adaptive_kernel", "kernel_kwargs", "adaptive_kwargs"]
)


def initialize_sampling_state(
    model: Model, observed: Optional[dict] = None, state: Optional[flow.SamplingState] = None
) -> Tuple[flow.SamplingState, List[str]]:
    _, state = flow.evaluate_meta_model(model, observed=observed, state=state)
    deterministic_names = list(state.deterministics_values)
    state, transformed_names = state.as_sampling_state()
    return state, deterministic_names + transformed_names


def initialize_state(
    model: Model, observed: Optional[dict] = None, state: Optional[flow.SamplingState] = None
):
    """(full state, sampling state, free discrete names, free continuous names, cont dists, disc dists)."""
    _, state = flow.evaluate_model_transformed(model)
    observed_rvs = set(state.observed_values)
    free_discrete = [n for n in state.discrete_distributions if n not in observed_rvs]
    free_continuous = [n for n in state.continuous_distributions if n not in observed_rvs]
    sampling_state, _ = state.as_sampling_state()
    return (
        state,
        sampling_state,
        free_discrete,
        free_continuous,
        state.continuous_distributions,
        state.discrete_distributions,
    )


class _Group(dict):
    """A dict of ndarrays ``[chain, draw, *core]`` with attribute access (xarray.Dataset stand-in)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    @property
    def data_vars(self):
        return self


class InferenceData:
    """Minimal stand-in for ``arviz.InferenceData`` (groups are dicts of ndarrays)."""

    _group_names = (
        "posterior", "sample_stats", "observed_data", "log_likelihood",
        "prior_predictive", "posterior_predictive",
    )

    def __init__(self, **groups):
        self._groups: List[str] = []
        for name in self._group_names:
            data = groups.get(name)
            if data is not None:
                setattr(self, name, _Group(data))
                self._groups.append(name)

    def groups(self) -> List[str]:
        return list(self._groups)

    def __repr__(self):
        lines = ["Inference data with groups:"]
        for g in self._groups:
            lines.append("\t> {} ({})".format(g, ", ".join(getattr(self, g).keys())))
        return "\n".join(lines)

    def __add__(self, other: "InferenceData") -> "InferenceData":
        merged = {g: dict(getattr(self, g)) for g in self._groups}
        for g in other._groups:
            merged.setdefault(g, {}).update(getattr(other, g))
        return InferenceData(**merged)

    def to_arviz(self):  # pragma: no cover - arviz is not in the image
        import arviz as az

        return az.from_dict(**{g: dict(getattr(self, g)) for g in self._groups})


def _to_numpy(v) -> np.ndarray:
    if hasattr(v, "detach"):
        return v.detach().cpu().numpy()
    if hasattr(v, "numpy") and not isinstance(v, np.ndarray):
        return np.asarray(v.numpy())
    return np.asarray(v)


def _fetch_all(groups):
    """Device tensors of several dicts -> NumPy arrays with ONE synchronisation: every tensor is
    copied asynchronously into pinned host memory (torch caches pinned blocks between calls), so a
    multi-gigabyte trace moves at PCIe speed instead of through pageable staging buffers."""
    out, pending = [], []
    for g in groups:
        if not isinstance(g, dict):
            out.append(g)
            continue
        conv = {}
        for k, v in g.items():
            if hasattr(v, "detach") and getattr(v, "is_cuda", False):
                import torch

                t = v.detach()
                host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                host.copy_(t, non_blocking=True)
                pending.append(t.device)
                conv[k] = host
            else:
                conv[k] = v
        out.append(conv)
    if pending:
        import torch

        # the copies were issued on each device's current stream: wait for those streams only (a device-wide
        # synchronisation also waits for whatever else the driver has pending on the device)
        for dev in set(pending):
            torch.cuda.current_stream(dev).synchronize()
    return [({k: _to_numpy(v) for k, v in g.items()} if isinstance(g, dict) else g) for g in out]


def trace_to_arviz(
    trace=None,
    sample_stats=None,
    observed_data=None,
    prior_predictive=None,
    posterior_predictive=None,
    log_likelihood=None,
    inplace=True,
):
    """Convert sampler output (``[draw, chain, ...]`` layout) to an ``InferenceData``."""
    if isinstance(trace, dict):
        trace = {k: v for k, v in trace.items() if "/" in k}
    if isinstance(log_likelihood, dict):
        log_likelihood = {k: v for k, v in log_likelihood.items() if "/" in k}
    trace, sample_stats, log_likelihood = _fetch_all([trace, sample_stats, log_likelihood])
    if isinstance(trace, dict):
        trace = {k: np.swapaxes(_to_numpy(v), 1, 0) for k, v in trace.items() if "/" in k}
    if isinstance(sample_stats, dict):
        sample_stats = {k: _to_numpy(v).T for k, v in sample_stats.items()}
    if isinstance(prior_predictive, dict):
        prior_predictive = {k: _to_numpy(v)[np.newaxis] for k, v in prior_predictive.items()}
    if isinstance(posterior_predictive, dict):
        if isinstance(trace, InferenceData) and inplace:
            return trace + InferenceData(posterior_predictive=posterior_predictive)
        trace = None
    if isinstance(log_likelihood, dict):
        log_likelihood = {
            k: np.swapaxes(_to_numpy(v), 1, 0) for k, v in log_likelihood.items() if "/" in k
        }
    if isinstance(observed_data, dict):
        observed_data = {k: _to_numpy(v) for k, v in observed_data.items()}
    return InferenceData(
        posterior=trace,
        sample_stats=sample_stats,
        prior_predictive=prior_predictive,
        posterior_predictive=posterior_predictive,
        log_likelihood=log_likelihood,
        observed_data=observed_data,
    )


def scope_remove_transformed_part_if_required(name: str, transformed_values: Dict[str, Any]):
    parts = name.split("/")
    if transformed_values and name in transformed_values:
        leaf = parts[-1][2:]
        parts[-1] = leaf[leaf.find("_") + 1:]
    return "/".join(parts), parts[-1]
