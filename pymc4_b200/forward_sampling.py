"""Posterior predictive sampling on the device.

Reference: ``sample_posterior_predictive`` (/root/reference/pymc4/forward_sampling.py:230-427) re-evaluates
the model once per posterior draw (``tf.vectorized_map`` over the flattened chain x draw axes) with the
posterior values substituted and the observed variables *sampled* from their distributions.  Here the
posterior is packed into states ``[chain, draw, D]`` and one kernel (generated per model,
``Model::predict``) evaluates the parameter tapes of every observed term and draws from Normal /
HalfNormal / HalfCauchy / Bernoulli / Poisson with a counter-based Philox stream keyed by
(state, element), so the same ``seed`` reproduces the same draws.

Not built (SURVEY section 8(f)1 is only half done): ``sample_prior_predictive`` and posterior predictive
draws of variables other than the observed ones (free variables are taken from the trace,
deterministics are already in it).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Union

import numpy as np

from . import ir as irmod
from .coroutine_model import Model
from .mcmc.utils import InferenceData, trace_to_arviz

__all__ = ["sample_posterior_predictive", "sample_prior_predictive"]


def sample_prior_predictive(*args, **kwargs):
    raise NotImplementedError(
        "sample_prior_predictive is listed under 'next' in SURVEY section 8(f) and is not built yet"
    )


def sample_posterior_predictive(
    model: Model,
    trace: InferenceData,
    var_names: Optional[Union[str, List[str]]] = None,
    observed: Optional[Dict[str, Any]] = None,
    use_auto_batching: bool = True,
    inplace: bool = True,
    seed: Optional[int] = None,
    dtype="float32",
    device: Optional[int] = None,
) -> InferenceData:
    """Draw the observed variables of ``model`` from their distributions for every draw of
    ``trace.posterior`` (``{name: [chain, draw, *core]}``); returns an ``InferenceData`` with a
    ``posterior_predictive`` group ``{name: [chain, draw, *observed shape]}`` (added to ``trace`` when
    ``inplace``).  Same signature as the reference plus ``seed`` / ``dtype`` / ``device``."""
    from .mcmc.samplers import _get_device_model

    if not isinstance(model, Model):
        raise TypeError("`sample_posterior_predictive` only supports `pymc4.Model` objects, but you've passed `{}`"
                        .format(type(model)))
    if var_names is not None and len(var_names) == 0:
        raise ValueError("Supplied an empty var_names list to sample from")
    if isinstance(var_names, str):
        var_names = [var_names]
    ir, _, _ = irmod.lower_model(model, observed=observed)
    obs_terms = {t.label: t for t in ir.terms if t.observed}
    if var_names is None:
        var_names = list(obs_terms)
    unknown = [v for v in var_names if v not in obs_terms]
    if unknown:
        defined = {rv.name for rv in ir.rvs} | {d.name for d in ir.dets} | set(obs_terms)
        if not set(unknown) <= defined:
            raise KeyError("The supplied var_names = {} are not defined in the model.\nDefined variables are = {}"
                           .format(sorted(set(unknown) - defined), sorted(defined)))
        raise NotImplementedError("posterior predictive draws are built for the observed variables only; "
                                  "{} are already in the trace".format(unknown))
    posterior = trace.posterior
    missing = [rv.name for rv in ir.rvs if rv.name not in posterior]
    if missing:
        raise TypeError("the trace lacks the sampled variables {} of the model".format(missing))
    dm = _get_device_model(ir, np.dtype(dtype), device)
    torch = dm.torch
    first = np.asarray(posterior[ir.rvs[0].name])
    lead = first.shape[: first.ndim - len(ir.rvs[0].shape)]
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    theta = torch.empty(tuple(lead) + (ir.D,), dtype=tdt, device=dm._dev())
    for rv in ir.rvs:
        part = torch.as_tensor(np.ascontiguousarray(posterior[rv.name]))
        theta[..., rv.offset: rv.offset + rv.size] = part.to(theta.device, tdt).reshape(tuple(lead) + (rv.size,))
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    draws = dm.posterior_predictive(theta, seed=seed, dtype=np.dtype(dtype)).cpu().numpy()
    output = {}
    for name in var_names:
        t = obs_terms[name]
        output[name] = draws[..., t.ll_offset: t.ll_offset + t.n].reshape(tuple(lead) + tuple(t.shape))
    return trace_to_arviz(trace=trace, posterior_predictive=output, inplace=inplace)
