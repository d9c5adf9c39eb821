"""Prior and posterior predictive sampling on the device.

Reference: ``sample_posterior_predictive`` (/root/reference/pymc4/forward_sampling.py:230-427) re-evaluates
the model once per posterior draw (``tf.vectorized_map`` over the flattened chain x draw axes) with the
posterior values substituted and the observed variables *sampled* from their distributions.  Here the
posterior is packed into states ``[chain, draw, D]`` and one kernel (generated per model,
``Model::predict``) evaluates the parameter tapes of every observed term and draws from Normal /
HalfNormal / HalfCauchy / Bernoulli / Poisson with a counter-based Philox stream keyed by
(state, element), so the same ``seed`` reproduces the same draws.

``sample_prior_predictive`` (forward_sampling.py:26-227) is ancestral sampling of the whole model, one warp
per sample (``Model::prior``).  Not built: posterior predictive draws of variables other than the observed
ones (free variables are taken from the trace, deterministics are already in it).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Union

import numpy as np

from . import ir as irmod
from .coroutine_model import Model
from .mcmc.utils import InferenceData, trace_to_arviz

__all__ = ["sample_posterior_predictive", "sample_prior_predictive"]


def sample_prior_predictive(
    model: Model,
    sample_shape: Union[int, tuple] = 1000,
    sample_from_observed: bool = True,
    var_names: Optional[Union[str, List[str]]] = None,
    state=None,
    use_auto_batching: bool = True,
    seed: Optional[int] = None,
    dtype="float32",
    device: Optional[int] = None,
) -> InferenceData:
    """Draw ``sample_shape`` ancestral samples of the model (reference forward_sampling.py:26-227): every
    free variable from its distribution given the ones drawn before it, the deterministics, and -- with
    ``sample_from_observed`` -- the observed variables.  Returns an ``InferenceData`` with a
    ``prior_predictive`` group ``{name: [1, *sample_shape, *core]}`` keyed by the untransformed names.

    Limitation: observed values enter the lowered model as constants, so a free variable that depends on
    an observed one still sees the observed data when ``sample_from_observed`` is set."""
    from .mcmc.samplers import _get_device_model

    if not isinstance(model, Model):
        raise TypeError("`sample_prior_predictive` only supports `pymc4.Model` objects, but you've passed `{}`"
                        .format(type(model)))
    if state is not None:
        raise NotImplementedError("sample_prior_predictive(state=...) is not supported")
    if isinstance(sample_shape, int):
        sample_shape = (sample_shape,)
    sample_shape = tuple(int(s) for s in sample_shape)
    if isinstance(var_names, str):
        var_names = [var_names]
    ir, _, _ = irmod.lower_model(model)
    obs_terms = {t.label: t for t in ir.terms if t.observed}
    dist_names = [rv.untransformed_name or rv.name for rv in ir.rvs]
    if sample_from_observed:
        dist_names = dist_names + list(obs_terms)
    back_transformed = {rv.untransformed_name for rv in ir.rvs if rv.untransformed_name}
    # the model's own return value is traced under the bare model name: like trace_to_arviz, keep scoped names only
    det_names = [d.name for d in ir.dets if d.name not in back_transformed and "/" in d.name]
    traced_observeds = set()
    if var_names is None:
        var_names = dist_names + det_names
    else:
        traced_observeds = {v for v in var_names if v in obs_terms}
    if not set(var_names) <= (set(dist_names + det_names) | traced_observeds):
        raise ValueError("Some of the supplied var_names are not defined in the supplied model {}.\n"
                         "List of unknown var_names: {}".format(model, sorted(set(var_names) - set(dist_names + det_names))))
    dm = _get_device_model(ir, np.dtype(dtype), device)
    torch = dm.torch
    M = int(np.prod(sample_shape)) if sample_shape else 1
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    x, obs = dm.prior_predictive(M, seed=seed, dtype=np.dtype(dtype),
                                 sample_observed=sample_from_observed)
    output: Dict[str, Any] = {}
    lead = sample_shape
    for rv in ir.rvs:
        name = rv.untransformed_name or rv.name
        if name in var_names:
            output[name] = x[:, rv.offset: rv.offset + rv.size].reshape(lead + rv.shape).cpu().numpy()
    if any(n in var_names for n in det_names):
        theta = _unconstrain(ir, x, torch)  # deterministics are tapes over the unconstrained state
        dets = dm.deterministics(theta, dtype=np.dtype(dtype)).cpu().numpy()
        for d in ir.dets:
            if d.name in var_names and d.name in det_names:
                output[d.name] = dets[:, d.out_offset: d.out_offset + d.n].reshape(lead + tuple(d.shape))
    gp_index = {t.label: k for k, t in enumerate(ir.gp_terms)}
    for name, t in obs_terms.items():
        if name in var_names:
            if obs is not None and name in gp_index:
                # MvNormal observation of a GP model: a vector per sample, loc + chol(K(x)) z on the device
                theta = _unconstrain(ir, x, torch)
                output[name] = dm.gp_predictive(theta, seed=seed, dtype=np.dtype(dtype), which=gp_index[name]) \
                    .reshape(lead + (t.gp.n,)).cpu().numpy()
            elif obs is not None:
                output[name] = obs[:, t.ll_offset: t.ll_offset + t.n].reshape(lead + tuple(t.shape)).cpu().numpy()
            else:  # explicitly requested but not resampled: the observed data itself
                output[name] = np.broadcast_to(np.asarray(ir.observed[name]), lead + np.shape(ir.observed[name])).copy()
    return trace_to_arviz(prior_predictive=output)


def _unconstrain(ir, x, torch):
    """Constrained values x [M, D] of the free variables -> unconstrained state (undo the transforms)."""
    theta = x.clone()
    for rv in ir.rvs:
        sl = slice(rv.offset, rv.offset + rv.size)
        if rv.transform == irmod.TR_EXP:
            theta[:, sl] = torch.log((x[:, sl] - rv.shift) / rv.scale)
        elif rv.transform == irmod.TR_SIGMOID:
            u = (x[:, sl] - rv.shift) / rv.scale
            theta[:, sl] = torch.log(u) - torch.log1p(-u)
    return theta


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
    posterior = trace.posterior
    # var_names may also name variables that are NOT re-drawn (reference forward_sampling.py:318-323: the values of
    # the executed model, i.e. the posterior values the trace was passed in with, and recomputed deterministics)
    det_by_name = {d.name: d for d in ir.dets}
    unknown = [v for v in var_names if v not in obs_terms and v not in det_by_name and v not in posterior]
    if unknown:
        defined = {rv.name for rv in ir.rvs} | set(det_by_name) | set(obs_terms)
        raise KeyError("The supplied var_names = {} are not defined in the model.\nDefined variables are = {}"
                       .format(sorted(unknown), sorted(defined)))
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
    gp_index = {t.label: k for k, t in enumerate(ir.gp_terms)}
    output = {}
    dets = None
    for name in var_names:
        if name not in obs_terms:
            if name in det_by_name:  # deterministics (and back-transformed variables) are recomputed from the draws
                if dets is None:
                    dets = dm.deterministics(theta, dtype=np.dtype(dtype)).cpu().numpy()
                d = det_by_name[name]
                output[name] = dets[..., d.out_offset: d.out_offset + d.n].reshape(tuple(lead) + tuple(d.shape))
            else:  # a sampled variable: its posterior values
                output[name] = np.asarray(posterior[name])
            continue
        t = obs_terms[name]
        if name in gp_index:  # MvNormal observation of a GP model: a vector per draw, loc + chol(K(theta)) z
            output[name] = dm.gp_predictive(theta, seed=seed, dtype=np.dtype(dtype), which=gp_index[name]).cpu().numpy()
        else:
            output[name] = draws[..., t.ll_offset: t.ll_offset + t.n].reshape(tuple(lead) + tuple(t.shape))
    return trace_to_arviz(trace=trace, posterior_predictive=output, inplace=inplace)
