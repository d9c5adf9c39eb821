"""Sampler registry and the NUTS / HMC samplers behind ``pm.sample``.

Class protocol, kwarg routing and return layout follow the reference
(pymc4/mcmc/samplers.py:25-262 registry + ``_BaseSampler``, :265-315 ``HMC``, :342-419 ``NUTS``,
:729-829 ``build_logp_and_deterministic_functions`` / ``vectorize_logp_function`` /
``tile_init``).  The one call that changes is ``_run_chains`` (:172-189): instead of
``tfp.mcmc.sample_chain`` it calls ``pm4_nuts_run`` / ``pm4_hmc_run`` of libpm4b200 through
ctypes, with torch tensors as device buffers.  ``hmc_simple`` / ``nuts_simple`` (:318-339, :422-447:
``SimpleStepSizeAdaptation``) and ``rwm`` (:450-478: random-walk Metropolis) run on the same kernels;
``compound`` (discrete latents, :481-726) partitions the free variables into blocks and calls ``pm4_compound_run``:
one transition of every block's kernel per compound transition, the other blocks' dimensions frozen.
"""
import abc
import inspect
import itertools
import logging
import os
from functools import partial
from typing import Any, Dict, List, Optional

from collections import OrderedDict

import numpy as np

from .. import flow
from .. import ir as irmod
from .._cstructs import (PM4_ADAPT_SIMPLE, PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, PM4_PROPOSAL_RANDOM_WALK, make_adapt,
                         make_hmc_cfg, make_nuts_cfg)
from ..coroutine_model import Model
from . import kernels as mcmc
from .utils import (
    KERNEL_KWARGS_SET,
    initialize_sampling_state,
    initialize_state,
    scope_remove_transformed_part_if_required,
    trace_to_arviz,
)

__all__ = ["HMC", "HMCSimple", "NUTS", "NUTSSimple", "RandomWalkM", "CompoundStep"]
reg_samplers: Dict[str, type] = {}

_log = logging.getLogger("pymc4.sampling")
if not _log.handlers:
    _log.addHandler(logging.StreamHandler())
_log.setLevel(logging.INFO)

# keyword arguments of this backend that the reference does not have
_BACKEND_KWARGS = ("dtype", "device", "chain_offset", "warps_per_block", "comm", "trace_slabs", "mass_windows")


def register_sampler(cls):
    reg_samplers[cls._name] = cls
    return cls


def _signature_keys(fn, skip: int = 0) -> set:
    return set(list(inspect.signature(fn).parameters.keys())[skip:])


class _BaseSampler(metaclass=abc.ABCMeta):
    _grad = False
    _name: str = ""
    _kernel: Any = None
    _adaptation: Any = None
    default_kernel_kwargs: dict = {}
    default_adapter_kwargs: dict = {}

    def __init__(self, model: Model, **kwargs):
        if not isinstance(model, Model):
            raise TypeError(
                "`sample` function only supports `pymc4.Model` objects, but you've passed `{}`".format(type(model))
            )
        _, _, disc_names, cont_names, _, _ = initialize_state(model)
        if self._grad is True and disc_names:
            raise ValueError("Discrete distributions can't be used with gradient-based sampler")
        self.model = model
        self.stat_names: List[str] = []
        self.parent_inds: List[int] = []
        self.backend_kwargs = {k: kwargs.pop(k) for k in _BACKEND_KWARGS if k in kwargs}
        self._assign_arguments(kwargs)
        self._check_arguments()
        self._bound_kwargs()

    # -- kwarg routing (reference :191-229) ----------------------------------
    def _assign_arguments(self, kwargs):
        keys = set(kwargs)
        adaptation_keys = _signature_keys(self._adaptation.__init__, 1) if self._adaptation else set()
        kernel_keys = _signature_keys(self._kernel.__init__, 1)
        chain_keys = _signature_keys(mcmc.sample_chain)
        self.adaptation_kwargs = {k: kwargs[k] for k in adaptation_keys & keys}
        self.kernel_kwargs = {k: kwargs[k] for k in kernel_keys & keys}
        self.chain_kwargs = {k: kwargs[k] for k in chain_keys & keys}

    def _check_arguments(self):
        a, k, c = self.adaptation_kwargs.keys(), self.kernel_kwargs.keys(), self.chain_kwargs.keys()
        if (a & k) or (a & c) or (k & c):
            raise ValueError("Ambiguity in setting kwargs for `kernel`, `adaptation_kernel`, `chain_sampler`")

    def _bound_kwargs(self, *args):
        for k, v in self.default_kernel_kwargs.items():
            self.kernel_kwargs.setdefault(k, v)
        for k, v in self.default_adapter_kwargs.items():
            self.adaptation_kwargs.setdefault(k, v)

    def __call__(self, *args, **kwargs):
        return self._sample(*args, **kwargs)

    @classmethod
    def _default_kernel_maker(cls):
        return KERNEL_KWARGS_SET(
            kernel=cls._kernel,
            adaptive_kernel=cls._adaptation,
            kernel_kwargs=cls.default_kernel_kwargs,
            adaptive_kwargs=cls.default_adapter_kwargs,
        )

    @abc.abstractmethod
    def trace_fn(self, current_state, pkr):
        """Pick the per-draw statistics out of the kernel results (a dict of device tensors)."""

    # -- the sampling call (reference :78-170) ---------------------------------
    def _sample(
        self,
        *,
        num_samples: int = 1000,
        num_chains: int = 10,
        burn_in: int = 100,
        observed: Optional[dict] = None,
        state: Optional[flow.SamplingState] = None,
        use_auto_batching: bool = True,
        xla: bool = False,
        seed: Optional[int] = None,
        is_compound: bool = False,
        trace_discrete: Optional[List[str]] = None,
        include_log_likelihood=False,
    ):
        if state is not None and observed is not None:
            raise ValueError("Can't use both `state` and `observed` arguments")
        dtype = np.dtype(self.backend_kwargs.get("dtype", "float32"))
        (logpfn, init, _deterministics_callback, deterministic_names, state_) = build_logp_and_deterministic_functions(
            self.model,
            num_chains=num_chains,
            state=state,
            observed=observed,
            collect_reduced_log_prob=use_auto_batching,
            parent_inds=self.parent_inds if is_compound else None,
            dtype=dtype,
            device=self.backend_kwargs.get("device", None),
        )
        init_state = list(init.values())
        init_keys = list(init.keys())

        # the device code evaluates one chain per warp whatever `use_auto_batching` says: a model
        # written for manual batching is also a valid per-chain model
        self.parallel_logpfn = vectorize_logp_function(logpfn)
        self.deterministics_callback = vectorize_logp_function(_deterministics_callback)
        init_state = tile_init(init_state, num_chains)

        self._num_samples = num_samples
        self.seed = seed
        self._dtype = dtype
        self._device_model = logpfn.device_model
        self._ir_own = logpfn.ir  # names come from THIS model: a cached device model may have been built for a same-structure model with other names
        if xla:
            _log.debug("`xla=True` has no effect: the sampler is already one fused CUDA kernel")

        results, sample_stats = self._run_chains(init_state, burn_in)

        posterior = dict(zip(init_keys, results))
        if trace_discrete:
            init_keys_ = [scope_remove_transformed_part_if_required(k, state_.transformed_values)[1] for k in init_keys]
            for name in trace_discrete:
                key = init_keys[init_keys_.index(name)]
                v = posterior[key]
                posterior[key] = v.astype(np.int32) if isinstance(v, np.ndarray) else v.to(dtype=self._device_model.torch.int32)

        deterministic_values = ()
        if len(sample_stats) > len(self.stat_names):
            deterministic_values = sample_stats[len(self.stat_names):]
            sample_stats = sample_stats[: len(self.stat_names)]
        sampler_stats = dict(zip(self.stat_names, sample_stats))
        if len(deterministic_names) > 0:
            det_names = [d.name for d in self._ir.dets]
            posterior.update(dict(zip(det_names, deterministic_values)))

        log_likelihood_dict = dict()
        if include_log_likelihood:
            log_likelihood_dict = _log_likelihood_from_states(self._device_model, self._last_states, self._dtype, ir=self._ir)
        self.last_run_info = self._run_info
        return trace_to_arviz(
            posterior,
            sampler_stats if is_compound is False else None,
            observed_data=state_.observed_values,
            log_likelihood=log_likelihood_dict if include_log_likelihood else None,
        )

    # -- step-size policy -------------------------------------------------------
    def _step_size_policy(self, step_size, num_chains: int):
        """Map the reference's ``step_size`` conventions onto (eps0, per-dim scale, mode).

        scalar                       -> one pooled step size adapted from the across-chain mean
                                        acceptance (tfp reduces over all chains);
        array with leading chain axis (``[C]``, ``[C, 1]``) -> independent per-chain adaptation;
        list of per-part arrays ``[C, *core]`` (notebooks/radon_hierarchical.ipynb:123-146)
                                     -> per-chain adaptation of a multiplier of a fixed per-dimension
                                        scale (all dimensions of a chain see the same update).
        """
        ir = self._ir
        self._chain_step_sizes = None
        if isinstance(step_size, (list, tuple)) and len(step_size) == len(ir.rvs):
            scale = np.ones(ir.D, dtype=np.float64)
            for rv, part in zip(ir.rvs, step_size):
                part = np.asarray(part.detach().cpu().numpy() if hasattr(part, "detach") else part, dtype=np.float64)
                full = np.broadcast_to(part, (num_chains,) + rv.shape) if part.ndim else np.full((num_chains,) + rv.shape, part)
                if not np.allclose(full, full[0]):
                    raise NotImplementedError("step sizes that differ between chains at start are not supported")
                scale[rv.offset: rv.offset + rv.size] = full[0].reshape(-1)
            return 1.0, scale, PM4_EPS_PER_CHAIN
        arr = np.asarray(step_size.detach().cpu().numpy() if hasattr(step_size, "detach") else step_size, dtype=np.float64)
        if arr.ndim == 0:
            return float(arr), None, PM4_EPS_POOLED
        if arr.shape[0] == num_chains and arr.size == num_chains:
            flat = arr.reshape(-1)
            if not np.all(flat > 0):
                raise ValueError("step sizes must be positive")
            # chains that start from different step sizes: the array travels to the device (cfg.step_size_dev)
            self._chain_step_sizes = None if np.allclose(flat, flat[0]) else flat.copy()
            return float(flat[0]), None, PM4_EPS_PER_CHAIN
        if arr.size == ir.D:
            return 1.0, arr.reshape(-1), PM4_EPS_POOLED
        raise ValueError("cannot interpret step_size of shape {} for {} chains, D={}".format(arr.shape, num_chains, ir.D))

    @property
    def _ir(self):
        """The lowered model of this sampler's own model.  `_get_device_model` shares device models between models of
        identical structure and layout, whose variable names may differ: offsets and shapes are the cached model's,
        names must be this model's."""
        own = getattr(self, "_ir_own", None)
        return own if own is not None else self._device_model.ir

    def _pack_init(self, init: List[np.ndarray]) -> np.ndarray:
        ir = self._ir
        C = int(np.shape(init[0])[0])
        theta0 = np.empty((C, ir.D), dtype=np.float64)
        for rv, part in zip(ir.rvs, init):
            theta0[:, rv.offset: rv.offset + rv.size] = np.asarray(part, dtype=np.float64).reshape(C, rv.size)
        return theta0

    def _unpack_states(self, states):
        """device [S, C, D] -> list of per-variable views [S, C, *core] in trace order."""
        S, C = states.shape[0], states.shape[1]
        return [states[:, :, rv.offset: rv.offset + rv.size].reshape((S, C) + rv.shape)
                for rv in self._ir.rvs]

    def _seed_value(self) -> int:
        if self.seed is None:
            return int.from_bytes(os.urandom(8), "little")
        return int(self.seed) & (2**64 - 1)

    def _adapt_cfg(self, mode: int):
        a = self.adaptation_kwargs
        if self._adaptation is None or "num_adaptation_steps" not in a:
            return make_adapt(mode=mode, enabled=False)
        spec = self._adaptation(inner_kernel=None, **{k: v for k, v in a.items() if not callable(v)})
        if isinstance(spec, mcmc.SimpleStepSizeAdaptation):
            return make_adapt(mode=mode, num_adaptation_steps=spec.num_adaptation_steps,
                              target_accept_prob=spec.target_accept_prob, kind=PM4_ADAPT_SIMPLE,
                              adaptation_rate=spec.adaptation_rate, enabled=True)
        # mass_windows (backend extension, NUTS): diagonal mass-matrix adaptation from the cross-chain spread at that many
        # points of the burn-in -- what the reference's radon notebook emulates by passing posterior standard deviations
        # as per-dimension step sizes (notebooks/radon_hierarchical.ipynb:123-146)
        return make_adapt(mode=mode, num_adaptation_steps=spec.num_adaptation_steps,
                          target_accept_prob=spec.target_accept_prob,
                          exploration_shrinkage=spec.exploration_shrinkage,
                          step_count_smoothing=spec.step_count_smoothing, decay_rate=spec.decay_rate, enabled=True,
                          mass_windows=int(self.backend_kwargs.get("mass_windows", 0)) if self._name.startswith("nuts") else 0,
                          shrinkage_target=spec.shrinkage_target)

    def _check_chain_kwargs(self):
        ck = self.chain_kwargs
        if int(ck.get("num_steps_between_results", 0) or 0) < 0:
            raise ValueError("num_steps_between_results must be non-negative")
        if ck.get("return_final_kernel_results", False) or ck.get("previous_kernel_results") is not None:
            raise NotImplementedError("resuming from kernel results is not supported")

    # thinning (tfp sample_chain: `num_steps_between_results + 1` kernel steps per kept result, the result being the
    # state after the last of them): the device runs every step and the kept draws are the strided view
    def _thin(self) -> int:
        return int(self.chain_kwargs.get("num_steps_between_results", 0) or 0)

    def _draws_to_run(self) -> int:
        return self._num_samples * (self._thin() + 1)

    def _thin_out(self, out: dict, keys):
        k = self._thin()
        if k:
            for key in keys:
                if out.get(key) is not None:
                    out[key] = out[key][k::k + 1]
        return out

    def _deterministic_values(self, states):
        ir = self._ir
        if not ir.dets:
            return ()
        out = self._device_model.deterministics(states, dtype=self._dtype)
        S, C = states.shape[0], states.shape[1]
        return tuple(out[:, :, d.out_offset: d.out_offset + d.n].reshape((S, C) + d.shape) for d in ir.dets)

    @abc.abstractmethod
    def _run_chains(self, init, burn_in):
        """Run the chains on the device; return (results, sample_stats) in the reference layout."""


@register_sampler
class HMC(_BaseSampler):
    """Hamiltonian Monte Carlo with a fixed number of leapfrog steps under dual-averaging
    step-size adaptation (defaults ``step_size=0.1, num_leapfrog_steps=3``; reference :265-315)."""

    _name = "hmc"
    _grad = True
    _adaptation = mcmc.DualAveragingStepSizeAdaptation
    _kernel = mcmc.HamiltonianMonteCarlo

    default_kernel_kwargs: dict = {"step_size": 0.1, "num_leapfrog_steps": 3}
    default_adapter_kwargs: dict = {}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.stat_names = ["mean_tree_accept"]

    def trace_fn(self, current_state, pkr):
        return (pkr["log_accept"],) + tuple(self._deterministic_values(current_state))

    def _run_chains(self, init, burn_in):
        self._check_chain_kwargs()
        dm = self._device_model
        theta0 = self._pack_init(init)
        C = theta0.shape[0]
        spec = self._kernel(target_log_prob_fn=self.parallel_logpfn, **self.kernel_kwargs)
        eps0, scale, mode = self._step_size_policy(spec.step_size, C)
        adapt = self._adapt_cfg(mode)
        cfg = make_hmc_cfg(C, self._draws_to_run(), burn_in, step_size=eps0,
                           num_leapfrog_steps=spec.num_leapfrog_steps, seed=self._seed_value(), adapt=adapt,
                           warps_per_block=self.backend_kwargs.get("warps_per_block", 0),
                           chain_offset=self.backend_kwargs.get("chain_offset", 0))
        if "comm" in self.backend_kwargs:
            dm.attach_comm(self.backend_kwargs["comm"])
        out = self._thin_out(dm.hmc(cfg, theta0, dtype=self._dtype, step_scale=scale, step_sizes=self._chain_step_sizes),
                             ("states", "lp", "log_accept", "accepted"))
        self._run_info = dict(step_size=out["step_size"], kernel="hmc",
                              leapfrogs_per_transition=spec.num_leapfrog_steps)
        self._last_states = out["states"]
        results = self._unpack_states(out["states"])
        return results, self.trace_fn(out["states"], out)


@register_sampler
class HMCSimple(HMC):
    """HMC under ``SimpleStepSizeAdaptation`` (multiplicative step-size updates; reference :318-339)."""

    _name = "hmc_simple"
    _adaptation = mcmc.SimpleStepSizeAdaptation


@register_sampler
class NUTS(_BaseSampler):
    """No-U-Turn sampler under dual-averaging step-size adaptation (default ``step_size=0.1``;
    stats ``lp, tree_size, diverging, energy, mean_tree_accept``; reference :342-419)."""

    _name = "nuts"
    _grad = True
    _adaptation = mcmc.DualAveragingStepSizeAdaptation
    _kernel = mcmc.NoUTurnSampler

    default_adapter_kwargs: dict = {
        "num_adaptation_steps": 100,
        "step_size_getter_fn": lambda pkr: pkr.step_size,
        "log_accept_prob_getter_fn": lambda pkr: pkr.log_accept_ratio,
        "step_size_setter_fn": lambda pkr, new_step_size: pkr._replace(step_size=new_step_size),
    }
    default_kernel_kwargs: dict = {"step_size": 0.1}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.stat_names = ["lp", "tree_size", "diverging", "energy", "mean_tree_accept"]

    def trace_fn(self, current_state, pkr):
        return (
            pkr["lp"],
            pkr["leapfrogs"],
            pkr["diverging"].to(dtype=self._device_model.torch.bool),
            pkr["energy"],
            pkr["log_accept"],
        ) + tuple(self._deterministic_values(current_state))

    def _run_chains(self, init, burn_in):
        self._check_chain_kwargs()
        dm = self._device_model
        theta0 = self._pack_init(init)
        C = theta0.shape[0]
        spec = self._kernel(target_log_prob_fn=self.parallel_logpfn, **self.kernel_kwargs)
        eps0, scale, mode = self._step_size_policy(spec.step_size, C)
        cfg = make_nuts_cfg(C, self._draws_to_run(), burn_in, step_size=eps0, seed=self._seed_value(),
                            max_tree_depth=spec.max_tree_depth, max_energy_diff=spec.max_energy_diff,
                            adapt=self._adapt_cfg(mode),
                            warps_per_block=self.backend_kwargs.get("warps_per_block", 0),
                            chain_offset=self.backend_kwargs.get("chain_offset", 0))
        if "comm" in self.backend_kwargs:  # chains of this call are one shard of a box-wide run
            dm.attach_comm(self.backend_kwargs["comm"])
        # a large trace leaves the device in slabs while the sampler keeps running (pm4_model_set_trace_sink): the
        # posterior arrays are then views of one pinned host buffer instead of per-variable copies made afterwards
        trace_bytes = self._draws_to_run() * C * dm.ir.D * np.dtype(self._dtype).itemsize
        slabs = int(self.backend_kwargs.get("trace_slabs", 8 if trace_bytes >= (64 << 20) else 0))
        out = dm.nuts(cfg, theta0, dtype=self._dtype, step_scale=scale, stream_trace=slabs, step_sizes=self._chain_step_sizes)
        if "states_host" in out and self._thin():
            out["states_host_ready"]()
        self._thin_out(out, ("states", "lp", "leapfrogs", "diverging", "energy", "log_accept", "depth", "states_host"))
        self._run_info = dict(step_size=out["step_size"], total_leapfrogs=out["total_leapfrogs"],
                              depth=out["depth"], kernel="nuts", mass_scale=out.get("mass_scale"))
        self._last_states = out["states"]
        stats = self.trace_fn(out["states"], out)
        if "states_host" in out:
            out["states_host_ready"]()
            host = out["states_host"].numpy()
            S = host.shape[0]
            results = [host[:, :, rv.offset: rv.offset + rv.size].reshape((S, C) + rv.shape) for rv in self._ir.rvs]
        else:
            results = self._unpack_states(out["states"])
        return results, stats


@register_sampler
class NUTSSimple(NUTS):
    """NUTS under ``SimpleStepSizeAdaptation`` (reference :422-447)."""

    _name = "nuts_simple"
    _adaptation = mcmc.SimpleStepSizeAdaptation
    default_adapter_kwargs: dict = {k: v for k, v in NUTS.default_adapter_kwargs.items()}


@register_sampler
class RandomWalkM(_BaseSampler):
    """Random-walk Metropolis (gradient-free): proposal = state + Normal(0, 1) perturbation, MH accept
    (tfp RandomWalkMetropolis with its default ``random_walk_normal_fn``; reference :450-478).
    ``stat_names = ["mean_accept"]``; ``trace_fn`` returns ``(log_accept_ratio, *deterministics)``."""

    _name = "rwm"
    _adaptation = None
    _kernel = mcmc.RandomWalkMetropolis
    _grad = False

    default_kernel_kwargs: dict = {}
    default_adapter_kwargs: dict = {}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.stat_names = ["mean_accept"]

    def trace_fn(self, current_state, pkr):
        return (pkr["log_accept"],) + tuple(self._deterministic_values(current_state))

    def _run_chains(self, init, burn_in):
        self._check_chain_kwargs()
        dm = self._device_model
        theta0 = self._pack_init(init)
        C = theta0.shape[0]
        spec = self._kernel(target_log_prob_fn=self.parallel_logpfn, **self.kernel_kwargs)
        cfg = make_hmc_cfg(C, self._draws_to_run(), burn_in, step_size=1.0, num_leapfrog_steps=0,
                           seed=self._seed_value(), adapt=make_adapt(enabled=False),
                           warps_per_block=self.backend_kwargs.get("warps_per_block", 0),
                           chain_offset=self.backend_kwargs.get("chain_offset", 0),
                           proposal=PM4_PROPOSAL_RANDOM_WALK, rw_scale=spec.scale)
        out = self._thin_out(dm.hmc(cfg, theta0, dtype=self._dtype), ("states", "lp", "log_accept", "accepted"))
        self._run_info = dict(kernel="rwm", accepted=out["accepted"])
        self._last_states = out["states"]
        results = self._unpack_states(out["states"])
        return results, self.trace_fn(out["states"], out)


@register_sampler
class CompoundStep(_BaseSampler):
    """The compound step (reference :481-726): every free variable gets a sampler -- the user's ``sampler_methods``
    entry, else random-walk Metropolis with the distribution's own proposal function (Bernoulli, Categorical), else
    NUTS when the distribution has a gradient and random-walk Metropolis when not; variables with the same sampler
    and arguments are merged into one block; a transition runs every block's kernel once, in order.

    ``stat_names = ["compound_results"]``; like the reference the sampler statistics are not put in the trace."""

    _name = "compound"
    _adaptation = None
    _kernel = mcmc._CompoundStepTF
    _grad = False

    default_adapter_kwargs: dict = {}
    default_kernel_kwargs: dict = {}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.stat_names = ["compound_results"]
        self._block_vars: List[List[str]] = []

    def trace_fn(self, current_state, pkr):
        return (pkr["log_accept"],) + tuple(self._deterministic_values(current_state))

    @staticmethod
    def _convert_sampler_methods(sampler_methods):
        if not sampler_methods:
            return {}
        sampler_methods_dict = {}
        for sampler_item in sampler_methods:
            if len(sampler_item) not in [2, 3]:
                raise ValueError(
                    "You need to provide `sampler_methods` as the tuple of the length in [2, 3]. If additional kwargs "
                    "for kernel are provided then the lenght of tuple is equal to 3."
                )
            if len(sampler_item) < 3:
                var, sampler = sampler_item
                kwargs = {}
            else:
                var, sampler, kwargs = sampler_item
            if isinstance(var, (list, tuple)):
                for var_ in var:
                    sampler_methods_dict[var_] = (sampler, kwargs)
            else:
                sampler_methods_dict[var] = (sampler, kwargs)
        return sampler_methods_dict

    def _merge_samplers(self, make_kernel_fn, part_kernel_kwargs):
        """Union variables whose sampler class and arguments are equal (proposal functions compared as instances);
        returns (one (sampler, kwargs) per set, set sizes) and stores the variable order in ``parent_inds``."""
        num_vars = len(make_kernel_fn)
        parents = list(range(num_vars))
        key = "new_state_fn"

        def same_kwargs(item1, item2):
            if (key in item1) != (key in item2):
                return False
            if key in item1 and not (item1[key].__self__ == item2[key].__self__):
                return False
            rest1 = {k: v for k, v in item1.items() if k != key}
            rest2 = {k: v for k, v in item2.items() if k != key}
            return rest1 == rest2

        def get_set(p):
            return p if parents[p] == p else get_set(parents[p])

        for i, j in itertools.combinations(range(num_vars), 2):
            if make_kernel_fn[i] == make_kernel_fn[j] and same_kwargs(part_kernel_kwargs[i], part_kernel_kwargs[j]):
                p1, p2 = get_set(i), get_set(j)
                if p1 != p2:
                    parents[max(p1, p2)] = min(p1, p2)
        parents = [get_set(p) for p in parents]
        kernels, used = [], {}
        for i, p in enumerate(parents):
            if p not in used:
                kernels.append((make_kernel_fn[i], part_kernel_kwargs[i]))
                used[p] = True
        parent_used, set_lengths = {}, []
        for p in parents:
            if p in parent_used:
                set_lengths[parent_used[p]] += 1
            else:
                parent_used[p] = len(set_lengths)
                set_lengths.append(1)
        self.parent_inds = sorted(range(len(parents)), key=lambda k: parents[k])
        return kernels, set_lengths

    def _assign_default_methods(self, *, sampler_methods: Optional[List] = None,
                                state: Optional[flow.SamplingState] = None, observed: Optional[dict] = None):
        converted = CompoundStep._convert_sampler_methods(sampler_methods)
        (_, state, _, _, continuous_distrs, discrete_distrs) = initialize_state(self.model, observed=observed, state=state)
        init_keys = list(state.all_unobserved_values.keys())

        make_kernel_fn: list = []
        part_kernel_kwargs: list = []
        func_names: list = []
        for key in init_keys:
            untrs_var, unscoped_tr_var = scope_remove_transformed_part_if_required(key, state.transformed_values)
            distr = continuous_distrs.get(untrs_var, None)
            if distr is None:
                distr = discrete_distrs[untrs_var]
            func = distr._default_new_state_part
            if unscoped_tr_var in converted:
                sampler, kwargs = converted[unscoped_tr_var]
                if not distr._grad_support and sampler._grad:
                    raise ValueError(
                        "The `{}` doesn't support gradient, please provide an appropriate sampler method".format(
                            unscoped_tr_var))
                make_kernel_fn.append(sampler)
                part_kernel_kwargs.append(dict(kwargs))
                user_fn = part_kernel_kwargs[-1].get("new_state_fn", None)
                if user_fn is not None:
                    func = user_fn
                if func and sampler._name == "rwm":
                    part_kernel_kwargs[-1]["new_state_fn"] = partial(func)()
            elif callable(func):
                make_kernel_fn.append(RandomWalkM)
                part_kernel_kwargs.append({"new_state_fn": partial(func)()})
            else:
                make_kernel_fn.append(NUTS if distr._grad_support else RandomWalkM)
                part_kernel_kwargs.append({})
            func_names.append(func._name if func else "default")

        kernels, set_lengths = self._merge_samplers(make_kernel_fn, part_kernel_kwargs)
        CompoundStep._log_variables(init_keys, kernels, set_lengths, self.parent_inds, func_names)
        self.kernel_kwargs["compound_samplers"] = kernels
        self.kernel_kwargs["compound_set_lengths"] = set_lengths
        ordered = [init_keys[i] for i in self.parent_inds]
        self._block_vars, at = [], 0
        for n in set_lengths:
            self._block_vars.append(ordered[at: at + n])
            at += n

    def __call__(self, *args, **kwargs):
        if "compound_samplers" not in self.kernel_kwargs:
            self._assign_default_methods(observed=kwargs.get("observed"), state=kwargs.get("state"))
        return self._sample(*args, is_compound=True, **kwargs)

    @staticmethod
    def _log_variables(var_keys, kernel_kwargs, set_lengths, parent_inds, func_names):
        var_keys = [var_keys[i] for i in parent_inds]
        func_names = [func_names[i] for i in parent_inds]
        log_output, curr = "", 0
        for i, ((kernel, _), set_len) in enumerate(zip(kernel_kwargs, set_lengths)):
            vars_ = var_keys[curr: curr + set_len]
            log_output += ("\n" if i > 0 else "") + " -- {}[vars={}, proposal_function={}]".format(
                kernel._name, [item.split("/")[1] for item in vars_], func_names[curr])
            curr += set_len
        _log.info(log_output)

    def _blocks(self, burn_in: int):
        """(block records for ``DeviceModel.compound``, one per merged sampler) from ``compound_samplers``."""
        ir = self._ir
        by_name = {rv.name: rv for rv in ir.rvs}
        blocks = []
        for (sampler, user_kwargs), names in zip(self.kernel_kwargs["compound_samplers"], self._block_vars):
            mask = np.zeros(ir.D, dtype=np.float64)
            for n in names:
                rv = by_name[n]
                mask[rv.offset: rv.offset + rv.size] = 1.0
            mk = sampler._default_kernel_maker()
            # the reference builds the sub-kernel from {**user kwargs, **class defaults} (tf_support.py:41): keep that
            # precedence for the keys the class defines, pass everything else through
            kk = {**{k: v for k, v in user_kwargs.items()}, **mk.kernel_kwargs}
            if sampler._name in ("nuts", "nuts_simple", "hmc", "hmc_simple"):
                discrete = [n for n in names if by_name[n].discrete]
                if discrete:
                    raise ValueError("Discrete distributions can't be used with gradient-based sampler: {}".format(discrete))
                ak = dict(mk.adaptive_kwargs)
                spec_kwargs = {k: v for k, v in kk.items() if k in _signature_keys(mk.kernel.__init__, 1)}
                spec = mk.kernel(target_log_prob_fn=None, **spec_kwargs)
                adapt_keys = _signature_keys(mk.adaptive_kernel.__init__, 1)
                ak.update({k: v for k, v in user_kwargs.items() if k in adapt_keys})
                ak.setdefault("num_adaptation_steps", burn_in)
                ad = mk.adaptive_kernel(inner_kernel=None, **{k: v for k, v in ak.items() if not callable(v)})
                eps0, scale, mode = self._step_size_policy(spec.step_size, self._num_chains)
                if isinstance(ad, mcmc.SimpleStepSizeAdaptation):
                    adapt = make_adapt(mode=mode, num_adaptation_steps=ad.num_adaptation_steps,
                                       target_accept_prob=ad.target_accept_prob, kind=PM4_ADAPT_SIMPLE,
                                       adaptation_rate=ad.adaptation_rate, enabled=True)
                else:
                    adapt = make_adapt(mode=mode, num_adaptation_steps=ad.num_adaptation_steps,
                                       target_accept_prob=ad.target_accept_prob,
                                       exploration_shrinkage=ad.exploration_shrinkage,
                                       step_count_smoothing=ad.step_count_smoothing, decay_rate=ad.decay_rate, enabled=True)
                blocks.append(dict(kind="nuts" if sampler._name.startswith("nuts") else "hmc", step_size=eps0,
                                   dim_scale=mask if scale is None else mask * scale, adapt=adapt,
                                   max_tree_depth=getattr(spec, "max_tree_depth", 10),
                                   max_energy_diff=getattr(spec, "max_energy_diff", 1000.0),
                                   num_leapfrog_steps=getattr(spec, "num_leapfrog_steps", 0)))
            elif sampler._name == "rwm":
                spec = mk.kernel(target_log_prob_fn=None,
                                 **{k: v for k, v in kk.items() if k in _signature_keys(mk.kernel.__init__, 1)})
                rw_kind, rw_scale = None, spec.scale
                if spec.proposal is not None:
                    rw_kind = np.zeros(ir.D, dtype=np.int32)
                    rw_kind[mask > 0] = spec.proposal.device_code()
                    rw_scale = spec.proposal.device_scale()
                blocks.append(dict(kind="rwm", rw_scale=rw_scale, rw_kind=rw_kind, dim_scale=mask))
            else:
                raise NotImplementedError("sampler `{}` cannot be a block of the compound step".format(sampler._name))
        return blocks

    def _run_chains(self, init, burn_in):
        self._check_chain_kwargs()
        dm = self._device_model
        theta0 = self._pack_init(init)
        self._num_chains = theta0.shape[0]
        out = dm.compound(self._blocks(burn_in), theta0, self._draws_to_run(), burn_in, seed=self._seed_value(),
                          dtype=self._dtype, warps_per_block=self.backend_kwargs.get("warps_per_block", 0),
                          chain_offset=self.backend_kwargs.get("chain_offset", 0))
        k = self._thin()
        if k:
            out["states"], out["lp"], out["log_accept"] = out["states"][k::k + 1], out["lp"][k::k + 1], out["log_accept"][:, k::k + 1]
        self._run_info = dict(kernel="compound", step_size=out["step_size"], blocks=self._block_vars)
        self._last_states = out["states"]
        return self._unpack_states(out["states"]), self.trace_fn(out["states"], out)


# ---------------------------------------------------------------------------
# log-density builder (reference :729-829)
# ---------------------------------------------------------------------------
_device_models: "OrderedDict[Any, Any]" = OrderedDict()
MAX_DEVICE_MODELS = 4  # every entry holds device copies of the data pools (and GLM / GP workspaces)


def clear_device_models():
    """Release the device copies (data pools, workspaces) of every cached model."""
    while _device_models:
        _, dm = _device_models.popitem(last=False)
        dm.close()


def _get_device_model(ir: irmod.ModelIR, dtype, device):
    from .._cabi import DeviceModel

    f64 = np.dtype(dtype) == np.float64
    dev = 0 if device is None else int(device)
    key = (ir.struct_hash, dev, f64)
    dm = _device_models.get(key)
    if dm is not None:
        if dm.ir.layout_key() != ir.layout_key():
            # same model structure, data of another shape (or other blob offsets / plan sizes): a fresh device model
            del _device_models[key]
            dm.close()
            dm = None
        elif dm.ir is not ir and not (np.array_equal(dm.ir.dataf, ir.dataf) and np.array_equal(dm.ir.datai, ir.datai)):
            try:
                dm.set_data(ir)
            except Exception:
                del _device_models[key]
                dm.close()
                dm = None
    if dm is None:
        dm = DeviceModel(ir, device=dev, f64=f64)
    _device_models[key] = dm
    _device_models.move_to_end(key)
    while len(_device_models) > MAX_DEVICE_MODELS:  # least recently used first
        _, old = _device_models.popitem(last=False)
        old.close()
    return dm


def build_logp_and_deterministic_functions(
    model,
    num_chains: Optional[int] = None,
    observed: Optional[dict] = None,
    state: Optional[flow.SamplingState] = None,
    collect_reduced_log_prob: bool = True,
    parent_inds: Optional[List] = None,
    dtype="float32",
    device: Optional[int] = None,
):
    """Lower ``model`` and return ``(logpfn, initial values, deterministics_callback,
    deterministic_names, sampling state)`` like the reference; both callables run on the GPU.

    ``logpfn(*values)`` / ``logpfn(**{name: value})`` accepts values with any number of leading
    batch axes and returns a device tensor of that batch shape.
    """
    if not isinstance(model, Model):
        raise TypeError(
            "`sample` function only supports `pymc4.Model` objects, but you've passed `{}`".format(type(model))
        )
    if state is not None and observed is not None:
        raise ValueError("Can't use both `state` and `observed` arguments")
    ir, init_state, deterministic_names = irmod.lower_model(model, observed=observed, state=state)
    unobserved_keys = list(init_state.all_unobserved_values.keys())
    # `parent_inds` (compound step): the reference reorders the state parts so that the variables of one sub-kernel
    # are adjacent; the device state keeps the model's order and the blocks are masks over its dimensions
    dtype = np.dtype(dtype)
    holder: Dict[str, Any] = {}

    def device_model():
        if "dm" not in holder:
            holder["dm"] = _get_device_model(ir, dtype, device)
        return holder["dm"]

    def _pack(values, kwargs):
        if kwargs and values:
            raise TypeError("Either list state should be passed or a dict one")
        if values:
            kwargs = dict(zip(unobserved_keys, values))
        parts, lead = [], None
        for rv in ir.rvs:
            v = kwargs[rv.name]
            v = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            nd = len(rv.shape)
            this_lead = v.shape[: v.ndim - nd] if nd else v.shape
            lead = this_lead if lead is None else np.broadcast_shapes(lead, this_lead)
            parts.append((rv, v))
        flat = np.empty(tuple(lead) + (ir.D,), dtype=dtype)
        for rv, v in parts:
            flat[..., rv.offset: rv.offset + rv.size] = np.broadcast_to(v, tuple(lead) + rv.shape).reshape(
                tuple(lead) + (rv.size,))
        return flat, tuple(lead)

    def logpfn(*values, **kwargs):
        flat, lead = _pack(values, kwargs)
        lp, _ = device_model().logp_grad(flat.reshape(-1, ir.D), dtype=dtype, want_grad=False)
        return lp.reshape(lead)

    def logp_and_grad(*values, **kwargs):
        flat, lead = _pack(values, kwargs)
        lp, grad = device_model().logp_grad(flat.reshape(-1, ir.D), dtype=dtype, want_grad=True)
        return lp.reshape(lead), grad.reshape(lead + (ir.D,))

    def deterministics_callback(*values, **kwargs):
        flat, lead = _pack(values, kwargs)
        out = device_model().deterministics(flat.reshape(-1, ir.D), dtype=dtype)
        return [out[:, d.out_offset: d.out_offset + d.n].reshape(lead + d.shape) for d in ir.dets]

    class _Fn:
        """callable with the lowered model attached (``.ir``, ``.device_model``)."""

        def __init__(self, fn):
            self._fn = fn
            self.ir = ir
            self.value_and_gradient = logp_and_grad

        def __call__(self, *a, **k):
            return self._fn(*a, **k)

        @property
        def device_model(self):
            return device_model()

    init = {k: np.asarray(v, dtype=dtype) for k, v in init_state.all_unobserved_values.items()}
    return (_Fn(logpfn), init, _Fn(deterministics_callback), deterministic_names, init_state)


def vectorize_logp_function(logpfn):
    """The device functions already broadcast over leading (chain / draw) axes."""

    def vectorized_logpfn(*state):
        return logpfn(*state)

    for attr in ("ir", "device_model", "value_and_gradient"):
        if hasattr(logpfn, attr):
            try:
                setattr(vectorized_logpfn, attr, getattr(logpfn, attr))
            except Exception:  # device_model may be unavailable on a CPU-only host
                pass
    return vectorized_logpfn


def tile_init(init, num_repeats):
    return [np.tile(np.expand_dims(np.asarray(t), 0), [num_repeats] + [1] * np.ndim(t)) for t in init]


def _log_likelihood_from_states(dm, states, dtype, ir=None) -> Dict[str, Any]:
    """states [S, C, D] (device) -> {observed name: [S, C, *shape]} device tensors: one kernel evaluates
    every observed term element-wise for every state (no reduction)."""
    ir = dm.ir if ir is None else ir  # labels of the caller's model (the device model may be shared, see _BaseSampler._ir)
    ll = dm.log_likelihood(states, dtype=dtype)
    lead = tuple(ll.shape[:-1])
    out = {}
    for t in ir.terms:
        if t.observed:
            out[t.label] = ll[..., t.ll_offset: t.ll_offset + t.n].reshape(lead + tuple(t.shape))
    return out


def calculate_log_likelihood(model: Model, posterior: Dict[str, Any], sampling_state: flow.SamplingState,
                             dtype="float32", device: Optional[int] = None) -> Dict[str, Any]:
    """Pointwise log-likelihood ``dist.log_prob(observed)`` of every observed variable for every draw
    and chain of ``posterior`` (``{sampled name: [draw, chain, *core]}``), on the GPU.

    Reference: pymc4/mcmc/samplers.py:832-854 (two nested ``tf.vectorized_map`` over chains and draws
    of the transformed executor); here the draws are packed into states ``[S, C, D]`` and one kernel
    evaluates the observed terms of the lowered model without the reduction."""
    ir, _, _ = irmod.lower_model(model, observed=dict(sampling_state.observed_values) or None)
    dm = _get_device_model(ir, np.dtype(dtype), device)
    torch = dm.torch
    first = next(iter(posterior.values()))
    S, C = int(first.shape[0]), int(first.shape[1])
    theta = torch.empty((S, C, ir.D), dtype=torch.float64 if np.dtype(dtype) == np.float64 else torch.float32,
                        device=dm._dev())
    for rv in ir.rvs:
        part = posterior[rv.name]
        part = part if isinstance(part, torch.Tensor) else torch.as_tensor(np.asarray(part))
        theta[:, :, rv.offset: rv.offset + rv.size] = part.to(theta.device, theta.dtype).reshape(S, C, rv.size)
    return _log_likelihood_from_states(dm, theta, np.dtype(dtype), ir=ir)
