"""Parameter carriers with the constructor signatures of the TFP kernels the reference binds.

The reference routes ``pm.sample(**kwargs)`` to the kernel, the step-size adaptation and
``sample_chain`` by *introspecting their signatures* (pymc4/mcmc/samplers.py:191-207).  These
classes keep that routing intact -- same parameter names and defaults as
``tfp.mcmc.NoUTurnSampler`` / ``HamiltonianMonteCarlo`` /
``DualAveragingStepSizeAdaptation`` / ``sample_chain`` (TFP ~0.12) -- but they
only hold configuration: the transition itself runs in the CUDA kernels
(csrc/pm4_kernels.cuh).
"""
from typing import Any, Callable, Optional


class NoUTurnSampler:
    def __init__(self, target_log_prob_fn, step_size, max_tree_depth=10, max_energy_diff=1000.0,
                 unrolled_leapfrog_steps=1, parallel_iterations=10, seed=None, name=None):
        if int(unrolled_leapfrog_steps) != 1:
            raise NotImplementedError("unrolled_leapfrog_steps != 1 is not supported by the CUDA NUTS kernel")
        self.target_log_prob_fn = target_log_prob_fn
        self.step_size = step_size
        self.max_tree_depth = int(max_tree_depth)
        self.max_energy_diff = float(max_energy_diff)
        self.unrolled_leapfrog_steps = 1
        self.parallel_iterations = parallel_iterations
        self.seed = seed
        self.name = name

    is_calibrated = True

    @property
    def parameters(self):
        return dict(step_size=self.step_size, max_tree_depth=self.max_tree_depth,
                    max_energy_diff=self.max_energy_diff)


class HamiltonianMonteCarlo:
    def __init__(self, target_log_prob_fn, step_size, num_leapfrog_steps, state_gradients_are_stopped=False,
                 step_size_update_fn=None, seed=None, store_parameters_in_results=False, name=None):
        if step_size_update_fn is not None:
            raise NotImplementedError("step_size_update_fn is not supported; use the dual-averaging adaptation")
        self.target_log_prob_fn = target_log_prob_fn
        self.step_size = step_size
        self.num_leapfrog_steps = int(num_leapfrog_steps)
        self.seed = seed
        self.name = name

    is_calibrated = True

    @property
    def parameters(self):
        return dict(step_size=self.step_size, num_leapfrog_steps=self.num_leapfrog_steps)


class DualAveragingStepSizeAdaptation:
    def __init__(self, inner_kernel, num_adaptation_steps, target_accept_prob=0.75, exploration_shrinkage=0.05,
                 shrinkage_target=None, step_count_smoothing=10, decay_rate=0.75,
                 step_size_setter_fn: Optional[Callable] = None, step_size_getter_fn: Optional[Callable] = None,
                 log_accept_prob_getter_fn: Optional[Callable] = None, reduce_fn: Any = None,
                 validate_args=False, name=None):
        self.shrinkage_target = None if shrinkage_target is None else float(shrinkage_target)
        if self.shrinkage_target is not None and not self.shrinkage_target > 0:
            raise ValueError("shrinkage_target must be positive")
        self.inner_kernel = inner_kernel
        self.num_adaptation_steps = int(num_adaptation_steps)
        self.target_accept_prob = float(target_accept_prob)
        self.exploration_shrinkage = float(exploration_shrinkage)
        self.step_count_smoothing = float(step_count_smoothing)
        self.decay_rate = float(decay_rate)
        self.name = name

    is_calibrated = True


class SimpleStepSizeAdaptation:
    """tfp.mcmc.SimpleStepSizeAdaptation: step *= (1 + adaptation_rate) when the (mean) acceptance
    probability exceeds the target, /= otherwise, for the first ``num_adaptation_steps`` transitions."""

    def __init__(self, inner_kernel, num_adaptation_steps, target_accept_prob=0.75, adaptation_rate=0.01,
                 step_size_setter_fn: Optional[Callable] = None, step_size_getter_fn: Optional[Callable] = None,
                 log_accept_prob_getter_fn: Optional[Callable] = None, reduce_fn: Any = None,
                 validate_args=False, name=None):
        self.inner_kernel = inner_kernel
        self.num_adaptation_steps = int(num_adaptation_steps)
        self.target_accept_prob = float(target_accept_prob)
        self.adaptation_rate = float(adaptation_rate)
        self.name = name

    is_calibrated = True


class RandomWalkMetropolis:
    """tfp.mcmc.RandomWalkMetropolis: proposal = state + perturbation, Metropolis-Hastings accept.  Only the
    default ``new_state_fn`` (``random_walk_normal_fn(scale=1.0)``) runs on the device; ``scale`` (a backend
    extension) sets its standard deviation."""

    def __init__(self, target_log_prob_fn, new_state_fn=None, seed=None, name=None, scale=1.0):
        # besides the default, the proposals of pymc4_b200.distributions.state_functions run on the device (compound step)
        proposal = getattr(new_state_fn, "__self__", None)
        if new_state_fn is not None and not hasattr(proposal, "device_code"):
            raise NotImplementedError("custom new_state_fn is not supported; the device proposals are "
                                      "state + scale * Normal(0, 1) (tfp random_walk_normal_fn) and pm.bernoulli_fn / "
                                      "pm.categorical_uniform_fn / pm.gaussian_round_fn")
        self.new_state_fn = new_state_fn
        self.proposal = proposal
        self.target_log_prob_fn = target_log_prob_fn
        self.scale = float(scale)
        self.seed = seed
        self.name = name

    is_calibrated = True

    @property
    def parameters(self):
        return dict(scale=self.scale)


class _CompoundStepTF:
    """Configuration carrier of the compound transition kernel (reference pymc4/mcmc/tf_support.py:52-163): one
    transition of every sub-kernel in turn, each on its subset of the free variables with the others held fixed.
    The transition runs in ``pm4_compound_run`` (csrc/pm4_core.cu)."""

    def __init__(self, target_log_prob_fn, compound_samplers, compound_set_lengths, name=None):
        self.target_log_prob_fn = target_log_prob_fn
        self.compound_samplers = list(compound_samplers)
        self.compound_set_lengths = list(compound_set_lengths)
        self.name = name

    is_calibrated = True


def sample_chain(num_results, current_state, previous_kernel_results=None, kernel=None, num_burnin_steps=0,
                 num_steps_between_results=0, trace_fn=None, return_final_kernel_results=False,
                 parallel_iterations=10, seed=None, name=None):
    """Signature carrier only (see module docstring); the chain loop runs on the device."""
    raise NotImplementedError("sample_chain is executed by libpm4b200 (pm4_nuts_run / pm4_hmc_run)")
