"""ctypes mirrors of the configuration structs in include/pm4b200.h."""
import ctypes

PM4_F32, PM4_F64 = 0, 1
PM4_EPS_POOLED, PM4_EPS_PER_CHAIN = 0, 1
PM4_ADAPT_DUAL_AVERAGING, PM4_ADAPT_SIMPLE = 0, 1
PM4_PROPOSAL_LEAPFROG, PM4_PROPOSAL_RANDOM_WALK = 0, 1
PM4_BLOCK_NUTS, PM4_BLOCK_HMC, PM4_BLOCK_RWM = 0, 1, 2


class AdaptCfg(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int32),
        ("num_adaptation_steps", ctypes.c_int32),
        ("target_accept_prob", ctypes.c_double),
        ("exploration_shrinkage", ctypes.c_double),
        ("step_count_smoothing", ctypes.c_double),
        ("decay_rate", ctypes.c_double),
        ("enabled", ctypes.c_int32),
        ("kind", ctypes.c_int32),
        ("adaptation_rate", ctypes.c_double),
        ("mass_windows", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("shrinkage_target", ctypes.c_double),
    ]


class HMCCfg(ctypes.Structure):
    _fields_ = [
        ("C", ctypes.c_int32),
        ("num_results", ctypes.c_int32),
        ("num_burnin", ctypes.c_int32),
        ("num_leapfrog_steps", ctypes.c_int32),
        ("step_size", ctypes.c_double),
        ("seed", ctypes.c_uint64),
        ("adapt", AdaptCfg),
        ("warps_per_block", ctypes.c_int32),
        ("chain_offset", ctypes.c_int32),
        ("proposal", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("rw_scale", ctypes.c_double),
        ("step_size_dev", ctypes.c_void_p),
    ]


class NUTSCfg(ctypes.Structure):
    _fields_ = [
        ("C", ctypes.c_int32),
        ("num_results", ctypes.c_int32),
        ("num_burnin", ctypes.c_int32),
        ("max_tree_depth", ctypes.c_int32),
        ("max_energy_diff", ctypes.c_double),
        ("step_size", ctypes.c_double),
        ("seed", ctypes.c_uint64),
        ("adapt", AdaptCfg),
        ("warps_per_block", ctypes.c_int32),
        ("chain_offset", ctypes.c_int32),
        ("step_size_dev", ctypes.c_void_p),
    ]


class CompoundBlock(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32),
        ("num_leapfrog_steps", ctypes.c_int32),
        ("max_tree_depth", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("max_energy_diff", ctypes.c_double),
        ("step_size", ctypes.c_double),
        ("rw_scale", ctypes.c_double),
        ("adapt", AdaptCfg),
        ("dim_scale_dev", ctypes.c_void_p),
        ("rw_kind_dev", ctypes.c_void_p),
    ]


def make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=0, target_accept_prob=0.75,
               exploration_shrinkage=0.05, step_count_smoothing=10.0, decay_rate=0.75,
               enabled=True, kind=PM4_ADAPT_DUAL_AVERAGING, adaptation_rate=0.01, mass_windows=0,
               shrinkage_target=None) -> AdaptCfg:
    return AdaptCfg(int(mode), int(num_adaptation_steps), float(target_accept_prob),
                    float(exploration_shrinkage), float(step_count_smoothing), float(decay_rate),
                    int(bool(enabled)), int(kind), float(adaptation_rate), int(mass_windows), 0,
                    float(shrinkage_target) if shrinkage_target else 0.0)


def make_nuts_cfg(C, num_results, num_burnin, step_size=0.1, seed=0, max_tree_depth=10,
                  max_energy_diff=1000.0, adapt=None, warps_per_block=0, chain_offset=0) -> NUTSCfg:
    adapt = adapt if adapt is not None else make_adapt(num_adaptation_steps=num_burnin)
    return NUTSCfg(int(C), int(num_results), int(num_burnin), int(max_tree_depth),
                   float(max_energy_diff), float(step_size), int(seed) & (2**64 - 1), adapt,
                   int(warps_per_block), int(chain_offset), None)


def make_hmc_cfg(C, num_results, num_burnin, step_size=0.1, num_leapfrog_steps=3, seed=0,
                 adapt=None, warps_per_block=0, chain_offset=0, proposal=PM4_PROPOSAL_LEAPFROG,
                 rw_scale=1.0) -> HMCCfg:
    adapt = adapt if adapt is not None else make_adapt(num_adaptation_steps=num_burnin)
    return HMCCfg(int(C), int(num_results), int(num_burnin), int(num_leapfrog_steps),
                  float(step_size), int(seed) & (2**64 - 1), adapt, int(warps_per_block),
                  int(chain_offset), int(proposal), 0, float(rw_scale), None)
