"""ctypes front-end of the CPU oracle (oracle/pm4_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  The product package
never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

from pymc4_b200._cstructs import CompoundBlock, HMCCfg, NUTSCfg, PM4_F32, PM4_F64
from pymc4_b200.ir import CIR, ModelIR, PackedIR

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpm4oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, -O2, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("pm4_oracle.c", "pm4_oracle_impl.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "pm4b200.h"))
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libpm4oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        irp = ctypes.POINTER(CIR)
        L.pm4o_logp_grad.argtypes = [irp, i32, i32, vp, vp, vp]
        L.pm4o_deterministics.argtypes = [irp, i32, i64, vp, vp]
        L.pm4o_log_likelihood.argtypes = [irp, i32, i64, i32, vp, vp]
        L.pm4o_posterior_predictive.argtypes = [irp, i32, i64, i32, ctypes.c_uint64, vp, vp]
        L.pm4o_gp_predictive.argtypes = [irp, i32, i64, ctypes.c_uint64, vp, vp]
        L.pm4o_prior_predictive.argtypes = [irp, i32, i64, i32, ctypes.c_uint64, vp, vp]
        L.pm4o_nuts_run.argtypes = [irp, i32, ctypes.POINTER(NUTSCfg)] + [vp] * 13
        L.pm4o_hmc_run.argtypes = [irp, i32, ctypes.POINTER(HMCCfg)] + [vp] * 10
        L.pm4o_compound_run.argtypes = [irp, i32, i32, i32, i32, ctypes.c_uint64, i32, ctypes.POINTER(CompoundBlock), i32] + [vp] * 6
        L.pm4o_philox4x32_10.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32,
                                         ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
        L.pm4o_philox4x32_10.restype = None
        L.pm4o_draw_momentum.argtypes = [i32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, i32, vp]
        L.pm4o_draw_momentum.restype = None
        _lib = L
    return _lib


def _np_dtype(dtype):
    return np.float64 if dtype == PM4_F64 else np.float32


def _code(dtype) -> int:
    return PM4_F64 if np.dtype(dtype) == np.float64 else PM4_F32


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("oracle {} failed with code {}".format(what, rc))


def set_threads(n: int) -> int:
    """OpenMP threads of the chain loops (overrides an inherited OMP_NUM_THREADS); returns the count in effect."""
    try:  # a BLAS initialised under OMP_NUM_THREADS=1 may have pinned the process to one core: undo it, the
        os.sched_setaffinity(0, range(os.cpu_count() or 1))  # OpenMP workers inherit the mask when they are created
    except (AttributeError, OSError):
        pass
    lib().pm4o_set_num_threads(int(n))
    return int(lib().pm4o_max_threads())


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*[int(v) & 0xFFFFFFFF for v in ctr])
    out = (ctypes.c_uint32 * 4)()
    lib().pm4o_philox4x32_10(c, int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF, out)
    return [int(v) for v in out]


def draw_momentum(seed, chain, t, D, dtype=np.float32):
    out = np.empty(D, dtype=dtype)
    lib().pm4o_draw_momentum(_code(dtype), int(seed), int(chain), int(t), int(D), _ptr(out))
    return out


class Oracle:
    """The oracle bound to one lowered model."""

    def __init__(self, ir: ModelIR):
        self.ir = ir
        self.packed = PackedIR(ir)

    def logp_grad(self, theta, dtype=np.float64, want_grad=True):
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=dtype)
        C, D = theta.shape
        assert D == self.ir.D
        lp = np.empty(C, dtype=dtype)
        grad = np.empty((C, D), dtype=dtype) if want_grad else None
        _check(lib().pm4o_logp_grad(self.packed.byref(), _code(dtype), C, _ptr(theta), _ptr(lp), _ptr(grad)),
               "logp_grad")
        return lp, grad

    def deterministics(self, theta, dtype=np.float64):
        theta = np.ascontiguousarray(theta, dtype=dtype).reshape(-1, self.ir.D)
        out = np.empty((theta.shape[0], self.ir.det_row), dtype=dtype)
        _check(lib().pm4o_deterministics(self.packed.byref(), _code(dtype), theta.shape[0], _ptr(theta), _ptr(out)),
               "deterministics")
        return out

    def log_likelihood(self, theta, dtype=np.float64):
        """theta [..., D] -> [..., ll_row]: pointwise log-likelihood of the observed variables."""
        theta = np.ascontiguousarray(theta, dtype=dtype)
        lead = theta.shape[:-1]
        flat = theta.reshape(-1, self.ir.D)
        out = np.empty((flat.shape[0], self.ir.ll_row), dtype=dtype)
        _check(lib().pm4o_log_likelihood(self.packed.byref(), _code(dtype), flat.shape[0], self.ir.ll_row,
                                         _ptr(flat), _ptr(out)), "log_likelihood")
        return out.reshape(lead + (self.ir.ll_row,))

    def posterior_predictive(self, theta, seed=0, dtype=np.float64):
        """theta [..., D] -> [..., ll_row]: one draw of every observed variable per state."""
        theta = np.ascontiguousarray(theta, dtype=dtype)
        lead = theta.shape[:-1]
        flat = theta.reshape(-1, self.ir.D)
        out = np.empty((flat.shape[0], self.ir.ll_row), dtype=dtype)
        _check(lib().pm4o_posterior_predictive(self.packed.byref(), _code(dtype), flat.shape[0], self.ir.ll_row,
                                               int(seed) & (2**64 - 1), _ptr(flat), _ptr(out)), "posterior_predictive")
        return out.reshape(lead + (self.ir.ll_row,))

    def gp_predictive(self, theta, seed=0, which=0):
        """theta [..., D] -> [..., n_points]: one draw of the observed vector of GP term ``which`` per state."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        lead = theta.shape[:-1]
        flat = theta.reshape(-1, self.ir.D)
        n = self.ir.gp_terms[which].gp.n
        out = np.empty((flat.shape[0], n), dtype=np.float64)
        _check(lib().pm4o_gp_predictive(self.packed.byref(), int(which), flat.shape[0], int(seed) & (2**64 - 1),
                                        _ptr(flat), _ptr(out)), "gp_predictive")
        return out.reshape(lead + (n,))

    def prior_predictive(self, M, seed=0, dtype=np.float64):
        """M ancestral samples of the model: (x [M, D] constrained values of the free variables,
        obs [M, ll_row] draws of the observed variables)."""
        x = np.empty((int(M), self.ir.D), dtype=dtype)
        obs = np.empty((int(M), self.ir.ll_row), dtype=dtype)
        _check(lib().pm4o_prior_predictive(self.packed.byref(), _code(dtype), int(M), self.ir.ll_row,
                                           int(seed) & (2**64 - 1), _ptr(x), _ptr(obs)), "prior_predictive")
        return x, obs

    def nuts(self, cfg: NUTSCfg, theta0, dtype=np.float32, step_scale=None, momenta=None, step_sizes=None):
        C, S, D = cfg.C, cfg.num_results, self.ir.D
        eps_c = None if step_sizes is None else np.ascontiguousarray(np.asarray(step_sizes).reshape(C), dtype=dtype)
        cfg.step_size_dev = eps_c.ctypes.data if eps_c is not None else None  # a HOST pointer for the oracle
        theta0 = np.ascontiguousarray(np.broadcast_to(np.asarray(theta0, dtype=dtype), (C, D)))
        step_scale = None if step_scale is None else np.ascontiguousarray(step_scale, dtype=dtype)
        momenta = None if momenta is None else np.ascontiguousarray(momenta, dtype=dtype)
        out = dict(
            states=np.empty((S, C, D), dtype=dtype), lp=np.empty((S, C), dtype=dtype),
            leapfrogs=np.empty((S, C), dtype=np.int32), diverging=np.empty((S, C), dtype=np.uint8),
            energy=np.empty((S, C), dtype=dtype), log_accept=np.empty((S, C), dtype=dtype),
            depth=np.empty((S, C), dtype=np.int32), step_size=np.empty(C, dtype=dtype),
            total_leapfrogs=np.empty(C, dtype=np.int64),
            mass_scale=np.ones(D, dtype=dtype),
        )
        rc = lib().pm4o_nuts_run(
            self.packed.byref(), _code(dtype), ctypes.byref(cfg), _ptr(theta0), _ptr(step_scale), _ptr(momenta),
            _ptr(out["states"]), _ptr(out["lp"]), _ptr(out["leapfrogs"]), _ptr(out["diverging"]),
            _ptr(out["energy"]), _ptr(out["log_accept"]), _ptr(out["depth"]), _ptr(out["step_size"]),
            _ptr(out["total_leapfrogs"]), _ptr(out["mass_scale"]))
        _check(rc, "nuts_run")
        return out

    def hmc(self, cfg: HMCCfg, theta0, dtype=np.float32, step_scale=None, momenta=None, uniforms=None, step_sizes=None):
        C, S, D, T = cfg.C, cfg.num_results, self.ir.D, cfg.num_results + cfg.num_burnin
        eps_c = None if step_sizes is None else np.ascontiguousarray(np.asarray(step_sizes).reshape(C), dtype=dtype)
        cfg.step_size_dev = eps_c.ctypes.data if eps_c is not None else None  # a HOST pointer for the oracle
        theta0 = np.ascontiguousarray(np.broadcast_to(np.asarray(theta0, dtype=dtype), (C, D)))
        step_scale = None if step_scale is None else np.ascontiguousarray(step_scale, dtype=dtype)
        momenta = None if momenta is None else np.ascontiguousarray(momenta, dtype=dtype)
        uniforms = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=dtype)
        out = dict(
            states=np.empty((S, C, D), dtype=dtype), lp=np.empty((S, C), dtype=dtype),
            log_accept=np.empty((S, C), dtype=dtype), accepted=np.empty((S, C), dtype=np.uint8),
            step_size=np.empty(C, dtype=dtype), traj=np.empty((T, C, D), dtype=dtype),
        )
        rc = lib().pm4o_hmc_run(
            self.packed.byref(), _code(dtype), ctypes.byref(cfg), _ptr(theta0), _ptr(step_scale), _ptr(momenta),
            _ptr(uniforms), _ptr(out["states"]), _ptr(out["lp"]), _ptr(out["log_accept"]),
            _ptr(out["accepted"]), _ptr(out["step_size"]), _ptr(out["traj"]))
        _check(rc, "hmc_run")
        return out


def _compound(self, blocks, theta0, num_results, num_burnin, seed=0, dtype=np.float64, chain_offset=0):
    """Oracle of ``pm4_compound_run``: ``blocks`` as for ``DeviceModel.compound`` (host arrays)."""
    from pymc4_b200._cstructs import PM4_BLOCK_HMC, PM4_BLOCK_NUTS, PM4_BLOCK_RWM, make_adapt

    theta0 = np.ascontiguousarray(theta0, dtype=dtype)
    C, D, S = theta0.shape[0], self.ir.D, int(num_results)
    arr = (CompoundBlock * len(blocks))()
    keep = []
    codes = {"nuts": PM4_BLOCK_NUTS, "hmc": PM4_BLOCK_HMC, "rwm": PM4_BLOCK_RWM}
    for k, b in enumerate(blocks):
        scale = np.ascontiguousarray(np.asarray(b["dim_scale"], dtype=np.float64).reshape(D), dtype=dtype)
        kinds = None if b.get("rw_kind") is None else np.ascontiguousarray(b["rw_kind"], dtype=np.int32).reshape(D)
        keep += [scale, kinds]
        arr[k] = CompoundBlock(codes[b["kind"]], int(b.get("num_leapfrog_steps", 0)), int(b.get("max_tree_depth", 10)), 0,
                               float(b.get("max_energy_diff", 1000.0)), float(b.get("step_size", 1.0)),
                               float(b.get("rw_scale", 1.0)), b.get("adapt") or make_adapt(enabled=False),
                               scale.ctypes.data, kinds.ctypes.data if kinds is not None else None)
    out = dict(states=np.empty((S, C, D), dtype=dtype), lp=np.empty((S, C), dtype=dtype),
               step_size=np.empty((len(blocks), C), dtype=dtype), log_accept=np.empty((len(blocks), S, C), dtype=dtype),
               leapfrogs=np.zeros((len(blocks), S, C), dtype=np.int32))
    rc = lib().pm4o_compound_run(self.packed.byref(), _code(dtype), C, S, int(num_burnin), ctypes.c_uint64(int(seed) & (2**64 - 1)),
                                 int(chain_offset), arr, len(blocks), _ptr(theta0), _ptr(out["states"]), _ptr(out["lp"]),
                                 _ptr(out["step_size"]), _ptr(out["log_accept"]), _ptr(out["leapfrogs"]))
    _check(rc, "compound_run")
    return out


Oracle.compound = _compound
