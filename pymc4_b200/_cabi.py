"""ctypes binding of libpm4b200.so (include/pm4b200.h) + torch hand-off of device buffers.

There is no CPU fallback: importing this module is cheap, but any compute call raises
``BackendUnavailable`` when the CUDA library cannot be built/loaded or no CUDA device is visible.
PyTorch is used only to allocate device memory, to pass ``data_ptr()`` / the current stream, and
to copy results to the host.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from typing import Dict, Optional

import numpy as np

from . import codegen
from ._cstructs import (CompoundBlock, HMCCfg, NUTSCfg, PM4_BLOCK_HMC, PM4_BLOCK_NUTS, PM4_BLOCK_RWM, PM4_F32, PM4_F64,
                       make_adapt)
from .ir import CIR, ModelIR, PackedIR

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpm4b200.so")
_CORE_SRC = os.path.join(_HERE, "csrc", "pm4_core.cu")
_GLM_SRC = os.path.join(_HERE, "csrc", "pm4_glm.cu")
_GP_SRC = os.path.join(_HERE, "csrc", "pm4_gp.cu")
_HEADER = os.path.join(_HERE, "..", "include", "pm4b200.h")
_lock = threading.Lock()
_lib = None

EXPORTS = [
    "pm4_model_create", "pm4_model_destroy", "pm4_model_dim", "pm4_model_set_data", "pm4_model_set_plans",
    "pm4_logp_grad",
    "pm4_deterministics", "pm4_log_likelihood", "pm4_model_ll_row", "pm4_posterior_predictive", "pm4_prior_predictive", "pm4_hmc_run", "pm4_nuts_run", "pm4_compound_run", "pm4_scratch_bytes", "pm4_launch_count",
    "pm4_last_error", "pm4_abi_version",
    "pm4_comm_create", "pm4_comm_connect", "pm4_comm_destroy", "pm4_model_attach_comm", "pm4_model_set_trace_sink", "pm4_model_get_mass_scale",
    "pm4_glm_work_bytes", "pm4_glm_logp_grad", "pm4_glm_gemm_tile",
    "pm4_glm_image_bytes", "pm4_glm_build_image", "pm4_glm_fused_work_bytes", "pm4_glm_fused_logp_grad",
    "pm4_gp_work_bytes", "pm4_gp_logp_grad", "pm4_gp_launches", "pm4_gp_sample", "pm4_gp_predictive", "pm4_gp_points",
]


class BackendUnavailable(RuntimeError):
    """The CUDA backend cannot run here (no device, or the library failed to build/load)."""


def build_core(force: bool = False) -> str:
    """Compile libpm4b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    deps = [_CORE_SRC, _GLM_SRC, _GP_SRC, _HEADER, os.path.join(_HERE, "csrc", "pm4_module.h"),
            os.path.join(_HERE, "csrc", "pm4_lockstep.cuh"), os.path.join(_HERE, "csrc", "pm4_philox.cuh")]
    with _lock:
        fresh = os.path.exists(LIB_PATH) and all(os.path.getmtime(d) <= os.path.getmtime(LIB_PATH) for d in deps)
        if fresh and not force:
            return LIB_PATH
        cmd = [codegen.find_nvcc()] + codegen.NVCC_FLAGS + ["-shared", "-o", LIB_PATH + ".tmp", _CORE_SRC, _GLM_SRC, _GP_SRC, "-ldl"]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise BackendUnavailable("nvcc failed building libpm4b200.so:\n" + proc.stdout[-4000:])
        os.replace(LIB_PATH + ".tmp", LIB_PATH)
        return LIB_PATH


def load_library():
    """dlopen the C-ABI library (building it first if the source is newer)."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        path = build_core()
    except RuntimeError as e:
        if not os.path.exists(LIB_PATH):
            raise BackendUnavailable(str(e)) from e
        path = LIB_PATH
    L = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.pm4_model_create.argtypes = [ctypes.POINTER(CIR), ctypes.c_char_p, i32, ctypes.POINTER(vp)]
    L.pm4_model_destroy.argtypes = [vp]
    L.pm4_model_dim.argtypes = [vp]
    L.pm4_model_set_data.argtypes = [vp, ctypes.POINTER(ctypes.c_double), i64, ctypes.POINTER(ctypes.c_int32), i64]
    L.pm4_model_set_plans.argtypes = [vp, ctypes.POINTER(CIR)]
    L.pm4_logp_grad.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.pm4_deterministics.argtypes = [vp, i32, i64, vp, vp, vp]
    L.pm4_log_likelihood.argtypes = [vp, i32, i64, vp, vp, vp]
    L.pm4_model_ll_row.argtypes = [vp]
    L.pm4_posterior_predictive.argtypes = [vp, i32, i64, vp, ctypes.c_uint64, vp, vp]
    L.pm4_prior_predictive.argtypes = [vp, i32, i64, ctypes.c_uint64, vp, vp, vp]
    L.pm4_gp_predictive.argtypes = [vp, i32, i32, i64, vp, ctypes.c_uint64, vp, vp]
    L.pm4_gp_points.argtypes = [vp, i32]
    L.pm4_hmc_run.argtypes = [vp, i32, ctypes.POINTER(HMCCfg)] + [vp] * 11
    L.pm4_nuts_run.argtypes = [vp, i32, ctypes.POINTER(NUTSCfg)] + [vp] * 13
    L.pm4_compound_run.argtypes = [vp, i32, i32, i32, i32, ctypes.c_uint64, i32, i32, ctypes.POINTER(CompoundBlock), i32] + [vp] * 6
    L.pm4_scratch_bytes.argtypes = [vp, i32, i32, i32]
    L.pm4_scratch_bytes.restype = i64
    L.pm4_launch_count.argtypes = [i32]
    L.pm4_launch_count.restype = i64
    L.pm4_last_error.restype = ctypes.c_char_p
    L.pm4_glm_work_bytes.argtypes = [i32, i32, i32]
    L.pm4_glm_work_bytes.restype = i64
    L.pm4_glm_logp_grad.argtypes = [vp, vp, i32, vp, i32, i32, vp, i32, i32, i32, vp, vp, i32, vp, vp, vp]
    L.pm4_glm_gemm_tile.argtypes = [vp, i32, vp, i32, vp, i32, i32, i32, vp]
    L.pm4_comm_create.argtypes = [i32, i32, i32, ctypes.POINTER(vp), ctypes.c_char_p]
    L.pm4_comm_connect.argtypes = [vp, ctypes.c_char_p]
    L.pm4_comm_destroy.argtypes = [vp]
    L.pm4_model_attach_comm.argtypes = [vp, vp]
    L.pm4_model_set_trace_sink.argtypes = [vp, vp, i32, vp]
    L.pm4_model_get_mass_scale.argtypes = [vp, i32, vp, vp]
    _lib = L
    return L


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise BackendUnavailable(
            "pymc4_b200 needs a CUDA device: there is no CPU fallback for the sampler hot path"
        )
    return torch


def launch_count(reset: bool = False) -> int:
    return int(load_library().pm4_launch_count(1 if reset else 0))


def _tdtype(torch, dtype):
    return torch.float64 if np.dtype(dtype) == np.float64 else torch.float32


def _code(dtype) -> int:
    return PM4_F64 if np.dtype(dtype) == np.float64 else PM4_F32


class DeviceModel:
    """A lowered model resident on one GPU (wraps ``pm4_model_t``)."""

    def __init__(self, ir: ModelIR, device: int = 0, f64: bool = False, module_path: Optional[str] = None):
        self.ir = ir
        self.torch = _torch()
        self.lib = load_library()
        self.device = int(device)
        self.packed = PackedIR(ir)
        self.module_path = module_path or codegen.build_module(ir, has_f64=f64)
        self.handle = ctypes.c_void_p()
        rc = self.lib.pm4_model_create(self.packed.byref(), self.module_path.encode(), self.device,
                                       ctypes.byref(self.handle))
        self._check(rc)

    # -- plumbing --------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError("libpm4b200: %s (code %d)" % (self.lib.pm4_last_error().decode(), rc))

    def close(self):
        self.__dict__.pop("_pinned_pool", None)
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.pm4_model_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _dev(self):
        return self.torch.device("cuda", self.device)

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self._dev()).cuda_stream)

    def to_device(self, arr, dtype):
        """Host array (or tensor) -> contiguous device tensor of the compute dtype."""
        torch = self.torch
        if arr is None:
            return None
        if isinstance(arr, torch.Tensor):
            return arr.to(device=self._dev(), dtype=_tdtype(torch, dtype)).contiguous()
        host = np.ascontiguousarray(arr, dtype=dtype)
        if not host.flags.writeable:  # e.g. a broadcast view: torch wants a writable buffer
            host = host.copy()
        host = torch.from_numpy(host)
        return host.to(self._dev(), non_blocking=False)

    @staticmethod
    def _p(t):
        return ctypes.c_void_p(0 if t is None else t.data_ptr())

    def _pinned(self, shape, dtype):
        """A pinned host tensor for a streamed trace.  Pinning gigabytes costs hundreds of milliseconds, so the model
        keeps the last buffer and hands it out again once nothing but the pool refers to its storage (the posterior
        arrays of a trace are NumPy views of it: while the caller holds the trace a new call gets a fresh buffer)."""
        torch = self.torch
        key = (tuple(int(v) for v in shape), dtype)
        use_count = getattr(torch._C, "_storage_Use_Count", None)
        pool = self.__dict__.setdefault("_pinned_pool", {})
        base = pool.get(key)
        if base is not None and use_count is not None:
            try:
                if use_count(base.untyped_storage()._cdata) <= 2:  # the pool's tensor + the temporary storage handle
                    return base.view(key[0])
            except Exception:
                pass
        base = torch.empty(key[0], dtype=dtype, pin_memory=True)
        pool.clear()  # one buffer at a time
        pool[key] = base
        return base.view(key[0])

    def empty(self, shape, dtype):
        return self.torch.empty(shape, dtype=dtype, device=self._dev())

    # -- API ---------------------------------------------------------------------
    def attach_comm(self, comm: Optional["Communicator"]):
        """Couple the pooled step size of this model's runs with the other ranks of ``comm`` (None detaches)."""
        self._check(self.lib.pm4_model_attach_comm(self.handle, comm.handle if comm is not None else None))
        self._comm = comm

    def set_data(self, ir: ModelIR):
        """New constant pools for the same structure (new observed values)."""
        if ir.struct_hash != self.ir.struct_hash:
            raise ValueError("set_data needs an IR with the same structure")
        if ir.layout_key() != self.ir.layout_key():
            raise ValueError("set_data needs an IR with the same pool sizes, element counts and blob offsets "
                             "(build a new DeviceModel for data of a different shape)")
        packed = PackedIR(ir)
        c = packed.c
        # plans first (the library validates them before it uploads anything), then the data pools; the Python
        # side switches to the new IR only when both calls went through
        self._check(self.lib.pm4_model_set_plans(self.handle, packed.byref()))
        self._check(self.lib.pm4_model_set_data(self.handle, c.dataf, c.n_dataf, c.datai, c.n_datai))
        self.packed = packed
        self.ir = ir

    def logp_grad(self, theta, dtype=np.float32, want_grad: bool = True):
        """theta [C, D] (host or device) -> (lp [C], grad [C, D]) device tensors."""
        torch = self.torch
        th = self.to_device(theta, dtype)
        if th.dim() == 1:
            th = th[None]
        C, D = th.shape
        if D != self.ir.D:
            raise ValueError("theta has %d columns, the model has D=%d" % (D, self.ir.D))
        td = _tdtype(torch, dtype)
        lp = self.empty((C,), td)
        grad = self.empty((C, D), td) if want_grad else None
        with torch.cuda.device(self._dev()):
            self._check(self.lib.pm4_logp_grad(self.handle, _code(dtype), C, self._p(th), self._p(lp),
                                               self._p(grad), self._stream()))
        return lp, grad

    def deterministics(self, theta, dtype=np.float32):
        """theta [..., D] device/host -> [..., det_row] device tensor."""
        torch = self.torch
        th = self.to_device(theta, dtype)
        lead = tuple(th.shape[:-1])
        flat = th.reshape(-1, self.ir.D)
        out = self.empty((flat.shape[0], self.ir.det_row), _tdtype(torch, dtype))
        with torch.cuda.device(self._dev()):
            self._check(self.lib.pm4_deterministics(self.handle, _code(dtype), flat.shape[0], self._p(flat),
                                                    self._p(out), self._stream()))
        return out.reshape(lead + (self.ir.det_row,))

    def log_likelihood(self, theta, dtype=np.float32):
        """theta [..., D] device/host -> [..., ll_row] device tensor (pointwise log-likelihood)."""
        torch = self.torch
        th = self.to_device(theta, dtype)
        lead = tuple(th.shape[:-1])
        flat = th.reshape(-1, self.ir.D)
        out = self.empty((flat.shape[0], self.ir.ll_row), _tdtype(torch, dtype))
        with torch.cuda.device(self._dev()):
            self._check(self.lib.pm4_log_likelihood(self.handle, _code(dtype), flat.shape[0], self._p(flat),
                                                    self._p(out), self._stream()))
        return out.reshape(lead + (self.ir.ll_row,))

    def posterior_predictive(self, theta, seed: int = 0, dtype=np.float32):
        """theta [..., D] device/host -> [..., ll_row] device tensor: one draw of every observed variable."""
        torch = self.torch
        th = self.to_device(theta, dtype)
        lead = tuple(th.shape[:-1])
        flat = th.reshape(-1, self.ir.D)
        out = self.empty((flat.shape[0], self.ir.ll_row), _tdtype(torch, dtype))
        with torch.cuda.device(self._dev()):
            self._check(self.lib.pm4_posterior_predictive(self.handle, _code(dtype), flat.shape[0], self._p(flat),
                                                          ctypes.c_uint64(int(seed) & (2**64 - 1)), self._p(out),
                                                          self._stream()))
        return out.reshape(lead + (self.ir.ll_row,))

    def gp_predictive(self, theta, seed: int = 0, dtype=np.float32, which: int = 0):
        """theta [..., D] (unconstrained) -> [..., n_points] device tensor: one draw of the observed vector of the
        model's GP term ``which`` per state (loc + chol(K(theta)) z, float64 factorisation)."""
        torch = self.torch
        th = self.to_device(theta, dtype)
        lead = tuple(th.shape[:-1])
        flat = th.reshape(-1, self.ir.D)
        n = self.lib.pm4_gp_points(self.handle, int(which))
        if n <= 0:
            raise ValueError("the model has no GP term %d" % which)
        out = self.empty((flat.shape[0], n), _tdtype(torch, dtype))
        with torch.cuda.device(self._dev()):
            self._check(self.lib.pm4_gp_predictive(self.handle, _code(dtype), int(which), flat.shape[0], self._p(flat),
                                                   ctypes.c_uint64(int(seed) & (2**64 - 1)), self._p(out), self._stream()))
        return out.reshape(lead + (n,))

    def prior_predictive(self, M: int, seed: int = 0, dtype=np.float32, sample_observed: bool = True):
        """M ancestral samples: (x [M, D] constrained free variables, obs [M, ll_row] or None), device tensors."""
        torch = self.torch
        td = _tdtype(torch, dtype)
        x = self.empty((int(M), self.ir.D), td)
        obs = self.empty((int(M), self.ir.ll_row), td) if sample_observed else None
        with torch.cuda.device(self._dev()):
            self._check(self.lib.pm4_prior_predictive(self.handle, _code(dtype), int(M),
                                                      ctypes.c_uint64(int(seed) & (2**64 - 1)), self._p(x),
                                                      self._p(obs), self._stream()))
        return x, obs

    def nuts(self, cfg: NUTSCfg, theta0, dtype=np.float32, step_scale=None, momenta=None,
             want_states: bool = True, stream_trace: int = 0, step_sizes=None) -> Dict[str, "object"]:
        """Run NUTS; every result is a device tensor in the reference layout [S, C, ...].

        ``stream_trace = n > 0``: the kept draws also travel to a pinned host tensor ``out["states_host"]`` in n slabs
        on a copy stream while the sampler keeps running; ``out["states_host_ready"]()`` waits for the last slab."""
        torch = self.torch
        td = _tdtype(torch, dtype)
        C, S, D = cfg.C, cfg.num_results, self.ir.D
        th0 = self.to_device(np.broadcast_to(np.asarray(theta0, dtype=dtype), (C, D))
                             if not isinstance(theta0, torch.Tensor) else theta0, dtype)
        sc = self.to_device(step_scale, dtype)
        mom = self.to_device(momenta, dtype)
        eps_c = self.to_device(None if step_sizes is None else np.asarray(step_sizes).reshape(C), dtype)  # initial epsilon per chain
        cfg.step_size_dev = eps_c.data_ptr() if eps_c is not None else None
        out = dict(
            states=self.empty((S, C, D), td) if want_states else None,
            lp=self.empty((S, C), td),
            leapfrogs=self.empty((S, C), torch.int32),
            diverging=self.empty((S, C), torch.uint8),
            energy=self.empty((S, C), td),
            log_accept=self.empty((S, C), td),
            depth=self.empty((S, C), torch.int32),
            step_size=self.empty((C,), td),
            total_leapfrogs=self.empty((C,), torch.int64),
        )
        host = copy_stream = None
        if stream_trace > 0 and want_states and S > 0 and not self.ir.lockstep:
            host = self._pinned((S, C, D), td)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self._dev())
            copy_stream = self._copy_stream
            self._check(self.lib.pm4_model_set_trace_sink(self.handle, ctypes.c_void_p(host.data_ptr()), int(stream_trace),
                                                          ctypes.c_void_p(copy_stream.cuda_stream)))
        with torch.cuda.device(self._dev()):
            rc = self.lib.pm4_nuts_run(
                self.handle, _code(dtype), ctypes.byref(cfg), self._p(th0), self._p(sc), self._p(mom),
                self._p(out["states"]), self._p(out["lp"]), self._p(out["leapfrogs"]), self._p(out["diverging"]),
                self._p(out["energy"]), self._p(out["log_accept"]), self._p(out["depth"]),
                self._p(out["step_size"]), self._p(out["total_leapfrogs"]), self._stream())
        if host is not None:
            self.lib.pm4_model_set_trace_sink(self.handle, None, 0, None)  # one-shot even when the run failed
        self._check(rc)
        if host is not None:
            out["states_host"] = host
            out["states_host_ready"] = copy_stream.synchronize
        if cfg.adapt.enabled and cfg.adapt.mass_windows > 0:  # the per-dimension step scale the run ended with
            out["mass_scale"] = self.empty((D,), td)
            with torch.cuda.device(self._dev()):
                self._check(self.lib.pm4_model_get_mass_scale(self.handle, _code(dtype), self._p(out["mass_scale"]),
                                                              self._stream()))
        out["_keepalive"] = (th0, sc, mom, eps_c)
        return out

    def hmc(self, cfg: HMCCfg, theta0, dtype=np.float32, step_scale=None, momenta=None, uniforms=None,
            want_traj: bool = False, step_sizes=None) -> Dict[str, "object"]:
        torch = self.torch
        td = _tdtype(torch, dtype)
        C, S, D, T = cfg.C, cfg.num_results, self.ir.D, cfg.num_results + cfg.num_burnin
        th0 = self.to_device(np.broadcast_to(np.asarray(theta0, dtype=dtype), (C, D))
                             if not isinstance(theta0, torch.Tensor) else theta0, dtype)
        sc = self.to_device(step_scale, dtype)
        mom = self.to_device(momenta, dtype)
        uni = self.to_device(uniforms, dtype)
        eps_c = self.to_device(None if step_sizes is None else np.asarray(step_sizes).reshape(C), dtype)
        cfg.step_size_dev = eps_c.data_ptr() if eps_c is not None else None
        out = dict(
            states=self.empty((S, C, D), td), lp=self.empty((S, C), td), log_accept=self.empty((S, C), td),
            accepted=self.empty((S, C), torch.uint8), step_size=self.empty((C,), td),
            traj=self.empty((T, C, D), td) if want_traj else None,
        )
        with torch.cuda.device(self._dev()):
            rc = self.lib.pm4_hmc_run(
                self.handle, _code(dtype), ctypes.byref(cfg), self._p(th0), self._p(sc), self._p(mom), self._p(uni),
                self._p(out["states"]), self._p(out["lp"]), self._p(out["log_accept"]), self._p(out["accepted"]),
                self._p(out["step_size"]), self._p(out["traj"]), self._stream())
        self._check(rc)
        out["_keepalive"] = (th0, sc, mom, uni, eps_c)
        return out


    def compound(self, blocks, theta0, num_results: int, num_burnin: int, seed: int = 0, dtype=np.float32,
                 warps_per_block: int = 0, chain_offset: int = 0) -> Dict[str, "object"]:
        """Compound step (``pm4_compound_run``): ``blocks`` is a list of dicts -- ``kind`` ("nuts" | "hmc" | "rwm"),
        ``dim_scale`` [D] (0 = dimension not in the block) and, by kind, ``step_size``, ``adapt``, ``max_tree_depth``,
        ``max_energy_diff``, ``num_leapfrog_steps`` / ``rw_scale``, ``rw_kind`` [D] int32 or None.  Returns device
        tensors: states [S, C, D], lp [S, C], step_size [n_blocks, C], log_accept [n_blocks, S, C]."""
        torch = self.torch
        td = _tdtype(torch, dtype)
        theta0 = np.asarray(theta0, dtype=dtype) if not isinstance(theta0, torch.Tensor) else theta0
        C, D, S = int(theta0.shape[0]), self.ir.D, int(num_results)
        th0 = self.to_device(theta0, dtype)
        arr = (CompoundBlock * len(blocks))()
        keep = [th0]
        codes = {"nuts": PM4_BLOCK_NUTS, "hmc": PM4_BLOCK_HMC, "rwm": PM4_BLOCK_RWM}
        for k, b in enumerate(blocks):
            scale = self.to_device(np.asarray(b["dim_scale"], dtype=np.float64).reshape(D), dtype)
            kinds = None
            if b.get("rw_kind") is not None:
                kinds = torch.from_numpy(np.ascontiguousarray(b["rw_kind"], dtype=np.int32).reshape(D)).to(self._dev())
            keep += [scale, kinds]
            arr[k] = CompoundBlock(codes[b["kind"]], int(b.get("num_leapfrog_steps", 0)), int(b.get("max_tree_depth", 10)), 0,
                                   float(b.get("max_energy_diff", 1000.0)), float(b.get("step_size", 1.0)),
                                   float(b.get("rw_scale", 1.0)), b.get("adapt") or make_adapt(enabled=False),
                                   scale.data_ptr(), kinds.data_ptr() if kinds is not None else None)
        out = dict(states=self.empty((S, C, D), td), lp=self.empty((S, C), td),
                   step_size=self.empty((len(blocks), C), td), log_accept=self.empty((len(blocks), S, C), td))
        with torch.cuda.device(self._dev()):
            rc = self.lib.pm4_compound_run(self.handle, _code(dtype), C, S, int(num_burnin),
                                           ctypes.c_uint64(int(seed) & (2**64 - 1)), int(chain_offset), int(warps_per_block),
                                           arr, len(blocks), self._p(th0), self._p(out["states"]), self._p(out["lp"]),
                                           self._p(out["step_size"]), self._p(out["log_accept"]), self._stream())
        self._check(rc)
        out["_keepalive"] = keep
        return out


class Communicator:
    """The box-wide exchange the sampler needs: one mailbox per rank in device memory, mapped into every peer
    with CUDA IPC and read over NVLink by the pooled dual-averaging kernel (include/pm4b200.h: pm4_comm_*).
    ``torch.distributed`` only carries the 64-byte IPC handles at set-up."""

    HANDLE_BYTES = 64

    def __init__(self, device: int, rank: int, world: int, group=None):
        torch = _torch()
        import torch.distributed as dist

        self.lib = load_library()
        self.rank, self.world = int(rank), int(world)
        self.handle = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(self.HANDLE_BYTES)
        rc = self.lib.pm4_comm_create(int(device), self.rank, self.world, ctypes.byref(self.handle), buf)
        if rc != 0:
            raise RuntimeError("libpm4b200: %s (code %d)" % (self.lib.pm4_last_error().decode(), rc))
        mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        if dist.get_backend(group) == "nccl":
            mine = mine.to(torch.device("cuda", int(device)))
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=group)
        handles = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
        rc = self.lib.pm4_comm_connect(self.handle, handles)
        if rc != 0:
            raise RuntimeError("libpm4b200: %s (code %d)" % (self.lib.pm4_last_error().decode(), rc))
        dist.barrier(group=group)

    def close(self):
        if self.handle and self.handle.value:
            self.lib.pm4_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()
