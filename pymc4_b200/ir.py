"""Lowering of a coroutine model's transformed log-density to the flat model IR.

What the reference does at run time -- ``logpfn`` = ``SamplingState.from_values`` ->
``evaluate_model_transformed`` -> ``collect_log_prob`` (pymc4/mcmc/samplers.py:762-770,
pymc4/flow/executor.py:171-187) -- is done here *once*, symbolically: the transformed
executor runs with every free variable bound to a symbolic view, and the distributions,
potentials and deterministics it registers are flattened into

  * a table of free variables (offset, size, transform) in the reference's trace order
    (untransformed first, then transformed; executor.py:24-35,163),
  * a list of log-density terms, each an element-wise SSA tape over ``n`` elements,
  * tapes for the traced outputs (deterministics and back-transformed variables;
    samplers.py:796-809),
  * two constant pools (reals, indices).

The layout of the packed structs is include/pm4b200.h.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import struct
import weakref
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import symbolic as sym
from .coroutine_model import Model
from .utils import NameParts

ABI_VERSION = 6

TR_IDENTITY, TR_EXP, TR_SIGMOID = 0, 1, 2

TERM_KINDS = {
    "potential": 0,
    "normal": 1,
    "halfnormal": 2,
    "halfcauchy": 3,
    "bernoulli": 4,
    "poisson": 5,
    "bernoulli_logit_glm": 6,
    "mvn_expquad": 7,
}
TERM_NPARAMS = {0: 0, 1: 2, 2: 1, 3: 1, 4: 1, 5: 1, 6: 0}

OP = {
    "ld_const": 0, "ld_data": 1, "ld_x": 2, "ld_z": 3,
    "add": 10, "sub": 11, "mul": 12, "div": 13, "pow": 14,
    "neg": 20, "exp": 21, "log": 22, "log1p": 23, "sqrt": 24, "square": 25,
    "sigmoid": 26, "softplus": 27, "tanh": 28, "abs": 29, "recip": 30, "lgamma": 31,
}
OP_NAME = {v: k for k, v in OP.items()}
MODE_SCALAR, MODE_ELEM, MODE_GATHER = 0, 1, 2


# ---------------------------------------------------------------------------
# Python-side IR
# ---------------------------------------------------------------------------
@dataclass
class RV:
    name: str            # sampled name, e.g. "m/__log_tau" or "m/mu"
    shape: Tuple[int, ...]
    offset: int
    size: int
    transform: int = TR_IDENTITY
    shift: float = 0.0
    scale: float = 1.0
    untransformed_name: Optional[str] = None  # "m/tau" for transformed variables
    discrete: bool = False  # a free discrete variable (float-typed integers): moved by the compound step's random-walk blocks only


@dataclass
class Op:
    opcode: int
    dst: int
    a: int = -1
    b: int = -1
    mode: int = 0
    base: int = 0
    blob: int = 0
    imm: float = 0.0
    sorted_idx: bool = False   # GATHER loads: index array is non-decreasing
    blob_len: int = 0
    uniform: bool = False      # value does not depend on the element index


@dataclass
class Term:
    kind: int
    n: int
    ops: List[Op]
    value_reg: int
    param_regs: List[int]
    label: str = ""
    plan: Optional["Plan"] = None
    glm: Optional["GLM"] = None     # dense Bernoulli-logit term (no tape): see GLM
    gp: Optional["GP"] = None       # marginal GP likelihood (no tape): see GP
    observed: bool = False          # density of an observed variable (enters the pointwise log-likelihood)
    ll_offset: int = 0              # position of element 0 in one state's log-likelihood row
    shape: Tuple[int, ...] = ()     # broadcast shape of the term's elements

    @property
    def n_regs(self) -> int:
        return max([o.dst for o in self.ops] + [-1]) + 1


@dataclass
class GLM:
    """``Bernoulli(probs=sigmoid(X @ beta))`` with a constant matrix X [n, P] and a free vector beta: the dense
    likelihood of the hierarchical logistic regression (BASELINE configs[2]).  It has no element-wise tape: the
    device evaluates it for all chains at once as two tensor-core GEMMs (csrc/pm4_glm.cu), the oracle with
    plain loops."""
    x_blob: int      # dataf offset of X, row-major [n, P]
    y_blob: int      # dataf offset of y [n]
    base: int        # theta offset of beta[0]
    P: int


@dataclass
class GP:
    """``MvNormal(loc, ExpQuad(l, a)(X, X) + noise I)`` observed: the marginal likelihood of GP regression
    (BASELINE configs[3]).  Like the GLM term it has no element-wise tape; the device evaluates it for all
    chains at once with batched float64 Cholesky kernels (csrc/pm4_gp.cu), the oracle with plain loops.

    ``par_blob`` points at 9 doubles in dataf: (l_elem, a_elem, noise_elem, l_const, a_const, noise_const,
    noise_squared, jitter, loc_blob) -- loc_blob is the dataf offset of loc [n] (forward sampling adds it back).  ``*_elem`` is the position in x (constrained space) of the scalar free variable
    the parameter is, or -1 when it is the constant next to it; the diagonal added to the kernel matrix is
    ``(noise_squared ? v * v : v) + jitter`` with v the noise parameter."""
    x_blob: int      # dataf offset of X, row-major [n, d]
    y_blob: int      # dataf offset of y - loc [n]
    par_blob: int
    n: int
    d: int
    elems: Tuple[int, int, int] = (-1, -1, -1)


@dataclass
class Plan:
    """Device execution plan of a term whose parameters gather from free variables.

    The term's elements are regrouped by gather target ("group"), groups are cut into chunks of
    at most ``chunk_cap`` elements, chunks are sorted by length and dealt 32*T at a time to the lanes
    of the chain's team of T warps (sliced-ELL).  Within a slice every lane walks one chunk, so the
    gathered values sit in registers and their adjoints accumulate in registers -- no scatter, no
    atomics, a fixed summation order.  Per-chunk sums are parked in the chain's ``part`` array and
    added up by the lane that owns the gradient element (pymc4_b200/plan.py).

    Structure (goes into the struct hash): which ops are planned gathers / record fields.
    Data (run-time pools): everything below.
    """
    term: int
    index_blob: int                 # datai offset of the primary index array
    gather_ops: List[int]           # op dst of GATHER loads sharing the primary index
    record_ops: List[int]           # op dst of ELEM data loads, in record order
    width: int                      # reals per record (1, 2 or 4)
    n_slices: int = 0
    n_groups: int = 0
    n_chunks: int = 0
    meta_off: int = 0               # offset into plani
    recs_off: int = 0               # offset into planf (reals)
    chunk_cap: int = 0


@dataclass
class Det:
    name: str
    shape: Tuple[int, ...]
    n: int
    ops: List[Op]
    value_reg: int
    out_offset: int

    @property
    def n_regs(self) -> int:
        return max([o.dst for o in self.ops] + [-1]) + 1


@dataclass
class ModelIR:
    D: int
    rvs: List[RV]
    terms: List[Term]
    dets: List[Det]
    dataf: np.ndarray
    datai: np.ndarray
    observed: Dict[str, np.ndarray] = field(default_factory=dict)
    planf: np.ndarray = field(default_factory=lambda: np.zeros(0))
    plani: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))
    # team layout (pymc4_b200/plan.py): warps per chain, scalar-read positions, part-array tables
    team_warps: int = 1
    scalar_pos: List[int] = field(default_factory=list)
    elem_off: int = 0      # offset of elem_tab[DPt] in plani
    spart_off: int = 0     # offset of spart[(n_scalar + 1) * T] in plani
    phdr_off: int = 0      # offset of the per-plan int4 headers in plani
    part_reals: int = 4
    part_k: int = 2        # fixed part entries per gradient element
    owned: Any = None      # owner-computes layout of the warp-per-chain path (pymc4_b200/owned.py), or None

    @property
    def plans(self) -> List["Plan"]:
        return [t.plan for t in self.terms if t.plan is not None]

    @property
    def det_row(self) -> int:
        return sum(d.n for d in self.dets)

    @property
    def glm_terms(self) -> List["Term"]:
        return [t for t in self.terms if t.glm is not None]

    @property
    def gp_terms(self) -> List["Term"]:
        return [t for t in self.terms if t.gp is not None]

    @property
    def lockstep(self) -> bool:
        """Models with a dense term run in lock-step over the chain axis (csrc/pm4_lockstep.cuh)."""
        return any(t.glm is not None or t.gp is not None for t in self.terms)

    @property
    def streamed(self) -> bool:
        """Plan pools too large to stage in shared memory (the device reads them from L2): the generated row
        loops prefetch."""
        return self.planf.size * 4 + self.plani.size * 4 > 96 * 1024

    @property
    def ll_row(self) -> int:
        return sum(t.n for t in self.terms if t.observed)

    # -- structural hash: everything the generated code depends on ---------
    def struct_key(self) -> bytes:
        h = [b"pm4ir-v%d" % ABI_VERSION, struct.pack("<ii?", self.D, self.team_warps, self.streamed)]
        for rv in self.rvs:
            h.append(struct.pack("<iiidd", rv.offset, rv.size, rv.transform, rv.shift, rv.scale))

        def ops_key(ops, n):
            for o in ops:
                h.append(struct.pack("<iiiiii?d", o.opcode, o.dst, o.a, o.b, o.mode, o.base,
                                     o.sorted_idx, o.imm))
            # n matters structurally only for terms aligned with a variable (ELEM x loads)
            if any(o.opcode in (OP["ld_x"], OP["ld_z"]) and o.mode == MODE_ELEM for o in ops):
                h.append(struct.pack("<i", n))

        for t in self.terms:
            h.append(struct.pack("<iii?", t.kind, t.value_reg, len(t.ops), t.observed))
            if t.observed:
                # the generated pointwise log-likelihood / predictive kernels bake the row layout in (LL_ROW, ll_offset)
                h.append(struct.pack("<ii", t.n, t.ll_offset))
            if t.glm is not None:
                h.append(struct.pack("<qqii", t.glm.x_blob, t.glm.y_blob, t.glm.base, t.glm.P))
            if t.gp is not None:
                h.append(struct.pack("<qqqii", t.gp.x_blob, t.gp.y_blob, t.gp.par_blob, t.gp.n, t.gp.d))
            h.append(struct.pack("<%di" % len(t.param_regs), *t.param_regs))
            ops_key(t.ops, t.n)
            if t.plan is not None:
                pl = t.plan
                h.append(b"plan" + struct.pack("<i", pl.width) + bytes(pl.gather_ops) + b"|" + bytes(pl.record_ops))
        for d in self.dets:
            h.append(struct.pack("<iii", d.value_reg, len(d.ops), d.n))
            ops_key(d.ops, d.n)
        if self.owned is not None:
            h.append(self.owned.key())
        return b"".join(h)

    def layout_key(self) -> tuple:
        """Everything a device model keeps besides the data VALUES: pool sizes, element counts, blob offsets.
        Two IRs of one structure whose layout keys agree can share a DeviceModel through ``set_data``."""
        return (self.struct_hash, self.dataf.shape, self.datai.shape, np.asarray(self.planf).shape,
                np.asarray(self.plani).shape, tuple(t.n for t in self.terms), tuple(d.n for d in self.dets),
                tuple(o.blob for t in self.terms for o in t.ops), tuple(o.blob for d in self.dets for o in d.ops))

    @property
    def struct_hash(self) -> int:
        return int.from_bytes(hashlib.sha256(self.struct_key()).digest()[:8], "little")

    @property
    def struct_hash_hex(self) -> str:
        return "%016x" % self.struct_hash

    # -- helpers -----------------------------------------------------------
    def rv_by_name(self, name: str) -> RV:
        for rv in self.rvs:
            if rv.name == name:
                return rv
        raise KeyError(name)

    def split_theta(self, theta: np.ndarray) -> Dict[str, np.ndarray]:
        """[..., D] -> {sampled name: [..., *shape]}"""
        lead = theta.shape[:-1]
        return {rv.name: theta[..., rv.offset: rv.offset + rv.size].reshape(lead + rv.shape)
                for rv in self.rvs}

    def join_theta(self, parts: Dict[str, Any]) -> np.ndarray:
        out = np.zeros(self.D, dtype=np.float64)
        for rv in self.rvs:
            out[rv.offset: rv.offset + rv.size] = np.asarray(parts[rv.name], dtype=np.float64).reshape(-1)
        return out

    def constrain(self, theta: np.ndarray) -> np.ndarray:
        """x = shift + scale * f(theta), NumPy (host bookkeeping / tests)."""
        x = np.array(theta, dtype=np.float64, copy=True)
        for rv in self.rvs:
            sl = slice(rv.offset, rv.offset + rv.size)
            if rv.transform == TR_EXP:
                x[..., sl] = rv.shift + rv.scale * np.exp(theta[..., sl])
            elif rv.transform == TR_SIGMOID:
                x[..., sl] = rv.shift + rv.scale / (1.0 + np.exp(-theta[..., sl]))
        return x


# ---------------------------------------------------------------------------
# expression -> tape lowering
# ---------------------------------------------------------------------------
class _Pools:
    def __init__(self):
        self.f: List[np.ndarray] = []
        self.i: List[np.ndarray] = []
        self.f_len = 0
        self.i_len = 0
        self._fkey: Dict[bytes, int] = {}
        self._fbig: Dict[Any, List[Tuple[int, np.ndarray]]] = {}
        self._ikey: Dict[bytes, int] = {}

    def add_f(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype=np.float64).reshape(-1)
        big = arr.size > (1 << 20)
        if big:
            # large blobs (a 10^6 x 100 design matrix): a sampled fingerprint finds the candidates, an exact
            # comparison confirms a duplicate -- hashing 800 MB per lowering would cost more than the upload
            key = (arr.size, arr[:64].tobytes(), arr[-64:].tobytes(), arr[:: arr.size // 4096].tobytes())
            for off, ref in self._fbig.get(key, ()):
                if np.array_equal(ref, arr):
                    return off
        else:
            key = arr.tobytes()
            if key in self._fkey:
                return self._fkey[key]
        # keep every blob 4-element aligned so 128-bit loads stay legal
        pad = (-self.f_len) % 4
        if pad:
            self.f.append(np.zeros(pad))
            self.f_len += pad
        off = self.f_len
        self.f.append(arr)
        self.f_len += arr.size
        if big:
            self._fbig.setdefault(key, []).append((off, arr))
        else:
            self._fkey[key] = off
        return off

    def add_i(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype=np.int32).reshape(-1)
        key = arr.tobytes()
        if key in self._ikey:
            return self._ikey[key]
        pad = (-self.i_len) % 4
        if pad:
            self.i.append(np.zeros(pad, dtype=np.int32))
            self.i_len += pad
        off = self.i_len
        self.i.append(arr)
        self.i_len += arr.size
        self._ikey[key] = off
        return off


class _TapeBuilder:
    """Lower expressions to one SSA tape over ``n`` elements."""

    def __init__(self, n: int, rvs: Dict[str, RV], pools: _Pools):
        self.n = n
        self.rvs = rvs
        self.pools = pools
        self.ops: List[Op] = []
        self._memo: Dict[Tuple[int, bytes], int] = {}
        self._keepalive: List[Any] = []

    def _emit(self, **kw) -> int:
        dst = len(self.ops)
        self.ops.append(Op(dst=dst, **kw))
        return dst

    def lower(self, node: sym.Expr, emap: np.ndarray) -> int:
        """Register holding ``node.reshape(-1)[emap[i]]`` for element i."""
        key = (id(node), hashlib.blake2b(emap.tobytes(), digest_size=8).digest())
        if key in self._memo:
            return self._memo[key]
        self._keepalive.append(node)
        reg = self._lower(node, emap)
        self._memo[key] = reg
        return reg

    def _lower(self, node: sym.Expr, emap: np.ndarray) -> int:
        n = self.n
        op = node.op
        if op == "const":
            vals = node.data.reshape(-1)[emap]
            if vals.size == 0 or np.all(vals == vals[0]) or (np.all(np.isnan(vals))):
                return self._emit(opcode=OP["ld_const"], imm=float(vals[0]) if vals.size else 0.0,
                                  uniform=True)
            blob = self.pools.add_f(vals)
            return self._emit(opcode=OP["ld_data"], mode=MODE_ELEM, blob=blob, blob_len=n)
        if op == "view":
            rv = self.rvs[node.rv]
            idx = node.index.reshape(-1)[emap].astype(np.int64)
            opcode = OP["ld_x"] if node.space == "x" else OP["ld_z"]
            if rv.transform == TR_IDENTITY:
                opcode = OP["ld_x"]  # identical arrays; keep one spelling
            if idx.size == 0 or np.all(idx == idx[0]):
                return self._emit(opcode=opcode, mode=MODE_SCALAR, base=rv.offset + int(idx[0]),
                                  uniform=True)
            if np.array_equal(idx, idx[0] + np.arange(n)):
                return self._emit(opcode=opcode, mode=MODE_ELEM, base=rv.offset + int(idx[0]))
            blob = self.pools.add_i(idx)
            return self._emit(opcode=opcode, mode=MODE_GATHER, base=rv.offset, blob=blob,
                              blob_len=n, sorted_idx=bool(np.all(np.diff(idx) >= 0)))
        if op in sym.UNARY_OPS:
            a = self.lower(node.args[0], emap)
            return self._emit(opcode=OP[op], a=a, uniform=self.ops[a].uniform)
        if op in sym.BINARY_OPS:
            regs = []
            for arg in node.args:
                regs.append(self.lower(arg, _broadcast_map(emap, node.shape, arg.shape)))
            return self._emit(opcode=OP[op], a=regs[0], b=regs[1],
                              uniform=self.ops[regs[0]].uniform and self.ops[regs[1]].uniform)
        if op == "broadcast":
            arg = node.args[0]
            return self.lower(arg, _broadcast_map(emap, node.shape, arg.shape))
        if op == "reshape":
            return self.lower(node.args[0], emap)
        if op == "take":
            return self.lower(node.args[0], node.index.reshape(-1).astype(np.int64)[emap])
        hint = ""
        if op == "matvec":
            hint = (".  A dense linear predictor runs on the device only as the whole likelihood "
                    "pm.Bernoulli(name, probs=pm.math.sigmoid(pm.math.matvec(X, beta)), observed=y) with a constant X and a "
                    "free vector beta (for an intercept, append a column of ones to X)")
        elif op == "expquad":
            hint = (".  A covariance matrix with free parameters runs on the device only inside an observed "
                    "pm.MvNormal(name, loc, covariance_matrix=ExpQuad(l, a)(X, X) + noise * np.eye(n), observed=y)")
        raise NotImplementedError(
            "op {!r} on random variables cannot be lowered to the element-wise model IR yet "
            "(supported: arithmetic, exp/log/sigmoid/..., static indexing, gather, broadcasting){}"
            .format(op, hint)
        )


def _broadcast_map(emap: np.ndarray, out_shape: Tuple[int, ...], arg_shape: Tuple[int, ...]) -> np.ndarray:
    """Map flat positions of a broadcast result to flat positions of one argument."""
    if tuple(arg_shape) == tuple(out_shape):
        return emap
    size = int(np.prod(arg_shape, dtype=np.int64)) if arg_shape else 1
    if size == 1:
        return np.zeros_like(emap)
    pos = np.arange(size, dtype=np.int64).reshape(arg_shape)
    return np.broadcast_to(pos, out_shape).reshape(-1)[emap]


def _identity_map(n: int) -> np.ndarray:
    return np.arange(n, dtype=np.int64)


def _match_glm(kind: str, value, params: List[Any], rvs, pools, label: str) -> Optional[Term]:
    """Bernoulli(probs=sigmoid(X @ beta)) observed, with constant X and beta a whole free vector -> GLM term."""
    if kind != "bernoulli" or len(params) != 1:
        return None
    p, v = sym.as_expr(params[0]), sym.as_expr(value)
    if p.op != "sigmoid" or p.args[0].op != "matvec" or v.op != "const":
        return None
    X, b = p.args[0].args
    if X.op != "const" or b.op != "view" or b.space != "x":
        return None
    rv = rvs[b.rv]
    idx = np.asarray(b.index).reshape(-1)
    if rv.transform != TR_IDENTITY or idx.size != rv.size or not np.array_equal(idx, np.arange(rv.size)):
        return None
    n, P = X.shape
    if v.shape != (n,):
        return None
    x_blob = pools.add_f(np.asarray(X.data, dtype=np.float64).reshape(-1))
    y_blob = pools.add_f(np.asarray(v.data, dtype=np.float64).reshape(-1))
    return Term(kind=TERM_KINDS["bernoulli_logit_glm"], n=int(n), ops=[], value_reg=-1, param_regs=[], label=label,
                shape=(int(n),), glm=GLM(x_blob=x_blob, y_blob=y_blob, base=rv.offset, P=int(P)))


def _scalar_param(e: "sym.Expr", rvs) -> Optional[Tuple[int, float]]:
    """(position in x, 0) for a scalar view of a free variable, (-1, value) for a constant, else None."""
    if e.op == "reshape" and e.size == 1:
        e = e.args[0]
    if e.size != 1:
        return None
    if e.op == "const":
        return -1, float(np.asarray(e.data).reshape(-1)[0])
    if e.op == "view" and e.space == "x":
        return int(rvs[e.rv].offset + int(np.asarray(e.index).reshape(-1)[0])), 0.0
    return None


def _match_gp(kind: str, value, params: List[Any], rvs, pools, label: str) -> Optional[Term]:
    """MvNormal(loc, expquad(l, a; X, X) + noise * I [+ jitter * I]) with a constant value and loc -> GP term."""
    if kind == "mvn_chol":
        raise NotImplementedError("MvNormalCholesky runs on the device with a constant scale_tril (a latent vector, or an "
                                  "observation with a free loc); a scale_tril that depends on free variables does not")
    if kind != "mvn":
        return None
    why = ("MvNormal runs on the device as the marginal likelihood of GP regression: an observed vector, a constant "
           "loc, covariance_matrix = ExpQuad(l, a)(X, X) + noise * np.eye(n) with l, a, noise scalars (free variables "
           "or constants; noise may be sigma or sigma ** 2)")
    loc, cov, v = sym.as_expr(params[0]), sym.as_expr(params[1]), sym.as_expr(value)
    if v.op != "const" or loc.op != "const":
        raise NotImplementedError(why)
    # flatten the sum
    addends, stack = [], [cov]
    while stack:
        e = stack.pop()
        if e.op == "add":
            stack.extend(e.args)
        else:
            addends.append(e)
    kern = [e for e in addends if e.op == "expquad"]
    rest = [e for e in addends if e.op != "expquad"]
    if len(kern) != 1 or not np.array_equal(kern[0].data[0], kern[0].data[1]):
        raise NotImplementedError(why)
    n = kern[0].shape[0]
    eye = None

    def is_eye_times(e):  # const c * I -> c
        nonlocal eye
        if e.op != "const" or e.shape != (n, n):
            return None
        dg = np.diag(e.data)
        if not np.all(dg == dg[0]) or np.count_nonzero(e.data) != (n if dg[0] != 0 else 0):
            return None
        return float(dg[0])

    jitter, noise = 0.0, None
    for e in rest:
        c = is_eye_times(e)
        if c is not None:
            jitter += c
            continue
        if e.op == "mul" and noise is None:
            a, b = e.args
            if is_eye_times(b) is None:
                a, b = b, a
            c = is_eye_times(b)
            squared = False
            if c == 1.0:
                if a.op == "square" or (a.op == "pow" and a.args[1].op == "const" and float(a.args[1].data) == 2.0):
                    a, squared = a.args[0], True
                sp = _scalar_param(a, rvs)
                if sp is not None:
                    noise = (sp, squared)
                    continue
        raise NotImplementedError(why)
    ls, amp = _scalar_param(kern[0].args[0], rvs), _scalar_param(kern[0].args[1], rvs)
    if ls is None or amp is None:
        raise NotImplementedError(why)
    if noise is None:
        noise = ((-1, 0.0), False)
    (ne, nc), squared = noise
    X = np.asarray(kern[0].data[0], dtype=np.float64)
    loc_vec = np.broadcast_to(np.asarray(loc.data, dtype=np.float64), (n,)).copy()
    y = np.asarray(v.data, dtype=np.float64).reshape(-1) - loc_vec
    if y.shape != (n,):
        raise ValueError("observed value of shape {} does not match the {} x {} covariance".format(v.shape, n, n))
    x_blob, y_blob, loc_blob = pools.add_f(X.reshape(-1)), pools.add_f(y), pools.add_f(loc_vec)
    par = np.array([ls[0], amp[0], ne, ls[1], amp[1], nc, 1.0 if squared else 0.0, jitter, float(loc_blob)], dtype=np.float64)
    par_blob = pools.add_f(par)
    return Term(kind=TERM_KINDS["mvn_expquad"], n=1, ops=[], value_reg=-1, param_regs=[], label=label, shape=(),
                gp=GP(x_blob=x_blob, y_blob=y_blob, par_blob=par_blob, n=int(n), d=int(X.shape[1]),
                      elems=(ls[0], amp[0], ne)))


def _mvn_constant_term(kind: str, value, params: List[Any], rvs, pools, label: str) -> Optional[Term]:
    """MvNormal / MvNormalCholesky with a CONSTANT covariance (Cholesky factor): a latent vector, or an observation
    with a free ``loc``.  With P = Sigma^-1 the density is ``-1/2 sum_ij P_ij r_i r_j - sum log L_ii - k/2 log 2 pi``,
    r = value - loc: an element-wise potential over the k (k + 1) / 2 pairs i >= j (off-diagonal pairs count twice),
    two gathers of r per element.  Reference: distributions/multivariate.py:159-199, 331-375."""
    if kind not in ("mvn", "mvn_chol"):
        return None
    mat = sym.as_expr(params[1])
    if mat.op != "const":
        return None
    M = np.asarray(mat.data, dtype=np.float64)
    L = np.linalg.cholesky(M) if kind == "mvn" else np.tril(M)
    k = L.shape[0]
    if np.any(np.diag(L) <= 0):
        raise ValueError("{}: the Cholesky factor needs a positive diagonal".format(label))
    Linv = np.linalg.solve(L, np.eye(k))
    P = Linv.T @ Linv
    v, loc = sym.as_expr(value), sym.as_expr(params[0])
    if v.ndim != 1 or v.shape[0] != k:
        raise NotImplementedError("{}: batches of multivariate normal vectors are not lowered (value shape {})".format(label, v.shape))
    r = v - (loc if loc.ndim == 1 else sym.broadcast_to(loc, (k,)))
    I, J = np.tril_indices(k)
    w = np.where(I == J, -0.5, -1.0) * P[I, J]
    const = (-np.sum(np.log(np.diag(L))) - 0.5 * k * np.log(2.0 * np.pi)) / len(I)
    quad = sym.const(w) * sym.gather(r, I) * sym.gather(r, J) + const
    return _lower_term("potential", quad, [], rvs, pools, label=label)


def _lower_term(kind: str, value, params: List[Any], rvs, pools, label="") -> Term:
    mv = _mvn_constant_term(kind, value, params, rvs, pools, label)
    if mv is not None:
        return mv
    glm = _match_glm(kind, value, params, rvs, pools, label)
    if glm is not None:
        return glm
    gp = _match_gp(kind, value, params, rvs, pools, label)
    if gp is not None:
        return gp
    exprs = [sym.as_expr(value)] + [sym.as_expr(p) for p in params]
    shape = tuple(np.broadcast_shapes(*[e.shape for e in exprs]))
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    tb = _TapeBuilder(n, rvs, pools)
    regs = [tb.lower(e, _broadcast_map(_identity_map(n), shape, e.shape)) for e in exprs]
    return Term(kind=TERM_KINDS[kind], n=n, ops=tb.ops, value_reg=regs[0], param_regs=regs[1:], label=label,
                shape=shape)


def _lower_det(name: str, value, rvs, pools, out_offset: int) -> Det:
    e = sym.as_expr(value)
    n = e.size
    tb = _TapeBuilder(n, rvs, pools)
    reg = tb.lower(e, _identity_map(n))
    return Det(name=name, shape=e.shape, n=n, ops=tb.ops, value_reg=reg, out_offset=out_offset)


# ---------------------------------------------------------------------------
# model -> IR
# ---------------------------------------------------------------------------
def _transform_record(dist) -> Tuple[int, float, float]:
    tr = dist.transform
    if tr is None:
        return TR_IDENTITY, 0.0, 1.0
    return int(tr.ir_code), float(tr.ir_shift), float(tr.ir_scale)


# ---------------------------------------------------------------------------
# lowering cache: a sampler object that is called again on the SAME model object (pm.sample in a loop, the bench's
# end-to-end arm) does not lower, concatenate and compare hundreds of megabytes of constant data again
# ---------------------------------------------------------------------------
_lower_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _data_fingerprint(model) -> tuple:
    """Identity, shape, dtype and a strided sample of every array the model object can reach (arguments bound by the
    template call, closure cells of the generator function): an in-place edit of the data between two calls changes it."""
    arrays, seen = [], set()

    def visit(v, depth=0):
        if isinstance(v, np.ndarray):
            if id(v) not in seen:
                seen.add(id(v))
                arrays.append(v)
        elif isinstance(v, (list, tuple)) and depth < 2 and len(v) <= 64:
            for e in v:
                visit(e, depth + 1)
        elif isinstance(v, dict) and depth < 2 and len(v) <= 64:
            for e in v.values():
                visit(e, depth + 1)

    fn = getattr(model, "genfn", None)
    while fn is not None:
        for a in getattr(fn, "args", ()) or ():
            visit(a)
        visit(getattr(fn, "keywords", None) or {})
        for cell in getattr(fn, "__closure__", None) or ():
            try:
                visit(cell.cell_contents)
            except ValueError:
                pass
        fn = getattr(fn, "func", None)
    out = []
    for a in arrays:
        flat = a.reshape(-1)
        step = max(1, flat.size // 1024)
        sample = np.ascontiguousarray(flat[::step][:2048])
        out.append((id(a), a.shape, str(a.dtype), hashlib.sha1(sample.tobytes()).hexdigest()))
    return tuple(out)


def lower_model(model: Model, observed: Optional[dict] = None, state=None,
                team_warps: Optional[int] = None) -> Tuple[ModelIR, Any, List[str]]:
    """Cached front of ``_lower_model`` (only for calls without ``observed`` / ``state`` overrides)."""
    cacheable = observed is None and state is None and os.environ.get("PM4_NO_LOWER_CACHE") is None
    if cacheable:
        try:
            fp = (_data_fingerprint(model), team_warps, os.environ.get("PM4_TEAM_WARPS"))
            hit = _lower_cache.get(model)
            if hit is not None and hit[0] == fp:
                return hit[1]
        except TypeError:
            cacheable = False
    result = _lower_model(model, observed, state, team_warps)
    if cacheable:
        try:
            _lower_cache[model] = (fp, result)
        except TypeError:
            pass
    return result


def _lower_model(model: Model, observed: Optional[dict] = None, state=None,
                 team_warps: Optional[int] = None) -> Tuple[ModelIR, Any, List[str]]:
    """Trace ``model`` and return (IR, initial sampling state, deterministic names).

    The initial state and the deterministic-name list are exactly what the reference's
    ``initialize_sampling_state`` returns (pymc4/mcmc/utils.py:14-36).
    """
    from . import flow
    from .mcmc.utils import initialize_sampling_state

    if not isinstance(model, Model):
        raise TypeError(
            "`sample` function only supports `pymc4.Model` objects, but you've passed `{}`".format(type(model))
        )
    init_state, deterministic_names = initialize_sampling_state(model, observed=observed, state=state)
    if not init_state.all_unobserved_values:
        raise ValueError(
            "Can not calculate a log probability: the model {} has no unobserved values.".format(model.name or "")
        )

    # pass 1 (meta, concrete) gave names + shapes; build symbolic stand-ins
    rvs: Dict[str, RV] = {}
    sym_values: Dict[str, sym.Expr] = {}
    offset = 0
    for name, value in init_state.all_unobserved_values.items():
        shape = tuple(np.shape(sym.as_numpy(value)))
        size = int(np.prod(shape, dtype=np.int64)) if shape else 1
        parts = NameParts.from_name(name)
        rvs[name] = RV(name=name, shape=shape, offset=offset, size=size,
                       untransformed_name=parts.full_untransformed_name if parts.is_transformed else None)
        sym_values[name] = sym.free_rv(name, shape, space="z" if parts.is_transformed else "x")
        offset += size
    D = offset

    # pass 2 (transformed executor, symbolic)
    st = flow.SamplingState.from_values(sym_values, observed_values=init_state.observed_values)
    _, st = flow.evaluate_model_transformed(model, state=st)

    # fill in transform records now that the distributions are known
    all_dists = dict(st.discrete_distributions)
    all_dists.update(st.continuous_distributions)
    for name, dist in all_dists.items():
        if name in st.observed_values:
            continue
        if dist.transform is not None:
            tname = NameParts.from_name(name).replace_transform(dist.transform.name).full_original_name
            if tname in rvs:
                rvs[tname].transform, rvs[tname].shift, rvs[tname].scale = _transform_record(dist)
                continue
        if name in rvs and (not dist._grad_support or name in st.discrete_distributions):
            # a free discrete variable: a coordinate of the state the gradient-based kernels keep frozen; the sampler
            # classes refuse it unless a random-walk block owns it (reference samplers.py:62-66, compound step :490-726)
            rvs[name].discrete = True

    pools = _Pools()
    terms: List[Term] = []
    for name, dist in all_dists.items():
        if dist._kind is None:
            raise NotImplementedError(
                "{} has no device log-density (supported kinds: {})".format(type(dist).__name__, sorted(TERM_KINDS))
            )
        value = st.all_values[name]
        if dist._kind == "categorical":  # log-mass as an element-wise expression of the value (distributions/discrete.py)
            terms.append(_lower_term("potential", dist.log_prob_expr(value), [], rvs, pools, label=name))
            continue
        params = [dist.params[p] for p in dist._param_names]
        terms.append(_lower_term(dist._kind, value, params, rvs, pools, label=name))
        terms[-1].observed = name in st.observed_values
    for k, pot in enumerate(st.potentials):
        terms.append(_lower_term("potential", pot.value, [], rvs, pools, label="potential_%d" % k))

    # traced outputs, in the order the reference's deterministics_callback returns them
    dets: List[Det] = []
    out_off = 0
    det_values = dict(st.deterministics_values)
    for tname in st.transformed_values:
        uname = NameParts.from_name(tname).full_untransformed_name
        if uname in st.untransformed_values:
            det_values[uname] = st.untransformed_values[uname]
    for name in deterministic_names:
        if name not in det_values:
            continue
        d = _lower_det(name, det_values[name], rvs, pools, out_off)
        dets.append(d)
        out_off += d.n

    ll_off = 0
    for t in terms:
        if t.observed:
            t.ll_offset = ll_off
            ll_off += t.n
    dataf = np.concatenate(pools.f) if pools.f else np.zeros(0)
    datai = np.concatenate(pools.i).astype(np.int32) if pools.i else np.zeros(0, dtype=np.int32)
    observed_np = {k: sym.as_numpy(v) for k, v in st.observed_values.items()}
    ir = ModelIR(D=D, rvs=list(rvs.values()), terms=terms, dets=dets, dataf=dataf, datai=datai,
                 observed=observed_np)
    from .owned import attach_owned
    from .plan import attach_plans

    attach_plans(ir, team_warps)
    attach_owned(ir)
    return ir, init_state, deterministic_names


# ---------------------------------------------------------------------------
# ctypes packing (layout: include/pm4b200.h)
# ---------------------------------------------------------------------------
class CRV(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_int32), ("size", ctypes.c_int32), ("transform", ctypes.c_int32),
                ("_pad", ctypes.c_int32), ("shift", ctypes.c_double), ("scale", ctypes.c_double)]


class CTerm(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("n", ctypes.c_int32), ("op_begin", ctypes.c_int32),
                ("op_end", ctypes.c_int32), ("n_regs", ctypes.c_int32), ("value_reg", ctypes.c_int32),
                ("param_reg", ctypes.c_int32 * 3), ("plan", ctypes.c_int32), ("observed", ctypes.c_int32),
                ("ll_offset", ctypes.c_int32), ("glm_base", ctypes.c_int32), ("glm_p", ctypes.c_int32),
                ("glm_x", ctypes.c_int64), ("glm_y", ctypes.c_int64), ("gp_par", ctypes.c_int64)]


class CPlan(ctypes.Structure):
    _fields_ = [("term", ctypes.c_int32), ("n_gather", ctypes.c_int32), ("width", ctypes.c_int32),
                ("_pad", ctypes.c_int32), ("meta", ctypes.c_int64), ("recs", ctypes.c_int64)]


class COp(ctypes.Structure):
    _fields_ = [("opcode", ctypes.c_int32), ("dst", ctypes.c_int32), ("a", ctypes.c_int32),
                ("b", ctypes.c_int32), ("mode", ctypes.c_int32), ("base", ctypes.c_int32),
                ("blob", ctypes.c_int64), ("imm", ctypes.c_double)]


class CDet(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("op_begin", ctypes.c_int32), ("op_end", ctypes.c_int32),
                ("n_regs", ctypes.c_int32), ("value_reg", ctypes.c_int32), ("out_offset", ctypes.c_int32)]


class CIR(ctypes.Structure):
    _fields_ = [("abi_version", ctypes.c_int32), ("D", ctypes.c_int32), ("n_rv", ctypes.c_int32),
                ("n_terms", ctypes.c_int32), ("n_ops", ctypes.c_int32), ("n_dets", ctypes.c_int32),
                ("det_row", ctypes.c_int32), ("_pad", ctypes.c_int32),
                ("rvs", ctypes.POINTER(CRV)), ("terms", ctypes.POINTER(CTerm)),
                ("ops", ctypes.POINTER(COp)), ("dets", ctypes.POINTER(CDet)),
                ("n_dataf", ctypes.c_int64), ("dataf", ctypes.POINTER(ctypes.c_double)),
                ("n_datai", ctypes.c_int64), ("datai", ctypes.POINTER(ctypes.c_int32)),
                ("struct_hash", ctypes.c_uint64),
                ("n_plans", ctypes.c_int32), ("_pad2", ctypes.c_int32), ("plans", ctypes.POINTER(CPlan)),
                ("n_planf", ctypes.c_int64), ("planf", ctypes.POINTER(ctypes.c_double)),
                ("n_plani", ctypes.c_int64), ("plani", ctypes.POINTER(ctypes.c_int32)),
                ("team_warps", ctypes.c_int32), ("part_reals", ctypes.c_int32),
                ("elem_off", ctypes.c_int64), ("spart_off", ctypes.c_int64), ("phdr_off", ctypes.c_int64)]


class PackedIR:
    """Owns the ctypes arrays a ``pm4_ir_t`` points into."""

    def __init__(self, ir: ModelIR):
        self.ir = ir
        all_ops: List[Op] = []
        plans: List[Plan] = []
        cterms = (CTerm * max(len(ir.terms), 1))()
        for k, t in enumerate(ir.terms):
            begin = len(all_ops)
            all_ops.extend(t.ops)
            pr = list(t.param_regs) + [-1] * (3 - len(t.param_regs))
            plan_idx = -1
            if t.plan is not None:
                plan_idx = len(plans)
                plans.append(t.plan)
            g, gp = t.glm, t.gp
            if gp is not None:  # the dense-term fields are shared: (n points, d features, X, y - loc, parameter record)
                dense = (gp.n, gp.d, gp.x_blob, gp.y_blob, gp.par_blob)
            else:
                dense = (g.base if g else 0, g.P if g else 0, g.x_blob if g else 0, g.y_blob if g else 0, 0)
            cterms[k] = CTerm(t.kind, t.n, begin, len(all_ops), t.n_regs, t.value_reg,
                              (ctypes.c_int32 * 3)(*pr), plan_idx, 1 if t.observed else 0, t.ll_offset, *dense)
        cdets = (CDet * max(len(ir.dets), 1))()
        for k, d in enumerate(ir.dets):
            begin = len(all_ops)
            all_ops.extend(d.ops)
            cdets[k] = CDet(d.n, begin, len(all_ops), d.n_regs, d.value_reg, d.out_offset)
        cops = (COp * max(len(all_ops), 1))()
        for k, o in enumerate(all_ops):
            cops[k] = COp(o.opcode, o.dst, o.a, o.b, o.mode, o.base, o.blob, o.imm)
        crvs = (CRV * max(len(ir.rvs), 1))()
        for k, rv in enumerate(ir.rvs):
            crvs[k] = CRV(rv.offset, rv.size, rv.transform, 0, rv.shift, rv.scale)
        self._dataf = np.ascontiguousarray(ir.dataf, dtype=np.float64)
        self._datai = np.ascontiguousarray(ir.datai, dtype=np.int32)
        cplans = (CPlan * max(len(plans), 1))()
        for k, pl in enumerate(plans):
            cplans[k] = CPlan(pl.term, len(pl.gather_ops), pl.width, 0, pl.meta_off, pl.recs_off)
        self._planf = np.ascontiguousarray(ir.planf, dtype=np.float64)
        self._plani = np.ascontiguousarray(ir.plani, dtype=np.int32)
        self._keep = (cterms, cdets, cops, crvs, cplans)
        self.c = CIR(
            ABI_VERSION, ir.D, len(ir.rvs), len(ir.terms), len(all_ops), len(ir.dets), ir.det_row, 0,
            ctypes.cast(crvs, ctypes.POINTER(CRV)), ctypes.cast(cterms, ctypes.POINTER(CTerm)),
            ctypes.cast(cops, ctypes.POINTER(COp)), ctypes.cast(cdets, ctypes.POINTER(CDet)),
            self._dataf.size, self._dataf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            self._datai.size, self._datai.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            ir.struct_hash,
            len(plans), 0, ctypes.cast(cplans, ctypes.POINTER(CPlan)),
            self._planf.size, self._planf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            self._plani.size, self._plani.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            ir.team_warps, ir.part_reals, ir.elem_off, ir.spart_off, ir.phdr_off,
        )

    def byref(self):
        return ctypes.byref(self.c)
