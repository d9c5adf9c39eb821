"""Model IR -> specialised CUDA module (nvcc, sm_100a).

The hand-written device code lives in ``csrc/`` (``pm4_dists.cuh``: log-densities with
analytic partials; ``pm4_kernels.cuh``: logp/grad, HMC and NUTS kernels; ``pm4_philox.cuh``:
RNG).  This module only emits the *glue*: a ``Model`` struct whose ``terms`` member is the
model's tape unrolled into straight-line code with the reverse sweep written out, so that
nvcc inlines the density functions into the persistent leapfrog / NUTS kernel and keeps
the chain state in registers.  One shared object is built per IR *structure*
(``ModelIR.struct_hash``); data values, element counts, pool offsets and execution plans stay
run-time arguments, so new observed data reuses the same module.

A chain is evaluated by a team of T warps (TL = 32 T lanes, ``ModelIR.team_warps``); lane ``tl`` of
the team holds elements tl, tl + TL, ... of every D-vector.  Every term is emitted in one of four
loop shapes:
  uniform   n == 1 and nothing depends on the element index: evaluated once, on team lane 0;
  aligned   the term is element-wise over ONE free vector variable: lane tl walks the elements it
            already holds in registers (value from ``x[r]``, adjoint into ``gr[r]``);
  planned   the term gathers from free variables through one index array: sliced-ELL walk over
            chunks of equal gather target (pymc4_b200/plan.py), gathered values and their
            adjoints in registers, records from the staged shared-memory pool, chunk sums parked
            in the chain's ``part`` array for the lane that owns the gradient element;
  general   team-strided loop over the elements (adjoints of gathers fall back to atomics on a
            shared-memory gradient array, ``NEEDS_GX``).
Adjoints of parameters read as chain-uniform scalars and the log-density itself are reduced with
one packed warp butterfly per warp and parked in ``part`` too (one entry per warp).
Normal terms with a chain-uniform scale use the factored form: only r = value - loc and the sums
of r, r*x ..., r*r are formed per element; 1/scale^2 multiplies the sums once.

Reference semantics being lowered: ``collect_log_prob`` (pymc4/flow/executor.py:171-187)
through ``TransformedSamplingExecutor`` (pymc4/flow/transformed_executor.py:22-138).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import tempfile
import threading
from typing import Dict, List, Optional, Tuple

from . import ir as I

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
CACHE_DIR = os.environ.get("PM4_KERNEL_CACHE", os.path.join(_HERE, "_kcache"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# experiments: extra nvcc flags for the model modules (part of the module hash)
NVCC_FLAGS += os.environ.get("PM4_EXTRA_NVCC_FLAGS", "").split()
# threads per block of the persistent single-precision kernels (registers per thread = 65536 / MAXT):
# 512 = 128 registers, no spills; measured best on the radon model (1024: 64 registers, 280 B of spills)
MAXT = int(os.environ.get("PM4_MAXT", "512"))
_build_lock = threading.Lock()
_path_locks: Dict[str, threading.Lock] = {}

_UNARY_FWD = {
    "neg": "-{a}", "exp": "pm4::m_exp({a})", "log": "pm4::m_log({a})", "log1p": "pm4::m_log1p({a})",
    "sqrt": "pm4::m_sqrt({a})", "square": "{a} * {a}", "sigmoid": "pm4::m_sigmoid<real>({a})",
    "softplus": "pm4::m_softplus<real>({a})", "tanh": "pm4::m_tanh({a})", "abs": "pm4::m_abs({a})",
    "recip": "pm4::m_rcp({a})", "lgamma": "pm4::m_lgamma({a})",
}
# d(out)/d(a) as an expression of a (input), r (output) and the incoming adjoint g
_UNARY_BWD = {
    "neg": "-{g}", "exp": "{g} * {r}", "log": "{g} * pm4::m_rcp({a})", "log1p": "{g} * pm4::m_rcp((real)1 + {a})",
    "sqrt": "{g} * pm4::m_rcp((real)2 * {r})", "square": "(real)2 * {g} * {a}", "sigmoid": "{g} * {r} * ((real)1 - {r})",
    "softplus": "{g} * pm4::m_sigmoid<real>({a})", "tanh": "{g} * ((real)1 - {r} * {r})",
    "abs": "(({a} > 0) ? {g} : (({a} < 0) ? -{g} : (real)0))", "recip": "-{g} * {r} * {r}",
    "lgamma": "{g} * pm4::m_digamma<real>({a})",
}
_KIND = {v: k for k, v in I.TERM_KINDS.items()}


def _lit(v: float) -> str:
    if v != v:
        return "(real)NAN"
    if v in (float("inf"), float("-inf")):
        return "(-pm4::m_inf<real>())" if v < 0 else "pm4::m_inf<real>()"
    return "(real)%r" % float(v)


class _Tape:
    """Forward / reverse expressions of one SSA tape."""

    def __init__(self, ops: List[I.Op], op_index0: int, tag: str):
        self.ops = ops
        self.op0 = op_index0
        self.tag = tag

    def v(self, reg: int) -> str:
        return "%s_v%d" % (self.tag, reg)

    def fwd(self, o: I.Op, idx: str, xread, dread) -> str:
        name = I.OP_NAME[o.opcode]
        if name == "ld_const":
            return _lit(o.imm)
        if name == "ld_data":
            return dread(o, idx)
        if name == "ld_z":
            raise NotImplementedError("unconstrained-space loads (ld_z) are not generated yet")
        if name == "ld_x":
            return xread(o, idx)
        a = self.v(o.a)
        if name in _UNARY_FWD:
            return _UNARY_FWD[name].format(a=a)
        b = self.v(o.b)
        return {"add": "%s + %s", "sub": "%s - %s", "mul": "%s * %s", "div": "%s * pm4::m_rcp(%s)",
                "pow": "pm4::m_pow(%s, %s)"}[name] % (a, b)

    def bwd(self, o: I.Op, g: str) -> List[Tuple[int, str]]:
        name = I.OP_NAME[o.opcode]
        r = self.v(o.dst)
        if name in _UNARY_BWD:
            return [(o.a, _UNARY_BWD[name].format(g=g, a=self.v(o.a), r=r))]
        if o.b < 0:
            return []
        a, b = self.v(o.a), self.v(o.b)
        if name == "add":
            return [(o.a, g), (o.b, g)]
        if name == "sub":
            return [(o.a, g), (o.b, "-" + g)]
        if name == "mul":
            return [(o.a, "%s * %s" % (g, b)), (o.b, "%s * %s" % (g, a))]
        if name == "div":
            return [(o.a, "%s * pm4::m_rcp(%s)" % (g, b)), (o.b, "-%s * %s * pm4::m_rcp(%s)" % (g, r, b))]
        if name == "pow":
            return [(o.a, "%s * %s * pm4::m_pow(%s, %s - (real)1)" % (g, b, a, b)),
                    (o.b, "((%s > 0) ? %s * %s * pm4::m_log(%s) : (real)0)" % (a, g, r, a))]
        return []


class _ModuleGen:
    def __init__(self, ir: I.ModelIR, has_f64: bool):
        self.ir = ir
        self.has_f64 = has_f64
        self.D = ir.D
        self.T = int(ir.team_warps)
        self.TL = 32 * self.T
        self.R = max(1, -(-ir.D // self.TL))
        self.lines: List[str] = []
        self.scalar_pos: List[int] = list(ir.scalar_pos)  # scalar x positions, part-array order
        self.needs_gx = False
        # element-independent terms run on ONE lane: lane 0 of the team's last warp (chunks are dealt
        # longest-first, so the first warp is the one with the most rows to walk)
        self.uni_lane = self.TL - 32

    def emit(self, s: str = "", ind: int = 2):
        self.lines.append("  " * ind + s)

    # -- constrain -----------------------------------------------------------
    def gen_constrain(self) -> str:
        out = ["  template <typename real>",
               "  static __device__ __forceinline__ void constrain(int r, int tl, real z, real& x, real& dxdz) {",
               "    const int j = tl + %d * r; (void)j;" % self.TL,
               "    x = z; dxdz = (real)1;"]
        # consecutive variables with the same transform share one range test (one branch, not one each)
        merged: List[I.RV] = []
        for rv in self.ir.rvs:
            prev = merged[-1] if merged else None
            if (prev is not None and prev.offset + prev.size == rv.offset
                    and (prev.transform, prev.shift, prev.scale) == (rv.transform, rv.shift, rv.scale)):
                merged[-1] = I.RV(name=prev.name, shape=(prev.size + rv.size,), offset=prev.offset,
                                  size=prev.size + rv.size, transform=prev.transform, shift=prev.shift, scale=prev.scale)
            else:
                merged.append(rv)
        for rv in merged:
            cond = "j >= %d && j < %d" % (rv.offset, rv.offset + rv.size)
            if rv.transform == I.TR_EXP:
                out.append("    if (%s) { const real e = pm4::m_exp(z); dxdz = %s * e; x = %s + dxdz; }"
                           % (cond, _lit(rv.scale), _lit(rv.shift)))
            elif rv.transform == I.TR_SIGMOID:
                out.append("    if (%s) { const real s = pm4::m_sigmoid<real>(z); x = %s + %s * s; "
                           "dxdz = %s * s * ((real)1 - s); }" % (cond, _lit(rv.shift), _lit(rv.scale), _lit(rv.scale)))
        out.append("  }")
        return "\n".join(out)

    # -- terms ---------------------------------------------------------------
    def _sacc(self, pos: int) -> str:
        if pos not in self.scalar_pos:
            raise AssertionError("scalar position %d missing from ModelIR.scalar_pos" % pos)
        return "sacc_%d" % pos

    def gen_terms(self) -> str:
        body: List[str] = []
        self.lines = body
        op0, plan_idx = 0, 0
        for k, t in enumerate(self.ir.terms):
            if t.gp is not None:
                # marginal GP likelihood: batched float64 Cholesky kernels (csrc/pm4_gp.cu) after this per-chain pass
                body.append("    // ---- term %d: %s (mvn_expquad) -- batched dense path" % (k, t.label))
                continue
            if t.glm is not None:
                # dense Bernoulli-logit term: evaluated for all chains at once on the tensor cores
                # (csrc/pm4_glm.cu, added by libpm4b200 after this per-chain pass); nothing to emit here
                body.append("    // ---- term %d: %s (bernoulli_logit_glm) -- batched tensor-core path" % (k, t.label))
                continue
            self.gen_one_term(k, t, op0, plan_idx)
            op0 += len(t.ops)
            if t.plan is not None:
                plan_idx += 1
        head = ["  template <typename real, class CtxT>",
                "  static __device__ __forceinline__ void terms(const real (&x)[R], real (&gr)[R], CtxT& c) {",
                "    const int lane = c.lane, tl = c.tl; (void)lane; (void)tl;",
                "    const pm4::ModelArgs<real>& A = c.A; (void)A;",
                "    auto c_blob = [&](int op) -> int { return __ldg(A.op_blob + op); }; (void)c_blob;",
                "    real lp = 0;"]
        for pos in self.scalar_pos:
            head.append("    const real sx_%d = c.xs[%d]; real sacc_%d = 0;" % (pos, pos, pos))
        # one transposing butterfly per warp reduces every scalar adjoint and the log-density together;
        # the first lane of the group that ends up holding value k parks the warp's total in the part
        # array (c.my_spart = spart[k] + warp), where the owner of the gradient element adds them up
        n_acc, P = len(self.scalar_pos), self.P
        tail = ["    real red[%d];" % P]
        for k, pos in enumerate(self.scalar_pos):
            tail.append("    red[%d] = sacc_%d;" % (k, pos))
        tail.append("    red[%d] = lp;" % n_acc)
        for k in range(n_acc + 1, P):
            tail.append("    red[%d] = 0;" % k)
        tail.append("    pm4::warp_sum_packed<%d, real>(red, lane);" % P)
        tail.append("    if (c.my_spart >= 0) c.part[c.my_spart] = red[0];")
        tail += ["  }"]
        return "\n".join(head + body + tail)

    @property
    def P(self) -> int:
        P = 1
        while P < len(self.scalar_pos) + 1:
            P *= 2
        if P > 32:
            raise NotImplementedError("more than 31 distinct scalar parameters in one model")
        return P

    def gen_one_term(self, k: int, t: I.Term, op0: int, plan_idx: int):
        ops, kind, tag = t.ops, t.kind, "t%d" % k
        tp = _Tape(ops, op0, tag)
        E = self.emit
        ld_x = I.OP["ld_x"]
        uniform = [o.uniform for o in ops]
        elem_bases = sorted({o.base for o in ops if o.opcode == ld_x and o.mode == I.MODE_ELEM})
        plan = t.plan
        if all(uniform) and t.n == 1:
            shape = "uniform"
        elif plan is not None:
            shape = "planned"
        elif len(elem_bases) == 1:
            shape = "aligned"
        else:
            shape = "general"

        n_par = I.TERM_NPARAMS[kind]
        preg = list(t.param_regs[:n_par])
        scale_reg = preg[1] if kind == I.TERM_KINDS["normal"] else (
            preg[0] if kind in (I.TERM_KINDS["halfnormal"], I.TERM_KINDS["halfcauchy"]) else None)
        scale_uniform = scale_reg is not None and uniform[scale_reg]
        # the factored form needs the scale to enter the term only as the scale parameter
        scale_feeds = scale_reg is not None and (any(o.a == scale_reg or o.b == scale_reg for o in ops)
                                                 or scale_reg in (t.value_reg, preg[0]))
        normal_fast = (kind == I.TERM_KINDS["normal"] and scale_uniform and shape != "uniform"
                       and not scale_feeds)
        K = "%s_K" % tag  # deferred adjoint multiplier 1/scale^2 of the factored Normal form

        def sink(expr: str) -> str:
            return "%s * (%s)" % (K, expr) if normal_fast else expr

        E("// ---- term %d: %s (%s), n=%d, %s%s" % (k, t.label, _KIND[kind], t.n, shape,
                                                      ", factored normal" if normal_fast else ""))
        E("{")
        touched_gx = [False]
        gather_names: Dict[int, str] = {}
        rec_field: Dict[int, str] = {}
        if shape == "planned":
            for j, dst in enumerate(plan.gather_ops):
                gather_names[dst] = "%s_g%d" % (tag, j)
            fields = ["x", "y", "z", "w"]
            for j, dst in enumerate(plan.record_ops):
                rec_field[dst] = ("%s_rec" % tag) if plan.width == 1 else "%s_rec.%s" % (tag, fields[j])

        def xread(o: I.Op, idx: str) -> str:
            if o.mode == I.MODE_SCALAR:
                self._sacc(o.base)
                return "sx_%d" % o.base
            if o.mode == I.MODE_ELEM:
                return "x[r]" if shape == "aligned" else "c.xs[%d + %s]" % (o.base, idx)
            if o.dst in gather_names:
                return gather_names[o.dst] + "v"
            return "c.xs[%d + A.datai[c_blob(%d) + %s]]" % (o.base, op0 + o.dst, idx)

        def dread(o: I.Op, idx: str) -> str:
            if o.dst in rec_field:
                return rec_field[o.dst]
            return "A.dataf[c_blob(%d) + %s]" % (op0 + o.dst, idx if o.mode == I.MODE_ELEM else "0")

        # registers that depend on a free variable (the others need no adjoint)
        depends = [False] * len(ops)
        for o in ops:
            nm = I.OP_NAME[o.opcode]
            if nm in ("ld_x", "ld_z"):
                depends[o.dst] = True
            elif nm not in ("ld_const", "ld_data"):
                depends[o.dst] = depends[o.a] or (o.b >= 0 and depends[o.b])

        # 1. prologue: element-independent forward ops
        for o in ops:
            if o.uniform:
                E("const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "0", xread, dread)), 3)
        if scale_uniform:
            E("const real %s_inv_s = pm4::m_rcp(%s); const real %s_log_s = pm4::m_log(%s);"
              % (tag, tp.v(scale_reg), tag, tp.v(scale_reg)), 3)
        if normal_fast:
            E("const real %s = %s_inv_s * %s_inv_s; real %s_S2 = 0;" % (K, tag, tag, tag), 3)

        # 2. the element body, generated into a buffer (we need to know which uniform registers
        #    receive adjoints before the loop header is written).  With `mask` set, the body is
        #    branch-free: it is evaluated for a padded row too and its contributions are zeroed.
        uacc: Dict[int, bool] = {}
        idx = "i"
        val = tp.v(t.value_reg)

        def gen_body(mask: Optional[str]) -> List[str]:
            saved, loop = self.lines, []
            self.lines = loop

            def m(expr: str) -> str:
                return "(%s ? %s : (real)0)" % (mask, expr) if mask else expr

            for o in ops:
                if not o.uniform:
                    E("const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, idx, xread, dread)), 5)
            adj: Dict[int, List[str]] = {}

            def add_adj(reg: int, expr: str):
                if reg >= 0 and depends[reg]:
                    adj.setdefault(reg, []).append(expr)

            if kind == I.TERM_KINDS["potential"]:
                E("lp += %s;" % m(val), 5)
                add_adj(t.value_reg, m("(real)1"))
            elif normal_fast:
                E("const real %s_r = %s;" % (tag, m("%s - %s" % (val, tp.v(preg[0])))), 5)
                E("%s_S2 += %s_r * %s_r;" % (tag, tag, tag), 5)
                add_adj(preg[0], "%s_r" % tag)
                add_adj(t.value_reg, "-%s_r" % tag)
            else:
                E("real %s_dx, %s_d0, %s_d1; (void)%s_dx; (void)%s_d0; (void)%s_d1;" % ((tag,) * 6), 5)
                if scale_reg is not None and not scale_uniform:
                    E("const real %s_inv_s = pm4::m_rcp(%s); const real %s_log_s = pm4::m_log(%s);"
                      % (tag, tp.v(scale_reg), tag, tp.v(scale_reg)), 5)
                p0 = tp.v(preg[0]) if n_par else ""
                call = {
                    "normal": "normal_lpdf<real>(%s, %s, %s_inv_s, %s_log_s, %s_dx, %s_d0, %s_d1)"
                              % (val, p0, tag, tag, tag, tag, tag),
                    "halfnormal": "halfnormal_lpdf<real>(%s, %s_inv_s, %s_log_s, %s_dx, %s_d0)" % (val, tag, tag, tag, tag),
                    "halfcauchy": "halfcauchy_lpdf<real>(%s, %s_inv_s, %s_log_s, %s_dx, %s_d0)" % (val, tag, tag, tag, tag),
                    "bernoulli": "bernoulli_lpdf<real>(%s, %s, %s_dx, %s_d0)" % (val, p0, tag, tag),
                    "poisson": "poisson_lpdf<real>(%s, %s, %s_dx, %s_d0)" % (val, p0, tag, tag),
                }[_KIND[kind]]
                E("const real %s_lp = pm4::%s;" % (tag, call), 5)
                E("lp += %s;" % m("%s_lp" % tag), 5)
                add_adj(preg[0], m("%s_d0" % tag))
                if kind == I.TERM_KINDS["normal"]:
                    add_adj(preg[1], m("%s_d1" % tag))
                add_adj(t.value_reg, m("%s_dx" % tag))

            def leaf_adj(o: I.Op, g: str):
                if o.mode == I.MODE_SCALAR:  # scalar loads are element-independent: handled via uacc
                    E("%s += %s;" % (self._sacc(o.base), sink(g)), 5)
                elif o.mode == I.MODE_ELEM:
                    if shape == "aligned":
                        E("gr[r] += %s;" % sink(g), 5)
                    else:
                        E("c.gx[%d + %s] += %s;" % (o.base, idx, sink(g)), 5)
                        touched_gx[0] = True
                        self.needs_gx = True
                elif o.dst in gather_names:
                    E("%sa += %s;" % (gather_names[o.dst], g), 5)  # scaled by K when flushed
                else:
                    E("atomicAdd(&c.gx[%d + A.datai[c_blob(%d) + %s]], %s);" % (o.base, op0 + o.dst, idx, sink(g)), 5)
                    touched_gx[0] = True
                    self.needs_gx = True

            for o in reversed(ops):
                if o.dst not in adj:
                    continue
                terms_ = adj.pop(o.dst)
                if o.uniform:
                    uacc[o.dst] = True
                    E("%s_ua%d += %s;" % (tag, o.dst, " + ".join(terms_)), 5)
                    continue
                g = "%s_a%d" % (tag, o.dst)
                E("const real %s = %s;" % (g, " + ".join(terms_)), 5)
                if o.opcode == ld_x:
                    leaf_adj(o, g)
                else:
                    for reg, expr in tp.bwd(o, g):
                        add_adj(reg, expr)
            self.lines = saved
            return loop

        loop = gen_body(None)

        for reg in sorted(uacc):
            E("real %s_ua%d = 0;" % (tag, reg), 3)

        # 3. the loop itself
        n_expr = str(t.n)
        if shape == "uniform":
            E("if (tl == %d) {" % self.uni_lane, 3)
            E("const int i = 0; (void)i;", 5)
            self.lines.extend(loop)
            E("}", 3)
        elif shape == "aligned":
            base = elem_bases[0]
            r_lo, r_hi = base // self.TL, (base + t.n - 1) // self.TL
            E("#pragma unroll", 3)
            E("for (int r = %d; r <= %d; ++r) {" % (r_lo, r_hi), 3)
            E("const int i = tl + %d * r - %d;" % (self.TL, base), 4)
            E("if (i >= 0 && i < %d) {" % t.n, 4)
            self.lines.extend(loop)
            E("}", 4)
            E("}", 3)
        elif shape == "planned":
            n_expr = "n_%s" % tag
            by_dst = {o.dst: o for o in ops}
            TL = self.TL
            n_g = len(gather_names)
            SW = -(-(2 + n_g) // 4)
            rec_t = "typename pm4::Rec<real, %d>::type" % plan.width
            E("const int4 %s_h0 = *reinterpret_cast<const int4*>(c.pi + A.phdr_off + %d);  // n_slices, slices, records, n"
              % (tag, 4 * plan_idx), 3)
            E("const int %s = %s_h0.w;" % (n_expr, tag), 3)
            E("const int4* %s_sl = reinterpret_cast<const int4*>(c.pi + %s_h0.y);" % (tag, tag), 3)
            E("const int4* %s_st = %s_sl + %s_h0.x + tl * %d;" % (tag, tag, tag, SW), 3)
            E("const %s* %s_pr = reinterpret_cast<const %s*>(c.pf + %s_h0.z) + tl;" % (rec_t, tag, rec_t, tag), 3)
            E("#pragma unroll 1", 3)
            E("for (int s = 0; s < %s_h0.x; ++s) {" % tag, 3)
            E("const int4 sl = %s_sl[s];        // rows all lanes walk, rows in the slice, first record row" % tag, 4)
            E("const int4 st = %s_st[s * %d];   // gather target, chunk length, part index per gather op" % (tag, TL * SW), 4)
            for w in range(1, SW):
                E("const int4 st%d = %s_st[s * %d + %d];" % (w, tag, TL * SW, w), 4)
            E("const int tgt = st.x < 0 ? 0 : st.x, len = st.y;", 4)
            for dst, nm in gather_names.items():
                E("const real %sv = c.xs[%d + tgt]; real %sa = 0;" % (nm, by_dst[dst].base, nm), 4)
            E("const %s* rp = %s_pr + sl.z * %d;" % (rec_t, tag, TL), 4)
            E("int i = 0;", 4)
            if self.ir.streamed:
                # records stream from L2 (pools too large to stage): keep the NEXT four rows in flight while
                # the current four are consumed (rows up to the slice's longest chunk exist, zero padded)
                E("%s %s_nx[4];" % (rec_t, tag), 4)
                E("#pragma unroll", 4)
                E("for (int u = 0; u < 4; ++u) %s_nx[u] = rp[(u < sl.y ? u : 0) * %d];" % (tag, TL), 4)
                E("#pragma unroll 1", 4)
                E("for (; i + 4 <= len; i += 4) {   // per-lane trip count: lanes with shorter chunks drop out", 4)
                E("%s %s_rq[4];" % (rec_t, tag), 5)
                E("#pragma unroll", 5)
                E("for (int u = 0; u < 4; ++u) { %s_rq[u] = %s_nx[u]; %s_nx[u] = rp[(i + 4 + u < sl.y ? i + 4 + u : 0) * %d]; }"
                  % (tag, tag, tag, TL), 5)
                E("#pragma unroll", 5)
                E("for (int u = 0; u < 4; ++u) {", 5)
                E("const %s %s_rec = %s_rq[u]; (void)%s_rec;" % (rec_t, tag, tag, tag), 5)
                self.lines.extend(loop)
                E("}", 5)
                E("}", 4)
                E("#pragma unroll", 4)
                E("for (int u = 0; u < 3; ++u) {   // the last rows of the chunk are already in registers", 4)
                E("if (i + u < len) {", 5)
                E("const %s %s_rec = %s_nx[u]; (void)%s_rec;" % (rec_t, tag, tag, tag), 5)
                self.lines.extend(loop)
                E("}", 5)
                E("}", 4)
            else:
                E("#pragma unroll 1", 4)
                E("for (; i + 4 <= len; i += 4) {   // per-lane trip count: lanes with shorter chunks drop out", 4)
                E("%s %s_rq[4];" % (rec_t, tag), 5)
                E("#pragma unroll", 5)
                E("for (int u = 0; u < 4; ++u) %s_rq[u] = rp[(i + u) * %d];" % (tag, TL), 5)
                E("#pragma unroll", 5)
                E("for (int u = 0; u < 4; ++u) {", 5)
                E("const %s %s_rec = %s_rq[u]; (void)%s_rec;" % (rec_t, tag, tag, tag), 5)
                self.lines.extend(loop)
                E("}", 5)
                E("}", 4)
                E("#pragma unroll 1", 4)
                E("for (; i < len; ++i) {", 4)
                E("const %s %s_rec = rp[i * %d]; (void)%s_rec;" % (rec_t, tag, TL, tag), 5)
                self.lines.extend(loop)
                E("}", 4)
            E("if (st.x >= 0) {  // park the chunk's adjoint sums where the owner of the element collects them", 4)
            fields = ["x", "y", "z", "w"]
            for j, (dst, nm) in enumerate(gather_names.items()):
                word, fld = divmod(2 + j, 4)
                E("c.part[st%s.%s] = %s;" % ("" if word == 0 else str(word), fields[fld], sink("%sa" % nm)), 5)
            E("}", 4)
            E("}", 3)
        else:
            n_expr = "n_%s" % tag
            E("const int %s = __ldg(A.term_n + %d);" % (n_expr, k), 3)
            E("for (int i = tl; i < %s; i += %d) {" % (n_expr, self.TL), 3)
            self.lines.extend(loop)
            E("}", 3)

        # 4. factored Normal: the sums enter the log-density and the scale adjoint once
        if normal_fast:
            E("lp += (real)-0.5 * %s * %s_S2;" % (K, tag), 3)
            E("if (tl == %d) lp -= (real)%s * (pm4::K<real>::HALF_LOG_2PI + %s_log_s);" % (self.uni_lane, n_expr, tag), 3)
            if depends[scale_reg]:
                uacc.setdefault(scale_reg, False)
                decl = "" if uacc[scale_reg] else "real "
                # d/d scale = sum (w^2 - 1) / scale = K * S2 / scale - n / scale
                E("%s%s_ua%d %s %s * %s_S2 * %s_inv_s - (tl == %d ? (real)%s * %s_inv_s : (real)0);"
                  % (decl, tag, scale_reg, "+=" if uacc[scale_reg] else "=", K, tag, tag, self.uni_lane, n_expr, tag), 3)
                uacc[scale_reg] = True

        # 5. epilogue: reverse sweep over the element-independent ops with lane-partial adjoints
        u_adj: Dict[int, List[str]] = {}
        for reg in uacc:
            # adjoints that flowed out of the loop carry the deferred factor, the scale adjoint does not
            if normal_fast and reg != scale_reg:
                u_adj[reg] = ["%s * %s_ua%d" % (K, tag, reg)]
            elif normal_fast and reg == scale_reg:
                u_adj[reg] = ["%s_ua%d" % (tag, reg)]
            else:
                u_adj[reg] = ["%s_ua%d" % (tag, reg)]
        for o in reversed(ops):
            if not o.uniform or o.dst not in u_adj or not depends[o.dst]:
                continue
            terms_ = u_adj.pop(o.dst)
            g = "%s_ub%d" % (tag, o.dst)
            E("const real %s = %s;" % (g, " + ".join(terms_)), 3)
            if o.opcode == ld_x:
                E("%s += %s;" % (self._sacc(o.base), g), 3)
            else:
                for reg, expr in tp.bwd(o, g):
                    if reg >= 0 and depends[reg]:
                        u_adj.setdefault(reg, []).append(expr)
        E("}")

    # -- traced outputs --------------------------------------------------------
    def gen_dets(self) -> str:
        out = ["  template <typename real>",
               "  static __device__ void dets(const real* __restrict__ th, real* __restrict__ out, const pm4::ModelArgs<real>& A) {",
               "    auto c_blob = [&](int op) -> int { return __ldg(A.op_blob + op); }; (void)c_blob;",
               "    real xv[D];"]
        for rv in self.ir.rvs:
            rng = "for (int j = %d; j < %d; ++j)" % (rv.offset, rv.offset + rv.size)
            if rv.transform == I.TR_EXP:
                out.append("    %s xv[j] = %s + %s * pm4::m_exp(th[j]);" % (rng, _lit(rv.shift), _lit(rv.scale)))
            elif rv.transform == I.TR_SIGMOID:
                out.append("    %s xv[j] = %s + %s * pm4::m_sigmoid<real>(th[j]);" % (rng, _lit(rv.shift), _lit(rv.scale)))
            else:
                out.append("    %s xv[j] = th[j];" % rng)
        op0 = sum(len(t.ops) for t in self.ir.terms)
        for k, d in enumerate(self.ir.dets):
            tp = _Tape(d.ops, op0, "d%d" % k)

            def xread(o: I.Op, idx: str, _op0=op0) -> str:
                if o.mode == I.MODE_SCALAR:
                    return "xv[%d]" % o.base
                if o.mode == I.MODE_ELEM:
                    return "xv[%d + %s]" % (o.base, idx)
                return "xv[%d + A.datai[c_blob(%d) + %s]]" % (o.base, _op0 + o.dst, idx)

            def dread(o: I.Op, idx: str, _op0=op0) -> str:
                return "A.dataf[c_blob(%d) + %s]" % (_op0 + o.dst, idx if o.mode == I.MODE_ELEM else "0")

            out.append("    // ---- output %d: %s" % (k, d.name))
            out.append("    for (int i = 0; i < %d; ++i) {" % d.n)
            for o in d.ops:
                out.append("      const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "i", xread, dread)))
            out.append("      out[%d + i] = %s;" % (d.out_offset, tp.v(d.value_reg)))
            out.append("    }")
            op0 += len(d.ops)
        out.append("  }")
        return "\n".join(out)

    # -- pointwise log-likelihood -------------------------------------------------
    def gen_loglik(self, predictive: bool = False) -> str:
        """One WARP per state: lanes stride over the elements of every observed term, evaluate the
        forward tape and either the log-density (partials are dead code) or -- ``predictive`` -- one
        draw from the distribution with those parameters, and store it -- no reduction."""
        out = ["  template <typename real>",
               ("  static __device__ void predict(const real* __restrict__ th, real* __restrict__ out, "
                "const pm4::ModelArgs<real>& A, int lane, uint32_t m_lo, uint32_t m_hi, uint32_t k0, uint32_t k1) {")
               if predictive else
               ("  static __device__ void loglik(const real* __restrict__ th, real* __restrict__ out, "
                "const pm4::ModelArgs<real>& A, int lane) {"),
               "    auto c_blob = [&](int op) -> int { return __ldg(A.op_blob + op); }; (void)c_blob;",
               "    auto X = [&](int j) -> real {  // constrained value of element j",
               "      const real z = th[j];"]
        for rv in self.ir.rvs:
            cond = "j >= %d && j < %d" % (rv.offset, rv.offset + rv.size)
            if rv.transform == I.TR_EXP:
                out.append("      if (%s) return %s + %s * pm4::m_exp(z);" % (cond, _lit(rv.shift), _lit(rv.scale)))
            elif rv.transform == I.TR_SIGMOID:
                out.append("      if (%s) return %s + %s * pm4::m_sigmoid<real>(z);" % (cond, _lit(rv.shift), _lit(rv.scale)))
        out += ["      return z;", "    }; (void)X;"]
        op0 = 0
        for k, t in enumerate(self.ir.terms):
            if not t.observed:
                op0 += len(t.ops)
                continue
            if t.gp is not None:
                # one density value per state, produced by the batched path (libpm4b200 refuses these entry points
                # for GP models until it routes them there)
                out.append("    // ---- observed term %d: %s (mvn_expquad): not evaluated per state" % (k, t.label))
                continue
            if t.glm is not None:
                g = t.glm
                out.append("    // ---- observed term %d: %s (bernoulli_logit_glm): eta = X[i, :] . beta" % (k, t.label))
                out.append("    {")
                out.append("      const int n = __ldg(A.term_n + %d);" % k)
                out.append("      for (int i = lane; i < n; i += 32) {")
                out.append("        real eta = 0;")
                out.append("        for (int q = 0; q < %d; ++q) eta += A.dataf[%d + (size_t)i * %d + q] * X(%d + q);"
                           % (g.P, g.x_blob, g.P, g.base))
                out.append("        const real prob = pm4::m_sigmoid<real>(eta), yv = A.dataf[%d + i];" % g.y_blob)
                if predictive:
                    out.append("        out[%d + i] = pm4::draw_value<real>(4, prob, (real)0, m_lo, m_hi, (uint32_t)(%d + i), k0, k1);"
                               % (t.ll_offset, t.ll_offset))
                else:
                    # log Bernoulli(y | sigmoid(eta)) in the cancellation-free form y eta - softplus(eta)
                    out.append("        out[%d + i] = yv * eta - pm4::m_softplus<real>(eta); (void)prob;" % t.ll_offset)
                out.append("      }")
                out.append("    }")
                continue
            tag = "l%d" % k
            tp = _Tape(t.ops, op0, tag)

            def xread(o: I.Op, idx: str, _op0=op0) -> str:
                if o.mode == I.MODE_SCALAR:
                    return "X(%d)" % o.base
                if o.mode == I.MODE_ELEM:
                    return "X(%d + %s)" % (o.base, idx)
                return "X(%d + A.datai[c_blob(%d) + %s])" % (o.base, _op0 + o.dst, idx)

            def dread(o: I.Op, idx: str, _op0=op0) -> str:
                return "A.dataf[c_blob(%d) + %s]" % (_op0 + o.dst, idx if o.mode == I.MODE_ELEM else "0")

            kind = _KIND[t.kind]
            n_par = I.TERM_NPARAMS[t.kind]
            preg = list(t.param_regs[:n_par])
            out.append("    // ---- observed term %d: %s (%s)" % (k, t.label, kind))
            out.append("    {")
            for o in t.ops:
                if o.uniform:
                    out.append("      const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "0", xread, dread)))
            out.append("      const int n = __ldg(A.term_n + %d);" % k)
            out.append("      for (int i = lane; i < n; i += 32) {")
            for o in t.ops:
                if not o.uniform:
                    out.append("        const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "i", xread, dread)))
            val = tp.v(t.value_reg)
            p0 = tp.v(preg[0]) if n_par else ""
            if predictive:
                p1 = tp.v(preg[1]) if n_par > 1 else "(real)0"
                expr = ("pm4::draw_value<real>(%d, %s, %s, m_lo, m_hi, (uint32_t)(%d + i), k0, k1)"
                        % (t.kind, p0 if n_par else "(real)0", p1, t.ll_offset))
            elif kind == "potential":
                expr = val
            else:
                out.append("        real dx_, d0_, d1_; (void)dx_; (void)d0_; (void)d1_;")
                if kind in ("normal", "halfnormal", "halfcauchy"):
                    sreg = tp.v(preg[1] if kind == "normal" else preg[0])
                    out.append("        const real inv_s = pm4::m_rcp(%s), log_s = pm4::m_log(%s);" % (sreg, sreg))
                expr = {
                    "normal": "pm4::normal_lpdf<real>(%s, %s, inv_s, log_s, dx_, d0_, d1_)" % (val, p0),
                    "halfnormal": "pm4::halfnormal_lpdf<real>(%s, inv_s, log_s, dx_, d0_)" % val,
                    "halfcauchy": "pm4::halfcauchy_lpdf<real>(%s, inv_s, log_s, dx_, d0_)" % val,
                    "bernoulli": "pm4::bernoulli_lpdf<real>(%s, %s, dx_, d0_)" % (val, p0),
                    "poisson": "pm4::poisson_lpdf<real>(%s, %s, dx_, d0_)" % (val, p0),
                }[kind]
            out.append("        out[%d + i] = %s;" % (t.ll_offset, expr))
            out.append("      }")
            out.append("    }")
            op0 += len(t.ops)
        out.append("  }")
        return "\n".join(out)

    # -- prior predictive ----------------------------------------------------------
    def gen_prior(self) -> str:
        """One WARP per sample: ancestral sampling.  Free variables are drawn in model order (lanes stride
        over the elements of a variable; parameters read the already drawn CONSTRAINED values from the
        sample's row ``x``), then the observed variables.  Philox element ids: LL_ROW + position for free
        variables, row position for observed ones (oracle: pm4o_prior_predictive)."""
        ld_x = I.OP["ld_x"]
        out = ["  template <typename real>",
               "  static __device__ void prior(real* __restrict__ x, real* __restrict__ obs, const pm4::ModelArgs<real>& A, "
               "int lane, uint32_t m_lo, uint32_t m_hi, uint32_t k0, uint32_t k1) {",
               "    auto c_blob = [&](int op) -> int { return __ldg(A.op_blob + op); }; (void)c_blob;",
               "    auto X = [&](int j) -> real { return x[j]; }; (void)X;",
               "    for (int j = lane; j < D; j += 32) x[j] = 0;",
               "    __syncwarp();"]
        op_of = {}
        op0 = 0
        for k, t in enumerate(self.ir.terms):
            op_of[k] = op0
            op0 += len(t.ops)

        def emit(k, t, target):
            base0 = op_of[k]
            tp = _Tape(t.ops, base0, "p%d" % k)

            def xread(o: I.Op, idx: str) -> str:
                if o.mode == I.MODE_SCALAR:
                    return "X(%d)" % o.base
                if o.mode == I.MODE_ELEM:
                    return "X(%d + %s)" % (o.base, idx)
                return "X(%d + A.datai[c_blob(%d) + %s])" % (o.base, base0 + o.dst, idx)

            def dread(o: I.Op, idx: str) -> str:
                return "A.dataf[c_blob(%d) + %s]" % (base0 + o.dst, idx if o.mode == I.MODE_ELEM else "0")

            n_par = I.TERM_NPARAMS[t.kind]
            preg = list(t.param_regs[:n_par])
            needed = set()  # only the ops the parameters depend on (the value itself is what we draw)
            stack = list(preg)
            while stack:
                r = stack.pop()
                if r in needed or r < 0:
                    continue
                needed.add(r)
                o = t.ops[r]
                if I.OP_NAME[o.opcode] not in ("ld_const", "ld_data", "ld_x", "ld_z"):
                    stack.extend([o.a, o.b])
            out.append("    // ---- %s term %d: %s (%s)" % ("free" if target == "x" else "observed", k, t.label, _KIND[t.kind]))
            out.append("    {")
            for o in t.ops:
                if o.uniform and o.dst in needed:
                    out.append("      const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "0", xread, dread)))
            out.append("      const int n = __ldg(A.term_n + %d);" % k)
            out.append("      for (int i = lane; i < n; i += 32) {")
            for o in t.ops:
                if not o.uniform and o.dst in needed:
                    out.append("        const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "i", xread, dread)))
            p0 = tp.v(preg[0]) if n_par else "(real)0"
            p1 = tp.v(preg[1]) if n_par > 1 else "(real)0"
            if target == "x":
                vo = t.ops[t.value_reg]
                pos = "%d + i" % vo.base if vo.mode == I.MODE_ELEM else "%d" % vo.base
                out.append("        x[%s] = pm4::draw_value<real>(%d, %s, %s, m_lo, m_hi, (uint32_t)(LL_ROW + %s), k0, k1);"
                           % (pos, t.kind, p0, p1, pos))
            else:
                out.append("        if (obs) obs[%d + i] = pm4::draw_value<real>(%d, %s, %s, m_lo, m_hi, (uint32_t)(%d + i), k0, k1);"
                           % (t.ll_offset, t.kind, p0, p1, t.ll_offset))
            out.append("      }")
            out.append("      __syncwarp();")
            out.append("    }")

        def emit_glm(k, t):
            g = t.glm
            out.append("    // ---- observed term %d: %s (bernoulli_logit_glm)" % (k, t.label))
            out.append("    {")
            out.append("      const int n = __ldg(A.term_n + %d);" % k)
            out.append("      for (int i = lane; i < n; i += 32) {")
            out.append("        real eta = 0;")
            out.append("        for (int q = 0; q < %d; ++q) eta += A.dataf[%d + (size_t)i * %d + q] * X(%d + q);"
                       % (g.P, g.x_blob, g.P, g.base))
            out.append("        if (obs) obs[%d + i] = pm4::draw_value<real>(4, pm4::m_sigmoid<real>(eta), (real)0, m_lo, m_hi, "
                       "(uint32_t)(%d + i), k0, k1);" % (t.ll_offset, t.ll_offset))
            out.append("      }")
            out.append("    }")

        for k, t in enumerate(self.ir.terms):
            if t.glm is not None or t.gp is not None:
                continue
            vo = t.ops[t.value_reg]
            if (not t.observed and t.kind != I.TERM_KINDS["potential"] and vo.opcode == ld_x
                    and vo.mode in (I.MODE_ELEM, I.MODE_SCALAR)):
                emit(k, t, "x")
        for k, t in enumerate(self.ir.terms):
            if t.observed and t.glm is not None:
                emit_glm(k, t)
            elif t.gp is not None:
                continue
            elif t.observed:
                emit(k, t, "obs")
        out.append("  }")
        return "\n".join(out)

    def gen_owned(self) -> str:
        from .codegen_owned import OwnedGen, not_owned_members

        if self.ir.owned is None:
            return not_owned_members()
        return OwnedGen(self.ir, self.ir.owned).members()

    def source(self) -> str:
        ir = self.ir
        name = "Model_%s" % ir.struct_hash_hex
        n_ops = sum(len(t.ops) for t in ir.terms) + sum(len(d.ops) for d in ir.dets)
        terms_src = self.gen_terms()
        parts = [
            "// generated by pymc4_b200/codegen.py -- do not edit.  IR structure hash %s" % ir.struct_hash_hex,
            "// free variables: " + ", ".join("%s[%d]@%d" % (rv.name, rv.size, rv.offset) for rv in ir.rvs),
            "// team: %d warps per chain" % self.T,
            '#include "pm4_common.cuh"',
            "struct %s {" % name,
            "  static constexpr int D = %d, T = %d, R = %d, DP = %d, DS = %d, DET_ROW = %d, N_TERMS = %d, N_OPS = %d, N_PLANS = %d;"
            % (self.D, self.T, self.R, self.TL * self.R, -(-self.D // 32) * 32, ir.det_row, len(ir.terms), n_ops,
               len(ir.plans)),
            "  static constexpr int N_ACC = %d, P = %d;  // scalar adjoints reduced with the log-density, packed width"
            % (len(self.scalar_pos), self.P),
            "  static constexpr int PK = %d;  // fixed part entries per gradient element" % int(ir.part_k),
            "  static constexpr bool NEEDS_GX = %s;  // some term scatters adjoints with atomics"
            % ("true" if self.needs_gx else "false"),
            "  static constexpr unsigned long long HASH = 0x%sull;" % ir.struct_hash_hex,
            self.gen_constrain(),
            terms_src,
            self.gen_dets(),
            "  static constexpr int LL_ROW = %d;  // pointwise log-likelihood scalars per state" % ir.ll_row,
            self.gen_loglik(),
            self.gen_loglik(predictive=True),
            self.gen_prior(),
            self.gen_owned(),
            "};",
            "#define PM4_MODEL %s" % name,
            "#define PM4_HAS_F64 %d" % (1 if self.has_f64 else 0),
            '#include "pm4_kernels.cuh"',
            "",
        ]
        return "\n".join(parts)


def generate_source(ir: I.ModelIR, has_f64: bool = False) -> str:
    return _ModuleGen(ir, has_f64).source()


_gen_hash_cache: Optional[str] = None


def generator_hash() -> str:
    """Content hash of everything a built module depends on besides the model structure: the
    device headers and the generator (a module built by an older generator may read execution
    plans in an older layout).  Part of the module file name, so stale modules are never loaded
    and file modification times (which do not survive a copy to the GPU box) play no role."""
    global _gen_hash_cache
    if _gen_hash_cache is None:
        import hashlib

        h = hashlib.sha256()
        deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
        deps += [os.path.join(_HERE, f) for f in ("codegen.py", "codegen_owned.py", "owned.py", "plan.py", "ir.py")]
        for d in deps:
            with open(d, "rb") as f:
                h.update(f.read())
        h.update((" ".join(NVCC_FLAGS) + " MAXT=%d" % MAXT).encode())
        _gen_hash_cache = h.hexdigest()[:8]
    return _gen_hash_cache


def module_path(ir: I.ModelIR, has_f64: bool = False) -> str:
    return os.path.join(CACHE_DIR, "pm4m_%s_%s%s.so" % (ir.struct_hash_hex, generator_hash(),
                                                        "_f64" if has_f64 else ""))


def find_nvcc() -> str:
    for cand in (os.environ.get("PM4_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the specialised sampler module cannot be built")


def _csrc_mtime() -> float:
    """Newest modification time of anything a built module depends on (device headers and the
    generator itself: a module built by an older generator may read plans in an older layout)."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(_HERE, f) for f in ("codegen.py", "codegen_owned.py", "owned.py", "plan.py", "ir.py")]
    return max(os.path.getmtime(d) for d in deps)


def build_module(ir: I.ModelIR, has_f64: bool = False, force: bool = False, verbose: bool = False,
                 extra_flags: Optional[List[str]] = None) -> str:
    """Return the path of the compiled module for this IR structure, building it if needed."""
    path = module_path(ir, has_f64)
    alt = module_path(ir, True)  # a module with double precision also serves float requests
    with _build_lock:  # one lock per module file: different models compile concurrently
        lock = _path_locks.setdefault(path, threading.Lock())
    with lock:
        for cand in (path, alt):
            if not force and os.path.exists(cand):
                return cand
        os.makedirs(CACHE_DIR, exist_ok=True)
        src_path = path[:-3] + ".cu"
        with open(src_path, "w") as f:
            f.write(generate_source(ir, has_f64))
        tmp = tempfile.NamedTemporaryFile(prefix="pm4m_", suffix=".so", dir=CACHE_DIR, delete=False).name
        cmd = ([find_nvcc()] + NVCC_FLAGS + ["-DPM4_MAXT=%d" % MAXT] + list(extra_flags or [])
               + ["-I", CSRC, "-shared", "-o", tmp, src_path])
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise RuntimeError("nvcc failed for %s:\n%s" % (src_path, proc.stdout[-6000:]))
        if verbose:
            print(proc.stdout)
        os.replace(tmp, path)
        return path
