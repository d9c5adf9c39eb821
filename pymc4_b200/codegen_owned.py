"""Code generation for the owner-computes layout of the warp-per-chain path (pymc4_b200/owned.py).

Emits, into the generated ``Model`` struct, ``constrain_o`` and ``terms_o``: the same joint log-density and
reverse sweep as ``Model::terms`` (codegen.py) -- the reference's ``collect_log_prob``
(/root/reference/pymc4/flow/executor.py:171-187) over the TFP closed forms of SURVEY App. B -- but on a state
held in the internal register layout:

* element-independent terms (the scalar priors, the log-det potentials) are evaluated by ALL lanes without a
  branch; their contributions are uniform and enter the reduction once, through lane 0;
* element-wise terms over a free vector walk the registers that hold the vector (``perm`` tells which element);
* the gather term walks its owned slices with the group's parameters in registers -- two observations per
  packed-FP32 instruction where the tape allows -- then the foreign pieces (parameters and sums by shuffle);
* one packed butterfly reduces the scalar adjoints and the log-density; total k lands on the lane that owns
  scalar k.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Tuple

from . import ir as I
from .owned import OwnedLayout

_PAIR_OPS = {"add", "sub", "mul", "neg", "square"}


def _fold(name: str, a: Optional[float], b: Optional[float]) -> Optional[float]:
    """value of a tape op on literal operands (the device math shims are inline PTX, which nvcc does not fold)"""
    import math

    try:
        if name == "neg": return -a
        if name == "exp": return math.exp(a)
        if name == "log": return math.log(a)
        if name == "log1p": return math.log1p(a)
        if name == "sqrt": return math.sqrt(a)
        if name == "square": return a * a
        if name == "recip": return 1.0 / a
        if name == "abs": return abs(a)
        if name == "tanh": return math.tanh(a)
        if name == "lgamma": return math.lgamma(a)
        if name == "sigmoid": return 1.0 / (1.0 + math.exp(-a))
        if name == "softplus": return math.log1p(math.exp(-abs(a))) + max(a, 0.0)
        if name == "add": return a + b
        if name == "sub": return a - b
        if name == "mul": return a * b
        if name == "div": return a / b
        if name == "pow": return a ** b
    except (ValueError, OverflowError, ZeroDivisionError, TypeError):
        return None
    return None


def _lit(v: float) -> str:
    from .codegen import _lit as L
    return L(v)


class OwnedGen:
    def __init__(self, ir: I.ModelIR, lay: OwnedLayout):
        from .codegen import _KIND, _Tape  # late import: codegen imports this module
        self.ir, self.lay = ir, lay
        self._Tape, self._KIND = _Tape, _KIND
        self.lines: List[str] = []
        self.scalar_pos = list(ir.scalar_pos)

    def emit(self, s: str = "", ind: int = 2):
        self.lines.append("  " * ind + s)

    # -- which internal registers hold a theta range ---------------------------
    def regs_of(self, base: int, n: int) -> List[int]:
        lay = self.lay
        for v, b in enumerate(lay.vec_bases):
            if base >= b and base + n <= b + lay.G:
                return list(range(v * lay.KS, (v + 1) * lay.KS))
        regs = sorted({lay.M0 + lay.misc_pos[j] // 32 for j in range(base, base + n)})
        return regs

    # -- constrain ---------------------------------------------------------------
    def gen_constrain(self) -> str:
        lay, ir = self.lay, self.ir
        out = ["  template <typename real>",
               "  static __device__ __forceinline__ void constrain_o(const real (&z)[RI], real (&x)[RI], real (&dxdz)[RI], int lane) {",
               "    (void)lane;"]
        rv_of = {}
        for rv in ir.rvs:
            for j in range(rv.offset, rv.offset + rv.size):
                rv_of[j] = rv
        for r in range(lay.RI):
            out.append("    x[%d] = z[%d]; dxdz[%d] = (real)1;" % (r, r, r))
            if r < lay.M0:
                rv = rv_of[lay.vec_bases[r // lay.KS]]
                groups = {(rv.transform, rv.shift, rv.scale): 0xFFFFFFFF}
            else:
                groups = {}
                for j, q in lay.misc_pos.items():
                    if q // 32 == r - lay.M0:
                        rv = rv_of[j]
                        key = (rv.transform, rv.shift, rv.scale)
                        groups[key] = groups.get(key, 0) | (1 << (q % 32))
            for (tr, shift, scale), mask in sorted(groups.items()):
                cond = "true" if mask == 0xFFFFFFFF else "(0x%08xu >> lane) & 1u" % mask
                if tr == I.TR_EXP:
                    out.append("    if (%s) { const real e = pm4::m_exp(z[%d]); dxdz[%d] = %s * e; x[%d] = %s + dxdz[%d]; }"
                               % (cond, r, r, _lit(scale), r, _lit(shift), r))
                elif tr == I.TR_SIGMOID:
                    out.append("    if (%s) { const real s = pm4::m_sigmoid<real>(z[%d]); x[%d] = %s + %s * s; "
                               "dxdz[%d] = %s * s * ((real)1 - s); }" % (cond, r, r, _lit(shift), _lit(scale), r, _lit(scale)))
        out.append("  }")
        return "\n".join(out)

    def id_regs(self) -> List[bool]:
        """register r holds identity-transformed elements only (dxdz == 1 exactly)"""
        lay, ir = self.lay, self.ir
        rv_of = {}
        for rv in ir.rvs:
            for j in range(rv.offset, rv.offset + rv.size):
                rv_of[j] = rv
        out = []
        for r in range(lay.RI):
            if r < lay.M0:
                out.append(rv_of[lay.vec_bases[r // lay.KS]].transform == I.TR_IDENTITY)
            else:
                out.append(all(rv_of[j].transform == I.TR_IDENTITY for j, q in lay.misc_pos.items() if q // 32 == r - lay.M0))
        return out

    # -- terms -------------------------------------------------------------------
    def gen_terms(self) -> str:
        lay = self.lay
        body: List[str] = []
        self.lines = body
        # element-independent terms of identical structure that differ only in WHICH scalar they read (the N(0, 1)
        # priors, the HalfCauchy priors, the log-det potentials) are evaluated together, each on the lane that owns
        # its scalar: one pass per group instead of one per term, adjoint straight into that lane's gradient register
        ld_x = I.OP["ld_x"]
        groups: Dict[tuple, List[int]] = {}
        op0s: List[int] = []
        op0 = 0
        for k, t in enumerate(self.ir.terms):
            op0s.append(op0)
            op0 += len(t.ops)
            if not (all(o.uniform for o in t.ops) and t.n == 1):
                continue
            xs = [o for o in t.ops if o.opcode == ld_x]
            if len(xs) != 1 or xs[0].mode != I.MODE_SCALAR or xs[0].base not in self.scalar_pos:
                continue
            if any(I.OP_NAME[o.opcode] in ("ld_data", "ld_z") for o in t.ops):
                continue
            sig = (t.kind, t.value_reg, tuple(t.param_regs),
                   tuple((o.opcode, o.dst, o.a, o.b, o.mode, float(o.imm)) for o in t.ops))
            groups.setdefault(sig, []).append(k)
        lanevec: Dict[int, Optional[List[int]]] = {}
        for sig, ks in groups.items():
            lanevec[ks[0]] = ks
            for k in ks[1:]:
                lanevec[k] = None  # folded into the group's first term
        for k, t in enumerate(self.ir.terms):
            if k in lanevec:
                if lanevec[k] is not None:
                    self.gen_one_term(k, t, op0s[k], group=lanevec[k])
                continue
            self.gen_one_term(k, t, op0s[k])
        n_acc, P, per = len(self.scalar_pos), lay.P, lay.per
        head = ["  template <typename real, class OC>",
                "  static __device__ __forceinline__ real terms_o(const real (&x)[RI], real (&gr)[RI], const OC& c) {",
                "    const int lane = c.lane; (void)lane;",
                "    const pm4::ModelArgs<real>& A = c.A; (void)A;",
                "    auto c_blob = [&](int op) -> int { return __ldg(A.op_blob + op); }; (void)c_blob;",
                "    real lp = 0, lp_u = 0;  // lane-partial and chain-uniform parts of the log-density"]
        for k, pos in enumerate(self.scalar_pos):
            head.append("    const real sx_%d = __shfl_sync(pm4::FULL, x[%d], %d); real sacc_%d = 0, usacc_%d = 0;"
                        % (pos, lay.M0, per * k, pos, pos))
        tail = ["    // uniform contributions enter the reduction once, through lane 0",
                "    if (lane == 0) {",
                "      lp += lp_u;"]
        for pos in self.scalar_pos:
            tail.append("      sacc_%d += usacc_%d;" % (pos, pos))
        tail.append("    }")
        tail.append("    real red[%d];" % P)
        for k, pos in enumerate(self.scalar_pos):
            tail.append("    red[%d] = sacc_%d;" % (k, pos))
        tail.append("    red[%d] = lp;" % n_acc)
        for k in range(n_acc + 1, P):
            tail.append("    red[%d] = 0;" % k)
        tail.append("    pm4::warp_sum_packed<%d, real>(red, lane);" % P)
        if n_acc:
            tail.append("    // total k sits on lanes [%d k, %d k + %d): lane %d k of register %d owns scalar k"
                        % (per, per, per, per, lay.M0))
            tail.append("    if ((lane %% %d) == 0 && lane < %d) gr[%d] += red[0];" % (per, per * n_acc, lay.M0))
        tail.append("    return __shfl_sync(pm4::FULL, red[0], %d);" % (per * n_acc))
        tail.append("  }")
        return "\n".join(head + body + tail)

    def gen_one_term(self, k: int, t: I.Term, op0: int, group: Optional[List[int]] = None):
        ir, lay = self.ir, self.lay
        ops, kind, tag = t.ops, t.kind, "t%d" % k
        tp = self._Tape(ops, op0, tag)
        E = self.emit
        ld_x = I.OP["ld_x"]
        uniform = [o.uniform for o in ops]
        elem_bases = sorted({o.base for o in ops if o.opcode == ld_x and o.mode == I.MODE_ELEM})
        plan = t.plan
        if all(uniform) and t.n == 1:
            shape = "uniform"
        elif plan is not None:
            shape = "planned"
        else:
            assert len(elem_bases) == 1, "owned layout: general terms are not eligible"
            shape = "aligned"
        KIND = self._KIND
        n_par = I.TERM_NPARAMS[kind]
        preg = list(t.param_regs[:n_par])
        scale_reg = preg[1] if kind == I.TERM_KINDS["normal"] else (
            preg[0] if kind in (I.TERM_KINDS["halfnormal"], I.TERM_KINDS["halfcauchy"]) else None)
        scale_uniform = scale_reg is not None and uniform[scale_reg]
        scale_feeds = scale_reg is not None and (any(o.a == scale_reg or o.b == scale_reg for o in ops)
                                                 or scale_reg in (t.value_reg, preg[0]))
        normal_fast = (kind == I.TERM_KINDS["normal"] and scale_uniform and shape != "uniform" and not scale_feeds)
        Kf = "%s_K" % tag
        lp_sink = "lp_u" if shape == "uniform" else "lp"
        if group is not None:
            lp_sink = "%s_lv_lp" % tag

        def sacc(pos: int) -> str:
            assert pos in self.scalar_pos
            if group is not None:
                return "%s_lv_g" % tag
            return ("usacc_%d" if shape == "uniform" else "sacc_%d") % pos

        def sink(expr: str) -> str:
            return "%s * (%s)" % (Kf, expr) if normal_fast else expr

        E("// ---- term %d: %s (%s), n=%d, %s%s" % (k, t.label, KIND[kind], t.n, shape,
                                                      ", factored normal" if normal_fast else ""))
        E("{")
        if group is not None:
            mask = 0
            for kk in group:
                pos = [o.base for o in ir.terms[kk].ops if o.opcode == ld_x][0]
                mask |= 1 << (lay.per * self.scalar_pos.index(pos))
            E("// evaluated by every lane on ITS element of register %d; terms %s live on the lanes of mask 0x%08x"
              % (lay.M0, ", ".join(str(kk) for kk in group), mask), 3)
            E("const bool %s_on = (0x%08xu >> lane) & 1u;" % (tag, mask), 3)
            E("real %s_lv_lp = 0, %s_lv_g = 0;" % (tag, tag), 3)
        by_dst = {o.dst: o for o in ops}
        gather_names: Dict[int, str] = {}
        rec_field: Dict[int, str] = {}
        if shape == "planned":
            for dst in plan.gather_ops:
                gather_names[dst] = "%s_g%d" % (tag, lay.vec_bases.index(by_dst[dst].base))
            for j, dst in enumerate(plan.record_ops):
                rec_field[dst] = "%s_c%d" % (tag, j)

        def xread(o: I.Op, idx: str) -> str:
            if o.mode == I.MODE_SCALAR:
                assert o.base in self.scalar_pos
                if group is not None:
                    return "x[%d]" % lay.M0
                return "sx_%d" % o.base
            if o.mode == I.MODE_ELEM:
                assert shape == "aligned"
                return "x[r]"
            assert o.dst in gather_names
            return gather_names[o.dst] + "v"

        def dread(o: I.Op, idx: str) -> str:
            if o.dst in rec_field:
                return rec_field[o.dst]
            return "A.dataf[c_blob(%d) + %s]" % (op0 + o.dst, idx if o.mode == I.MODE_ELEM else "0")

        depends = [False] * len(ops)
        for o in ops:
            nm = I.OP_NAME[o.opcode]
            if nm in ("ld_x", "ld_z"):
                depends[o.dst] = True
            elif nm not in ("ld_const", "ld_data"):
                depends[o.dst] = depends[o.a] or (o.b >= 0 and depends[o.b])

        # 1. prologue: element-independent forward ops (all lanes); ops on literals are folded here
        cval: Dict[int, float] = {}
        for o in ops:
            if o.uniform:
                nm = I.OP_NAME[o.opcode]
                if nm == "ld_const":
                    cval[o.dst] = float(o.imm)
                elif nm not in ("ld_data", "ld_x", "ld_z") and o.a in cval and (o.b < 0 or o.b in cval):
                    v = _fold(nm, cval[o.a], cval.get(o.b))
                    if v is not None and v == v and abs(v) != float("inf"):
                        cval[o.dst] = v
                if o.dst in cval:
                    E("const real %s = %s;" % (tp.v(o.dst), _lit(cval[o.dst])), 3)
                else:
                    E("const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "0", xread, dread)), 3)
        if scale_uniform:
            import math
            if scale_reg in cval and cval[scale_reg] > 0:
                E("const real %s_inv_s = %s; const real %s_log_s = %s;"
                  % (tag, _lit(1.0 / cval[scale_reg]), tag, _lit(math.log(cval[scale_reg]))), 3)
            else:
                E("const real %s_inv_s = pm4::m_rcp(%s); const real %s_log_s = pm4::m_log(%s);"
                  % (tag, tp.v(scale_reg), tag, tp.v(scale_reg)), 3)
        if normal_fast:
            E("const real %s = %s_inv_s * %s_inv_s; real %s_S2 = 0;" % (Kf, tag, tag, tag), 3)

        uacc: Dict[int, bool] = {}
        val = tp.v(t.value_reg)

        # 2. scalar element body (strings over `real`)
        def gen_body() -> List[str]:
            saved, loop = self.lines, []
            self.lines = loop
            for o in ops:
                if not o.uniform:
                    E("const real %s = %s;" % (tp.v(o.dst), tp.fwd(o, "i", xread, dread)), 5)
            adj: Dict[int, List[str]] = {}

            def add_adj(reg: int, expr: str):
                if reg >= 0 and depends[reg]:
                    adj.setdefault(reg, []).append(expr)

            if kind == I.TERM_KINDS["potential"]:
                E("%s += %s;" % (lp_sink, val), 5)
                add_adj(t.value_reg, "(real)1")
            elif normal_fast:
                E("const real %s_r = %s - %s;" % (tag, val, tp.v(preg[0])), 5)
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
                }[KIND[kind]]
                E("const real %s_lp = pm4::%s;" % (tag, call), 5)
                E("%s += %s_lp;" % (lp_sink, tag), 5)
                add_adj(preg[0], "%s_d0" % tag)
                if kind == I.TERM_KINDS["normal"]:
                    add_adj(preg[1], "%s_d1" % tag)
                add_adj(t.value_reg, "%s_dx" % tag)

            def leaf_adj(o: I.Op, g: str):
                if o.mode == I.MODE_SCALAR:
                    E("%s += %s;" % (sacc(o.base), sink(g)), 5)
                elif o.mode == I.MODE_ELEM:
                    E("gr[r] += %s;" % sink(g), 5)
                else:
                    E("%sa += %s;" % (gather_names[o.dst], g), 5)  # scaled by K when flushed

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

        loop = gen_body()
        for reg in sorted(uacc):
            E("real %s_ua%d = 0;" % (tag, reg), 3)

        n_expr = str(t.n)
        if shape == "uniform":
            E("{", 3)
            E("const int i = 0; (void)i;", 5)
            self.lines.extend(loop)
            E("}", 3)
        elif shape == "aligned":
            base = elem_bases[0]
            for r in self.regs_of(base, t.n):
                E("{", 3)
                E("constexpr int r = %d;" % r, 4)
                E("const int i = c.pj[r] - %d;" % base, 4)
                E("if (i >= 0 && i < %d) {" % t.n, 4)
                self.lines.extend(loop)
                E("}", 4)
                E("}", 3)
        else:
            n_expr = "c.n_term"
            self.gen_walk(t, tp, tag, loop, gather_names, rec_field, normal_fast, Kf, uacc, depends, preg, uniform, sink)

        # 4. factored Normal: the sums enter the log-density and the scale adjoint once
        if normal_fast:
            E("lp += (real)-0.5 * %s * %s_S2;" % (Kf, tag), 3)
            E("if (lane == 0) lp -= (real)%s * (pm4::K<real>::HALF_LOG_2PI + %s_log_s);" % (n_expr, tag), 3)
            if depends[scale_reg]:
                uacc.setdefault(scale_reg, False)
                decl = "" if uacc[scale_reg] else "real "
                E("%s%s_ua%d %s %s * %s_S2 * %s_inv_s - (lane == 0 ? (real)%s * %s_inv_s : (real)0);"
                  % (decl, tag, scale_reg, "+=" if uacc[scale_reg] else "=", Kf, tag, tag, n_expr, tag), 3)
                uacc[scale_reg] = True

        # 5. epilogue: reverse sweep over the element-independent ops
        u_adj: Dict[int, List[str]] = {}
        for reg in uacc:
            if normal_fast and reg != scale_reg:
                u_adj[reg] = ["%s * %s_ua%d" % (Kf, tag, reg)]
            else:
                u_adj[reg] = ["%s_ua%d" % (tag, reg)]
        for o in reversed(ops):
            if not o.uniform or o.dst not in u_adj or not depends[o.dst]:
                continue
            terms_ = u_adj.pop(o.dst)
            g = "%s_ub%d" % (tag, o.dst)
            E("const real %s = %s;" % (g, " + ".join(terms_)), 3)
            if o.opcode == ld_x:
                E("%s += %s;" % (sacc(o.base), g), 3)
            else:
                for reg, expr in tp.bwd(o, g):
                    if reg >= 0 and depends[reg]:
                        u_adj.setdefault(reg, []).append(expr)
        if group is not None:
            E("lp += %s_on ? %s_lv_lp : (real)0;" % (tag, tag), 3)
            E("gr[%d] += %s_on ? %s_lv_g : (real)0;" % (lay.M0, tag, tag), 3)
        E("}")

    # -- the gather term: owned slices, then foreign pieces -------------------------
    def pair_ok(self, t: I.Term, normal_fast: bool, uniform: List[bool]) -> bool:
        if not normal_fast:
            return False
        for o in t.ops:
            if o.uniform:
                continue
            nm = I.OP_NAME[o.opcode]
            if nm in ("ld_data", "ld_x"):
                if nm == "ld_data" and o.mode != I.MODE_ELEM:
                    return False
                continue
            if nm not in _PAIR_OPS:
                return False
        return True

    def gen_pair_body(self, t, tp, tag, gather_names, rec_field, uacc, depends, preg, uniform) -> List[str]:
        """Two observations per instruction: the non-uniform part of the tape over pm4::Pk<real> (float: packed
        FP32 fma / add / mul), factored-Normal form.  Uniform registers enter as broadcast pairs `<v>_p`."""
        ops = t.ops
        ld_x = I.OP["ld_x"]
        out: List[str] = []
        use_count: Dict[int, int] = {}
        for o in ops:
            for s in (o.a, o.b):
                if s >= 0 and I.OP_NAME[o.opcode] not in ("ld_const", "ld_data", "ld_x", "ld_z"):
                    use_count[s] = use_count.get(s, 0) + 1
        for reg in (t.value_reg, preg[0]):
            use_count[reg] = use_count.get(reg, 0) + 1

        def pv(reg: int) -> str:  # pair name of a register
            return ("%s_p%d" if uniform[reg] else "%s_q%d") % (tag, reg)

        by_dst = {o.dst: o for o in ops}
        fused: set = set()  # mul results folded into the add / sub that consumes them

        def is_fusable_mul(reg: int) -> bool:
            o = by_dst.get(reg)
            return (o is not None and not o.uniform and I.OP_NAME[o.opcode] == "mul" and use_count.get(reg, 0) == 1)

        # forward
        for o in ops:
            if o.uniform:
                continue
            nm = I.OP_NAME[o.opcode]
            d = pv(o.dst)
            if nm == "ld_data":
                out.append("const P2 %s = %s;" % (d, rec_field[o.dst]))
            elif nm == "ld_x":
                out.append("const P2 %s = %svp;" % (d, gather_names[o.dst]))
            elif nm == "mul":
                if is_fusable_mul(o.dst) and self._consumer_is_addsub(ops, o.dst):
                    fused.add(o.dst)
                    continue
                out.append("const P2 %s = pm4::pk_mul(%s, %s);" % (d, pv(o.a), pv(o.b)))
            elif nm == "add":
                if o.b in fused:
                    m = by_dst[o.b]
                    out.append("const P2 %s = pm4::pk_fma(%s, %s, %s);" % (d, pv(m.a), pv(m.b), pv(o.a)))
                elif o.a in fused:
                    m = by_dst[o.a]
                    out.append("const P2 %s = pm4::pk_fma(%s, %s, %s);" % (d, pv(m.a), pv(m.b), pv(o.b)))
                else:
                    out.append("const P2 %s = pm4::pk_add(%s, %s);" % (d, pv(o.a), pv(o.b)))
            elif nm == "sub":
                if o.b in fused:
                    m = by_dst[o.b]
                    out.append("const P2 %s = pm4::pk_fnma(%s, %s, %s);" % (d, pv(m.a), pv(m.b), pv(o.a)))
                elif o.a in fused:
                    m = by_dst[o.a]
                    out.append("const P2 %s = pm4::pk_fms(%s, %s, %s);" % (d, pv(m.a), pv(m.b), pv(o.b)))
                else:
                    out.append("const P2 %s = pm4::pk_sub(%s, %s);" % (d, pv(o.a), pv(o.b)))
            elif nm == "neg":
                out.append("const P2 %s = pm4::pk_neg(%s);" % (d, pv(o.a)))
            elif nm == "square":
                out.append("const P2 %s = pm4::pk_mul(%s, %s);" % (d, pv(o.a), pv(o.a)))
            else:
                raise AssertionError(nm)
        # factored Normal: r = value - loc
        out.append("const P2 %s_r = pm4::pk_sub(%s, %s);" % (tag, pv(t.value_reg), pv(preg[0])))
        out.append("%s_S2p = pm4::pk_fma(%s_r, %s_r, %s_S2p);" % (tag, tag, tag, tag))
        # reverse sweep: adjoint terms are (sign, name) or (sign, name, name)
        adj: Dict[int, List[Tuple]] = {}

        def add_adj(reg: int, term: Tuple):
            if reg >= 0 and depends[reg]:
                adj.setdefault(reg, []).append(term)

        add_adj(preg[0], (1, "%s_r" % tag))
        add_adj(t.value_reg, (-1, "%s_r" % tag))

        def accumulate(acc: str, terms_: List[Tuple], declare: bool) -> None:
            """acc (+)= sum of terms, products folded into fused multiply-adds"""
            cur = acc if not declare else None
            for sgn, *names in terms_:
                if len(names) == 1:
                    if cur is None:
                        expr = names[0] if sgn > 0 else "pm4::pk_neg(%s)" % names[0]
                    else:
                        expr = ("pm4::pk_add(%s, %s)" if sgn > 0 else "pm4::pk_sub(%s, %s)") % (cur, names[0])
                else:
                    if cur is None:
                        expr = "pm4::pk_mul(%s, %s)" % (names[0], names[1])
                        if sgn < 0:
                            expr = "pm4::pk_neg(%s)" % expr
                    else:
                        expr = ("pm4::pk_fma(%s, %s, %s)" if sgn > 0 else "pm4::pk_fnma(%s, %s, %s)") % (names[0], names[1], cur)
                if declare and cur is None:
                    out.append("P2 %s = %s;" % (acc, expr))
                else:
                    out.append("%s = %s;" % (acc, expr))
                cur = acc

        for o in reversed(ops):
            if o.dst not in adj:
                continue
            terms_ = adj.pop(o.dst)
            nm = I.OP_NAME[o.opcode]
            if o.uniform:
                uacc[o.dst] = True
                accumulate("%s_uap%d" % (tag, o.dst), terms_, False)
                continue
            if nm == "ld_x":
                accumulate("%sa" % gather_names[o.dst], terms_, False)
                continue
            if nm == "ld_data":
                continue
            if len(terms_) == 1 and len(terms_[0]) == 2:
                sgn, g = terms_[0]
            else:
                g = "%s_b%d" % (tag, o.dst)
                accumulate(g, terms_, True)
                sgn = 1
            if nm == "add":
                add_adj(o.a, (sgn, g)); add_adj(o.b, (sgn, g))
            elif nm == "sub":
                add_adj(o.a, (sgn, g)); add_adj(o.b, (-sgn, g))
            elif nm == "mul":
                add_adj(o.a, (sgn, g, pv(o.b))); add_adj(o.b, (sgn, g, pv(o.a)))
            elif nm == "neg":
                add_adj(o.a, (-sgn, g))
            elif nm == "square":
                two_a = "%s_t%d" % (tag, o.dst)
                out.append("const P2 %s = pm4::pk_add(%s, %s);" % (two_a, pv(o.a), pv(o.a)))
                add_adj(o.a, (sgn, g, two_a))
        return out

    @staticmethod
    def _consumer_is_addsub(ops, reg: int) -> bool:
        for o in ops:
            nm = I.OP_NAME[o.opcode]
            if nm in ("ld_const", "ld_data", "ld_x", "ld_z"):
                continue
            if o.a == reg or o.b == reg:
                return nm in ("add", "sub") and not o.uniform
        return False

    def gen_walk(self, t, tp, tag, loop, gather_names, rec_field, normal_fast, Kf, uacc, depends, preg, uniform, sink):
        """ONE copy of the row walk inside a run-time loop over the slices (owned, then foreign): uniform trip counts,
        rows beyond a lane's chunk are predicated off, eight pair rows per computed jump."""
        lay = self.lay
        E = self.emit
        W, KS, NV = lay.width, lay.KS, lay.NV
        ops = t.ops
        vecs = sorted(set(gather_names.values()), key=lambda nm: int(nm.rsplit("_g", 1)[1]))  # <tag>_g<v>
        vidx = {nm: int(nm.rsplit("_g", 1)[1]) for nm in vecs}
        pair = self.pair_ok(t, normal_fast, uniform)
        rec_t = "typename pm4::PairRec<real, %d>::type" % W
        E("// records: [pair row][lane] of %d column(s) x 2 observations" % W, 3)
        E("using P2 = pm4::Pk<real>; (void)sizeof(P2);", 3)
        E("const %s* const %s_rec = reinterpret_cast<const %s*>(c.rec) + lane;" % (rec_t, tag, rec_t), 3)
        pair_lines: List[str] = []
        uacc_pair: Dict[int, bool] = {}
        if pair:
            pair_lines = self.gen_pair_body(t, tp, tag, gather_names, rec_field, uacc_pair, depends, preg, uniform)
            text = "\n".join(pair_lines)
            for o in ops:
                if o.uniform and re.search(r"\b%s_p%d\b" % (tag, o.dst), text):
                    E("const P2 %s_p%d = pm4::pk_bcast<real>(%s);" % (tag, o.dst, tp.v(o.dst)), 3)
            E("P2 %s_S2p = pm4::pk_bcast<real>((real)0);" % tag, 3)
            for reg in sorted(uacc_pair):
                E("P2 %s_uap%d = pm4::pk_bcast<real>((real)0);" % (tag, reg), 3)
                if reg not in uacc:
                    uacc[reg] = True
                    E("real %s_ua%d = 0;" % (tag, reg), 3)

        def sel(regs: List[int]) -> str:
            """x[regs[s]] for the run-time (warp-uniform) slice index s"""
            expr = "x[%d]" % regs[-1]
            for k in range(len(regs) - 2, -1, -1):
                expr = "(s == %d ? x[%d] : %s)" % (k, regs[k], expr)
            return expr

        def scalar_row(indent: int) -> None:
            for ln in loop:
                st = ln.strip()
                for nm in vecs:
                    st = st.replace("%sa += " % nm, "%ss += " % nm)
                E(st, indent)

        def walk_slice(s_expr: str, own_slice: Optional[int]) -> None:
            """rows of slice s_expr; own_slice = compile-time owned slice index, None for a foreign slice (run-time)"""
            ind = 4
            E("const int sb0 = c.oi[c.slb_off + %s], npr = c.oi[c.slb_off + %s + 1] - sb0;  // first pair row, pair rows (uniform)"
              % (s_expr, s_expr), ind)
            if own_slice is None:
                E("const int4 fm = *reinterpret_cast<const int4*>(c.oi + c.fm_off + (f * 32 + lane) * 4);  // rows, owner lane, segment mask, head lane of this owner's segment", ind)
                E("const int len = fm.x, npl = len >> 1;", ind)
            else:
                E("const int len = c.oi[c.len_off + %d + lane], npl = len >> 1;" % (32 * own_slice), ind)
            for nm in vecs:
                v = vidx[nm]
                if own_slice is None:
                    E("const real %sv = __shfl_sync(pm4::FULL, x[%d], fm.y);  // the owner's slot is in slice 0" % (nm, v * KS), ind)
                else:
                    E("const real %sv = x[%d];" % (nm, v * KS + own_slice), ind)
                if pair:
                    E("const P2 %svp = pm4::pk_bcast<real>(%sv);" % (nm, nm), ind)
            E("const %s* rp = %s_rec + sb0 * 32;" % (rec_t, tag), ind)
            if pair:
                for nm in vecs:
                    E("P2 %sa = pm4::pk_bcast<real>((real)0);" % nm, ind)
                # rows beyond this lane's chunk are predicated off by a bit test on an opaque per-lane mask (a comparison
                # against the row index would be turned into a chain of divergent branches); blocks of 4, 2, 1 pair rows
                E("unsigned mk = npl >= 32 ? 0xffffffffu : ((1u << npl) - 1u);  // bit k: pair row k belongs to this lane", ind)

                def body(kk: int, i2: int) -> None:
                    E("if ((mk >> %d) & 1u) {" % kk, i2)
                    E("const %s rq = rp[%d * 32];" % (rec_t, kk), i2 + 1)
                    for j in range(W):
                        E("const P2 %s_c%d = pm4::rec_col<real, %d>(rq, %d);" % (tag, j, W, j), i2 + 1)
                    for ln in pair_lines:
                        E(ln, i2 + 1)
                    E("}", i2)

                E("int k = 0;", ind)
                E("#pragma unroll 1", ind)
                E("for (; k + 4 <= npr; k += 4, rp += 4 * 32, mk >>= 4) {  // uniform trip count", ind)
                for kk in range(4):
                    body(kk, ind + 1)
                E("}", ind)
                E("if (npr & 2) {", ind)
                for kk in range(2):
                    body(kk, ind + 1)
                E("rp += 2 * 32; mk >>= 2;", ind + 1)
                E("}", ind)
                E("if (npr & 1) {", ind)
                body(0, ind + 1)
                E("}", ind)
                for nm in vecs:
                    E("real %ss = pm4::pk_hsum(%sa);" % (nm, nm), ind)
                E("if (len & 1) {  // odd last observation: first half of the zero-padded pair", ind)
                E("const %s rq = %s_rec[(sb0 + npl) * 32];" % (rec_t, tag), ind + 2)
                for j in range(W):
                    E("const real %s_c%d = pm4::rec_col<real, %d>(rq, %d).x;" % (tag, j, W, j), ind + 2)
                E("const int i = 0; (void)i;", ind + 2)
                scalar_row(ind + 2)
                E("}", ind)
            else:
                for nm in vecs:
                    E("real %ss = 0;" % nm, ind)
                E("#pragma unroll 1", ind)
                E("for (int i2 = 0; i2 < 2 * npr; ++i2) {  // uniform trip count, rows beyond the lane's chunk predicated off", ind)
                E("if (i2 < len) {", ind + 1)
                E("const %s rq = rp[(i2 >> 1) * 32];" % rec_t, ind + 2)
                E("const int i = 0; (void)i;", ind + 2)
                for j in range(W):
                    E("const pm4::Pk<real> %s_cp%d = pm4::rec_col<real, %d>(rq, %d);" % (tag, j, W, j), ind + 2)
                    E("const real %s_c%d = (i2 & 1) ? %s_cp%d.y : %s_cp%d.x;" % (tag, j, tag, j, tag, j), ind + 2)
                scalar_row(ind + 2)
                E("}", ind + 1)
                E("}", ind)
            if own_slice is not None:
                for nm in vecs:
                    E("gr[%d] += %s;" % (vidx[nm] * KS + own_slice, sink("%ss" % nm)), ind)
            else:
                E("#pragma unroll 1", ind)
                E("for (int q = 0; q < c.Q; ++q) {  // segmented shuffle reduction over the pieces of one group ...", ind)
                for nm in vecs:
                    E("{ const real tq = __shfl_down_sync(pm4::FULL, %ss, 1u << q); if ((fm.z >> q) & 1) %ss += tq; }" % (nm, nm), ind + 2)
                E("}", ind)
                for nm in vecs:
                    E("{ const real tq = __shfl_sync(pm4::FULL, %ss, fm.w < 0 ? lane : fm.w); if (fm.w >= 0) gr[%d] += %s; }  // ... then the owner pulls the total"
                      % (nm, vidx[nm] * KS, sink("tq")), ind)

        for s_own in range(KS):
            E("{  // owned slice %d: this lane's group, parameters and gradient accumulators in its registers" % s_own, 3)
            walk_slice(str(s_own), s_own)
            E("}", 3)
        E("#pragma unroll 1", 3)
        E("for (int f = 0; f < c.NF; ++f) {  // foreign pieces of the groups capped in owned slice 0", 3)
        walk_slice("%d + f" % KS, None)
        E("}", 3)
        if pair:
            E("%s_S2 += pm4::pk_hsum(%s_S2p);" % (tag, tag), 3)
            for reg in sorted(uacc_pair):
                E("%s_ua%d += pm4::pk_hsum(%s_uap%d);" % (tag, reg, tag, reg), 3)

    # -- everything that goes into the Model struct ---------------------------------
    def members(self) -> str:
        lay = self.lay
        ids = self.id_regs()
        consts = [
            "  // owner-computes layout of the warp-per-chain path (pymc4_b200/owned.py)",
            "  static constexpr bool OWNED = true;",
            "  static constexpr int RI = %d, NV = %d, KS = %d, M0 = %d, PO = %d, OW = %d;"
            % (lay.RI, lay.NV, lay.KS, lay.M0, lay.P, max(lay.width, 1)),
            "  static constexpr unsigned ID_REGS = 0x%xu;  // registers holding identity-transformed elements only"
            % sum(1 << r for r, ok in enumerate(ids) if ok),
        ]
        return "\n".join(consts + [self.gen_constrain(), self.gen_terms()])


def not_owned_members() -> str:
    return "\n".join([
        "  static constexpr bool OWNED = false;",
        "  static constexpr int RI = 1, NV = 0, KS = 0, M0 = 0, PO = 1, OW = 1;",
        "  static constexpr unsigned ID_REGS = 0u;",
    ])
