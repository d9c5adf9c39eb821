"""Model IR -> specialised CUDA module (nvcc, sm_100a).

The hand-written device code lives in ``csrc/`` (``pm4_dists.cuh``: log-densities with
analytic partials; ``pm4_kernels.cuh``: logp/grad, HMC and NUTS kernels; ``pm4_philox.cuh``:
RNG).  This module only emits the *glue*: a ``Model`` struct whose ``terms`` member is the
model's tape unrolled into straight-line code with the reverse sweep written out, so that
nvcc inlines the density functions into the persistent leapfrog / NUTS kernel and keeps
the chain state in registers.  One shared object is built per IR *structure*
(``ModelIR.struct_hash``); data values, element counts and pool offsets stay run-time
arguments, so new observed data reuses the same module.

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
_build_lock = threading.Lock()

_UNARY_FWD = {
    "neg": "-{a}", "exp": "pm4::m_exp({a})", "log": "pm4::m_log({a})", "log1p": "pm4::m_log1p({a})",
    "sqrt": "pm4::m_sqrt({a})", "square": "{a} * {a}", "sigmoid": "pm4::m_sigmoid<real>({a})",
    "softplus": "pm4::m_softplus<real>({a})", "tanh": "pm4::m_tanh({a})", "abs": "pm4::m_abs({a})",
    "recip": "pm4::m_rcp({a})", "lgamma": "pm4::m_lgamma({a})",
}
# d(out)/d(a) as an expression of a (input) and r (output)
_UNARY_BWD = {
    "neg": "-{g}", "exp": "{g} * {r}", "log": "{g} / {a}", "log1p": "{g} / ((real)1 + {a})",
    "sqrt": "{g} / ((real)2 * {r})", "square": "(real)2 * {g} * {a}", "sigmoid": "{g} * {r} * ((real)1 - {r})",
    "softplus": "{g} * pm4::m_sigmoid<real>({a})", "tanh": "{g} * ((real)1 - {r} * {r})",
    "abs": "(({a} > 0) ? {g} : (({a} < 0) ? -{g} : (real)0))", "recip": "-{g} * {r} * {r}",
    "lgamma": "{g} * pm4::m_digamma<real>({a})",
}


def _lit(v: float) -> str:
    if v != v:
        return "(real)NAN"
    if v in (float("inf"), float("-inf")):
        return "(-pm4::m_inf<real>())" if v < 0 else "pm4::m_inf<real>()"
    return "(real)%r" % float(v)


class _TermGen:
    """Emit forward + reverse code for one tape."""

    def __init__(self, gen: "_ModuleGen", ops: List[I.Op], op_index0: int, tag: str):
        self.gen = gen
        self.ops = ops
        self.op0 = op_index0
        self.tag = tag

    def v(self, reg: int) -> str:
        return "%s_v%d" % (self.tag, reg)

    # forward expression of one op; `idx` is the element index expression, `xsrc` how x is read
    def fwd(self, o: I.Op, idx: str, xread) -> str:
        name = I.OP_NAME[o.opcode]
        if name == "ld_const":
            return _lit(o.imm)
        if name == "ld_data":
            blob = "c_blob(%d)" % (self.op0 + o.dst)
            return "A.dataf[%s + %s]" % (blob, idx if o.mode == I.MODE_ELEM else "0")
        if name in ("ld_x", "ld_z"):
            if name == "ld_z":
                raise NotImplementedError("unconstrained-space loads (ld_z) are not generated yet")
            return xread(o, idx)
        a = self.v(o.a)
        if name in _UNARY_FWD:
            return _UNARY_FWD[name].format(a=a)
        b = self.v(o.b)
        return {"add": "%s + %s", "sub": "%s - %s", "mul": "%s * %s", "div": "%s / %s",
                "pow": "pm4::m_pow(%s, %s)"}[name] % (a, b)

    # contributions of adjoint g of op o to its inputs: list of (input reg, expr)
    def bwd(self, o: I.Op, g: str) -> List[Tuple[int, str]]:
        name = I.OP_NAME[o.opcode]
        r = self.v(o.dst)
        if name in _UNARY_BWD:
            return [(o.a, _UNARY_BWD[name].format(g=g, a=self.v(o.a), r=r))]
        a, b = (self.v(o.a), self.v(o.b)) if o.b >= 0 else (None, None)
        if name == "add":
            return [(o.a, g), (o.b, g)]
        if name == "sub":
            return [(o.a, g), (o.b, "-" + g)]
        if name == "mul":
            return [(o.a, "%s * %s" % (g, b)), (o.b, "%s * %s" % (g, a))]
        if name == "div":
            return [(o.a, "%s / %s" % (g, b)), (o.b, "-%s * %s / %s" % (g, r, b))]
        if name == "pow":
            return [(o.a, "%s * %s * pm4::m_pow(%s, %s - (real)1)" % (g, b, a, b)),
                    (o.b, "((%s > 0) ? %s * %s * pm4::m_log(%s) : (real)0)" % (a, g, r, a))]
        return []


class _ModuleGen:
    def __init__(self, ir: I.ModelIR, has_f64: bool):
        self.ir = ir
        self.has_f64 = has_f64
        self.D = ir.D
        self.R = max(1, (ir.D + 31) // 32)
        self.lines: List[str] = []
        self.scalar_pos: List[int] = []  # distinct scalar x positions used anywhere

    def emit(self, s: str = "", ind: int = 2):
        self.lines.append("  " * ind + s)

    # -- constrain -----------------------------------------------------------
    def gen_constrain(self) -> str:
        out = ["  template <typename real>",
               "  static __device__ __forceinline__ void constrain(int r, int lane, real z, real& x, real& dxdz) {",
               "    const int j = lane + 32 * r; (void)j;",
               "    x = z; dxdz = (real)1;"]
        for rv in self.ir.rvs:
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
            self.scalar_pos.append(pos)
        return "sacc_%d" % pos

    def gen_terms(self) -> str:
        body: List[str] = []
        self.lines = body
        op0 = 0
        for k, t in enumerate(self.ir.terms):
            self.gen_one_term(k, t, op0)
            op0 += len(t.ops)
        head = ["  template <typename real>",
                "  static __device__ __forceinline__ real terms(const real (&x)[R], real (&gr)[R], pm4::Ctx<real>& c) {",
                "    const int lane = c.lane; (void)lane;",
                "    const pm4::ModelArgs<real>& A = c.A; (void)A;",
                "    auto c_blob = [&](int op) -> int { return __ldg(A.op_blob + op); }; (void)c_blob;",
                "    real lp = 0;"]
        for pos in self.scalar_pos:
            head.append("    const real sx_%d = c.xs[%d]; real sacc_%d = 0;" % (pos, pos, pos))
        tail = []
        for pos in self.scalar_pos:
            tail.append("    { const real tot = pm4::warp_sum(sacc_%d); if (lane == %d) gr[%d] += tot; }"
                        % (pos, pos & 31, pos >> 5))
        tail += ["    return lp;", "  }"]
        return "\n".join(head + body + tail)

    def gen_one_term(self, k: int, t: I.Term, op0: int):
        ops = t.ops
        tag = "t%d" % k
        tg = _TermGen(self, ops, op0, tag)
        kind = t.kind
        uniform = [o.uniform for o in ops]
        elem_bases = sorted({o.base for o in ops
                             if I.OP_NAME[o.opcode] == "ld_x" and o.mode == I.MODE_ELEM})
        all_uniform = all(uniform)
        aligned = (len(elem_bases) == 1)
        E = self.emit
        E("// ---- term %d: %s (%s), n=%s" % (k, t.label, [n for n, c in I.TERM_KINDS.items() if c == kind][0], t.n))
        E("{")
        touched_gx = [False]

        def xread(o: I.Op, idx: str) -> str:
            if o.mode == I.MODE_SCALAR:
                self._sacc(o.base)
                return "sx_%d" % o.base
            if o.mode == I.MODE_ELEM:
                if aligned:
                    return "x[r]"
                return "c.xs[%d + %s]" % (o.base, idx)
            return "c.xs[%d + A.datai[c_blob(%d) + %s]]" % (o.base, op0 + o.dst, idx)

        # adjoint bookkeeping -------------------------------------------------
        # uniform registers accumulate lane-partial adjoints across the element loop
        def is_u(reg: int) -> bool:
            return uniform[reg]

        # 1. prologue: uniform forward ops
        for o in ops:
            if o.uniform:
                E("const real %s = %s;" % (tg.v(o.dst), tg.fwd(o, "0", xread)), 3)
        u_need_adj = [False] * len(ops)

        # registers needing adjoints: reachable from value/params
        roots = [t.value_reg] + [p for p in t.param_regs if p >= 0]

        # lpdf partial wiring
        n_par = I.TERM_NPARAMS[kind]
        preg = list(t.param_regs[:n_par])
        scale_reg = None
        if kind in (I.TERM_KINDS["normal"],):
            scale_reg = preg[1]
        elif kind in (I.TERM_KINDS["halfnormal"], I.TERM_KINDS["halfcauchy"]):
            scale_reg = preg[0]
        scale_uniform = scale_reg is not None and is_u(scale_reg)
        if scale_uniform:
            E("const real %s_inv_s = pm4::m_rcp(%s); const real %s_log_s = pm4::m_log(%s);"
              % (tag, tg.v(scale_reg), tag, tg.v(scale_reg)), 3)
        # accumulators for uniform registers (declared lazily)
        uacc_declared: Dict[int, bool] = {}

        def uacc(reg: int) -> str:
            if reg not in uacc_declared:
                uacc_declared[reg] = True
            return "%s_ua%d" % (tag, reg)

        # we need to know which uniform regs receive adjoints before emitting the loop, so
        # generate the loop body into a buffer first
        saved = self.lines
        loop: List[str] = []
        self.lines = loop

        idx = "i"
        for o in ops:
            if not o.uniform:
                E("const real %s = %s;" % (tg.v(o.dst), tg.fwd(o, idx, xread)), 4)
        # lpdf
        val = tg.v(t.value_reg)
        adj_terms: Dict[int, List[str]] = {}

        def add_adj(reg: int, expr: str):
            adj_terms.setdefault(reg, []).append(expr)

        if kind == I.TERM_KINDS["potential"]:
            E("lp += %s;" % val, 4)
            add_adj(t.value_reg, "(real)1")
        else:
            E("real %s_dx, %s_d0, %s_d1; (void)%s_dx; (void)%s_d0; (void)%s_d1;" % ((tag,) * 6), 4)
            if scale_reg is not None and not scale_uniform:
                E("const real %s_inv_s = pm4::m_rcp(%s); const real %s_log_s = pm4::m_log(%s);"
                  % (tag, tg.v(scale_reg), tag, tg.v(scale_reg)), 4)
            if kind == I.TERM_KINDS["normal"]:
                E("lp += pm4::normal_lpdf<real>(%s, %s, %s_inv_s, %s_log_s, %s_dx, %s_d0, %s_d1);"
                  % (val, tg.v(preg[0]), tag, tag, tag, tag, tag), 4)
                add_adj(preg[0], "%s_d0" % tag)
                add_adj(preg[1], "%s_d1" % tag)
            elif kind == I.TERM_KINDS["halfnormal"]:
                E("lp += pm4::halfnormal_lpdf<real>(%s, %s_inv_s, %s_log_s, %s_dx, %s_d0);"
                  % (val, tag, tag, tag, tag), 4)
                add_adj(preg[0], "%s_d0" % tag)
            elif kind == I.TERM_KINDS["halfcauchy"]:
                E("lp += pm4::halfcauchy_lpdf<real>(%s, %s_inv_s, %s_log_s, %s_dx, %s_d0);"
                  % (val, tag, tag, tag, tag), 4)
                add_adj(preg[0], "%s_d0" % tag)
            elif kind == I.TERM_KINDS["bernoulli"]:
                E("lp += pm4::bernoulli_lpdf<real>(%s, %s, %s_dx, %s_d0);" % (val, tg.v(preg[0]), tag, tag), 4)
                add_adj(preg[0], "%s_d0" % tag)
            elif kind == I.TERM_KINDS["poisson"]:
                E("lp += pm4::poisson_lpdf<real>(%s, %s, %s_dx, %s_d0);" % (val, tg.v(preg[0]), tag, tag), 4)
                add_adj(preg[0], "%s_d0" % tag)
            else:  # pragma: no cover
                raise NotImplementedError(kind)
            add_adj(t.value_reg, "%s_dx" % tag)

        # which registers depend on a free variable at all (others need no adjoint)
        depends = [False] * len(ops)
        for o in ops:
            nm = I.OP_NAME[o.opcode]
            if nm in ("ld_x", "ld_z"):
                depends[o.dst] = True
            elif nm not in ("ld_const", "ld_data"):
                depends[o.dst] = depends[o.a] or (o.b >= 0 and depends[o.b])

        def leaf_adj(o: I.Op, g: str, ind: int):
            """adjoint g arrives at a load of x."""
            if o.mode == I.MODE_SCALAR:
                E("%s += %s;" % (self._sacc(o.base), g), ind)
            elif o.mode == I.MODE_ELEM:
                if aligned:
                    E("gr[r] += %s;" % g, ind)
                else:
                    E("c.gx[%d + %s] += %s;" % (o.base, idx, g), ind)
                    touched_gx[0] = True
            else:
                E("atomicAdd(&c.gx[%d + A.datai[c_blob(%d) + %s]], %s);" % (o.base, op0 + o.dst, idx, g), ind)
                touched_gx[0] = True

        # 2. reverse sweep over non-uniform ops (inside the loop); adjoints that reach a
        #    uniform register are accumulated for the epilogue
        for o in reversed(ops):
            if o.dst not in adj_terms or not depends[o.dst]:
                continue
            terms_ = adj_terms.pop(o.dst)
            if o.uniform:
                E("%s += %s;" % (uacc(o.dst), " + ".join(terms_)), 4)
                u_need_adj[o.dst] = True
                continue
            g = "%s_a%d" % (tag, o.dst)
            E("const real %s = %s;" % (g, " + ".join(terms_)), 4)
            nm = I.OP_NAME[o.opcode]
            if nm == "ld_x":
                leaf_adj(o, g, 4)
            else:
                for reg, expr in tg.bwd(o, g):
                    if reg >= 0 and depends[reg]:
                        add_adj(reg, expr)
        self.lines = saved

        # declare uniform accumulators
        for reg in sorted(uacc_declared):
            E("real %s_ua%d = 0;" % (tag, reg), 3)

        # 3. emit the loop in the right shape
        if all_uniform and t.n == 1:
            E("if (lane == 0) {", 3)
            E("const int i = 0; (void)i;", 4)
            body_ind = loop
            self.lines.extend(body_ind)
            E("}", 3)
        elif aligned:
            base = elem_bases[0]
            r_lo, r_hi = base // 32, (base + t.n - 1) // 32
            E("#pragma unroll", 3)
            E("for (int r = %d; r <= %d; ++r) {" % (r_lo, r_hi), 3)
            E("const int i = lane + 32 * r - %d;" % base, 4)
            E("if (i >= 0 && i < %d) {" % t.n, 4)
            self.lines.extend(loop)
            E("}", 4)
            E("}", 3)
        else:
            E("const int n_%s = __ldg(A.term_n + %d);" % (tag, k), 3)
            E("for (int i = lane; i < n_%s; i += 32) {" % tag, 3)
            self.lines.extend(loop)
            E("}", 3)

        # 4. epilogue: reverse sweep over the uniform ops with lane-partial adjoints
        u_adj: Dict[int, List[str]] = {reg: ["%s_ua%d" % (tag, reg)] for reg in uacc_declared}
        for o in reversed(ops):
            if not o.uniform or o.dst not in u_adj or not depends[o.dst]:
                continue
            terms_ = u_adj.pop(o.dst)
            g = "%s_ub%d" % (tag, o.dst)
            E("const real %s = %s;" % (g, " + ".join(terms_)), 3)
            nm = I.OP_NAME[o.opcode]
            if nm == "ld_x":
                E("%s += %s;" % (self._sacc(o.base), g), 3)
            else:
                for reg, expr in tg.bwd(o, g):
                    if reg >= 0 and depends[reg]:
                        u_adj.setdefault(reg, []).append(expr)
        if touched_gx[0]:
            E("__syncwarp();", 3)
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
            tg = _TermGen(self, d.ops, op0, "d%d" % k)

            def xread(o: I.Op, idx: str, _op0=op0) -> str:
                if o.mode == I.MODE_SCALAR:
                    return "xv[%d]" % o.base
                if o.mode == I.MODE_ELEM:
                    return "xv[%d + %s]" % (o.base, idx)
                return "xv[%d + A.datai[c_blob(%d) + %s]]" % (o.base, _op0 + o.dst, idx)

            out.append("    // ---- output %d: %s" % (k, d.name))
            out.append("    for (int i = 0; i < %d; ++i) {" % d.n)
            for o in d.ops:
                out.append("      const real %s = %s;" % (tg.v(o.dst), tg.fwd(o, "i", xread)))
            out.append("      out[%d + i] = %s;" % (d.out_offset, tg.v(d.value_reg)))
            out.append("    }")
            op0 += len(d.ops)
        out.append("  }")
        return "\n".join(out)

    def source(self) -> str:
        ir = self.ir
        name = "Model_%s" % ir.struct_hash_hex
        n_ops = sum(len(t.ops) for t in ir.terms) + sum(len(d.ops) for d in ir.dets)
        terms_src = self.gen_terms()
        parts = [
            "// generated by pymc4_b200/codegen.py -- do not edit.  IR structure hash %s" % ir.struct_hash_hex,
            "// free variables: " + ", ".join("%s[%d]@%d" % (rv.name, rv.size, rv.offset) for rv in ir.rvs),
            '#include "pm4_dists.cuh"',
            '#include "pm4_philox.cuh"',
            '#include "pm4_module.h"',
            "#include <cuda_runtime.h>",
            "namespace pm4 { template <typename real> struct Ctx; template <typename real> struct ModelArgs;",
            "template <typename real> __device__ __forceinline__ real warp_sum(real v); }",
            "#define PM4_MODEL %s" % name,
            "#define PM4_HAS_F64 %d" % (1 if self.has_f64 else 0),
            "struct %s {" % name,
            "  static constexpr int D = %d, R = %d, DP = %d, DET_ROW = %d, N_TERMS = %d, N_OPS = %d;"
            % (self.D, self.R, 32 * self.R, ir.det_row, len(ir.terms), n_ops),
            "  static constexpr unsigned long long HASH = 0x%sull;" % ir.struct_hash_hex,
            self.gen_constrain(),
            terms_src,
            self.gen_dets(),
            "};",
            '#include "pm4_kernels.cuh"',
            "",
        ]
        return "\n".join(parts)


def generate_source(ir: I.ModelIR, has_f64: bool = False) -> str:
    return _ModuleGen(ir, has_f64).source()


def module_path(ir: I.ModelIR, has_f64: bool = False) -> str:
    return os.path.join(CACHE_DIR, "pm4m_%s%s.so" % (ir.struct_hash_hex, "_f64" if has_f64 else ""))


def find_nvcc() -> str:
    for cand in (os.environ.get("PM4_NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the specialised sampler module cannot be built")


def _csrc_mtime() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
               if f.endswith((".cuh", ".h")))


def build_module(ir: I.ModelIR, has_f64: bool = False, force: bool = False, verbose: bool = False) -> str:
    """Return the path of the compiled module for this IR structure, building it if needed."""
    path = module_path(ir, has_f64)
    # a module with double precision also serves float requests
    alt = module_path(ir, True)
    with _build_lock:
        for cand in (path, alt):
            if not force and os.path.exists(cand) and os.path.getmtime(cand) >= _csrc_mtime():
                return cand
        os.makedirs(CACHE_DIR, exist_ok=True)
        src_path = os.path.join(CACHE_DIR, "pm4m_%s%s.cu" % (ir.struct_hash_hex, "_f64" if has_f64 else ""))
        with open(src_path, "w") as f:
            f.write(generate_source(ir, has_f64))
        tmp = tempfile.NamedTemporaryFile(prefix="pm4m_", suffix=".so", dir=CACHE_DIR, delete=False).name
        cmd = [find_nvcc()] + NVCC_FLAGS + ["-I", CSRC, "-shared", "-o", tmp, src_path]
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
