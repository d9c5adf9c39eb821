"""Owner-computes execution plan of the warp-per-chain path (csrc/pm4_warp.cuh, ``Model::OWNED``).

With ONE warp per chain (``ModelIR.team_warps == 1``) the chain's state does not have to sit in
theta order.  The likelihood of a hierarchical model, ``Normal(y | a[idx] + b[idx] * x, eps)``
(/root/reference/notebooks/radon_hierarchical.ipynb:72-95; the reference evaluates it as ``tf.gather``
+ broadcasted element-wise ops and differentiates through the gather with a scatter-add), is then
laid out so that *the lane that walks a group's observations is the lane that holds the group's
parameters*: no gather, no scatter, no hand-over through shared memory.

Internal layout of a D-vector, ``RI`` registers per lane (``perm[reg * 32 + lane]`` = element or -1):

* registers ``[v * KS, (v + 1) * KS)``: the v-th gathered vector (a, b, ...), ``KS = ceil(G / 32)``.
  Groups are sorted by their number of observations; *owned slice* ``s`` holds ranks ``32 s .. 32 s + 31``,
  rank ``r`` on lane ``r % 32``.  In owned slice ``s`` every lane walks the observations of ITS group with the
  group's parameters in registers, and the adjoint sums land in the gradient registers directly.
* the biggest groups (all in slice 0) would make their lanes the critical path, so slice 0 is capped
  at ``L0`` rows; the rest of such a group is cut into *foreign pieces* walked by other lanes in
  ``NF`` extra slices.  A foreign lane gets the group's parameters with one shuffle per vector from the
  owner lane; the pieces of one group sit on neighbouring lanes of a foreign slice, so their sums are
  combined by a segmented shuffle reduction of ``Q`` steps and the owner pulls the total from the
  segment's first lane.
* *misc* registers ``[NV * KS, RI)``: everything else.  Scalar parameter k (an element some term reads
  as a chain-uniform scalar) sits on lane ``per * k`` of the first misc register (``per = 32 / P``), which is
  exactly where the packed butterfly reduction of the scalar adjoints delivers total k.

Records are stored as PAIRS of observations (one 16-byte load feeds two packed-FP32 lanes:
``fma.rn.f32x2``), ``[pair row][lane][column][2]``; an odd last row is zero padded and handled by a
scalar tail.  Everything data dependent lives in the plan pools (appended to ``planf`` / ``plani``);
the generated code depends only on the structure (NV, KS, the misc layout).

The CPU oracle ignores all of this and evaluates the original tape in theta order, which is what makes
every parity test a test of this layout.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import ir as I

MAX_RI = 10          # registers per lane and vector the warp path accepts
HDR_INTS = 8         # owned header: enabled, i_off, i_len, f_off, f_len, NF, Q (segmented-reduction steps), n


@dataclass
class OwnedLayout:
    """Structure-level part of the owned plan (what the generated code is specialised on)."""
    term: int                       # index of the planned term (-1: the model has none)
    NV: int                         # gathered vectors
    KS: int                         # owned slices = registers per gathered vector
    G: int                          # elements per gathered vector
    vec_bases: List[int]            # theta offset of every gathered vector, in register order
    width: int                      # record columns (1 or 2)
    P: int                          # packed reduction width (power of two > number of scalars)
    RM: int                         # misc registers
    misc_pos: Dict[int, int] = field(default_factory=dict)  # element -> misc position (reg * 32 + lane, reg relative to M0)

    @property
    def M0(self) -> int:
        return self.NV * self.KS

    @property
    def RI(self) -> int:
        return self.M0 + self.RM

    @property
    def per(self) -> int:
        return 32 // self.P

    def key(self) -> bytes:
        items = sorted(self.misc_pos.items())
        return ("owned:%d,%d,%d,%d,%s,%d,%d,%d|%s" % (
            self.term, self.NV, self.KS, self.G, self.vec_bases, self.width, self.P, self.RM, items)).encode()


def _term_shape(t: "I.Term") -> str:
    ld_x = I.OP["ld_x"]
    if all(o.uniform for o in t.ops) and t.n == 1:
        return "uniform"
    if t.plan is not None:
        return "planned"
    if len({o.base for o in t.ops if o.opcode == ld_x and o.mode == I.MODE_ELEM}) == 1:
        return "aligned"
    return "general"


def owned_layout(ir: "I.ModelIR") -> Optional[OwnedLayout]:
    """The structure-level layout, or None when the model does not qualify (it then runs on the
    xs / part hand-over path of plan.py)."""
    if ir.team_warps != 1 or ir.lockstep:
        return None
    ld_x, ld_z, ld_data = I.OP["ld_x"], I.OP["ld_z"], I.OP["ld_data"]
    planned = [k for k, t in enumerate(ir.terms) if t.plan is not None]
    if len(planned) > 1:
        return None
    for t in ir.terms:
        if _term_shape(t) == "general" or any(o.opcode == ld_z for o in t.ops):
            return None
    scalars = list(ir.scalar_pos)
    P = 1
    while P < len(scalars) + 1:
        P *= 2
    if P > 32:
        return None
    vec_bases: List[int] = []
    KS = G = width = 0
    term = -1
    owned_elems: set = set()
    if planned:
        term = planned[0]
        t = ir.terms[term]
        pl = t.plan
        by_dst = {o.dst: o for o in t.ops}
        for dst in pl.gather_ops:
            b = by_dst[dst].base
            if b not in vec_bases:
                vec_bases.append(b)
        sizes = []
        for b in vec_bases:
            rv = [r for r in ir.rvs if r.offset == b]
            if not rv:
                return None  # a gather into the middle of a variable
            sizes.append(rv[0].size)
        if len(set(sizes)) != 1:
            return None
        G = sizes[0]
        idx = np.asarray(ir.datai[pl.index_blob: pl.index_blob + t.n])
        if idx.size and (idx.min() < 0 or idx.max() >= G):
            return None
        width = len(pl.record_ops)
        if width < 1 or width > 2:
            return None
        KS = -(-G // 32)
        for b in vec_bases:
            owned_elems.update(range(b, b + G))
        if owned_elems & set(scalars):
            return None
    NV = len(vec_bases)
    per = 32 // P
    misc_pos: Dict[int, int] = {}
    reserved = {per * k for k in range(len(scalars))}
    for k, pos in enumerate(scalars):
        misc_pos[pos] = per * k
    q = 0
    for j in range(ir.D):
        if j in owned_elems or j in misc_pos:
            continue
        while q in reserved:
            q += 1
        misc_pos[j] = q
        q += 1
    top = max(list(misc_pos.values()) + [0])
    RM = top // 32 + 1
    lay = OwnedLayout(term=term, NV=NV, KS=KS, G=G, vec_bases=vec_bases, width=width, P=P, RM=RM, misc_pos=misc_pos)
    if lay.RI > MAX_RI:
        return None
    return lay


def _pieces(left: int, cap: int) -> List[int]:
    if left <= 0:
        return []
    k = -(-left // cap)
    base, extra = divmod(left, k)
    return [base + (1 if c < extra else 0) for c in range(k)]


def _choose_caps(sizes_desc: np.ndarray) -> Tuple[int, int]:
    """(L0, Lf): row cap of owned slice 0 and of a foreign piece, minimising the rows every lane walks
    (plus a per-slice and per-step overhead, in row units of ~3 instructions)."""
    top = np.asarray(sizes_desc[:32], dtype=np.int64)
    rest_cost = sum(int(sizes_desc[s]) for s in range(32, len(sizes_desc), 32))
    best = None
    hi = int(top.max(initial=0))
    cands = sorted({int(v) for v in top if v > 0} | {max(1, hi // k) for k in range(1, 17)})
    for L0 in cands:
        left = top[top > L0] - L0
        if left.size == 0:
            cost = L0 + rest_cost
            if best is None or cost < best[0]:
                best = (cost, L0, max(L0, 1))
            continue
        for Lf in range(2, 49):
            k = -(-left // Lf)                       # pieces per group, near-equal lengths
            longest = -(-left // k)
            n_pieces = int(k.sum())
            nf = -(-n_pieces // 32)
            # the pieces are dealt longest-first, 32 per foreign slice: slice f is as long as its longest piece
            lens = np.sort(np.repeat(longest, k))[::-1]
            f_rows = int(lens[::32].sum())
            steps = int(np.ceil(np.log2(max(int(k.max()), 1)))) if k.max() > 1 else 0
            cost = L0 + rest_cost + f_rows + 7.0 * nf + 2.0 * steps * nf + 0.0 * int(longest.max())
            if best is None or cost < best[0]:
                best = (cost, L0, Lf)
    return best[1], best[2]


def attach_owned(ir: "I.ModelIR") -> Optional[OwnedLayout]:
    """Compute the owned plan of ``ir`` (after plan.attach_plans), append its pools to planf / plani and write
    the owned header behind the per-plan headers.  Sets ``ir.owned`` (None when the model does not qualify)."""
    lay = owned_layout(ir)
    ir.owned = lay
    hdr_at = ir.phdr_off + 4 * max(len(ir.plans), 1)
    plani = list(np.asarray(ir.plani, dtype=np.int32))
    assert len(plani) == hdr_at, (len(plani), hdr_at)
    hdr = [0] * HDR_INTS
    if lay is None:
        ir.plani = np.asarray(plani + hdr, dtype=np.int32)
        return None
    RI, KS, NV, M0 = lay.RI, lay.KS, lay.NV, lay.M0
    perm = -np.ones(RI * 32, dtype=np.int64)
    for j, q in lay.misc_pos.items():
        perm[M0 * 32 + q] = j
    own_len = np.zeros(max(KS, 1) * 32, dtype=np.int64)
    sl_base: List[int] = []     # first pair row of every slice (owned, then foreign)
    fmeta: List[List[int]] = []  # per foreign slice, per lane: (len, owner lane)
    rec_rows: List[np.ndarray] = []
    n_term = 0
    NF = Q = 0
    W = max(lay.width, 1)
    if lay.term >= 0:
        t = ir.terms[lay.term]
        pl = t.plan
        n_term = t.n
        by_dst = {o.dst: o for o in t.ops}
        cols = [np.asarray(ir.dataf[by_dst[d].blob: by_dst[d].blob + t.n], dtype=np.float64) for d in pl.record_ops]
        idx = np.asarray(ir.datai[pl.index_blob: pl.index_blob + t.n], dtype=np.int64)
        G = lay.G
        sizes = np.bincount(idx, minlength=G)
        order = np.argsort(idx, kind="stable")
        starts = np.concatenate([[0], np.cumsum(sizes)])
        rank = np.argsort(-sizes, kind="stable")  # groups by size, largest first (ties: by group index)
        sizes_desc = sizes[rank]
        L0, Lf = _choose_caps(sizes_desc)

        def pair_rows(chunks: List[np.ndarray]) -> np.ndarray:
            """[pair rows, 32, W, 2] records of one slice; chunk k belongs to lane k."""
            rows = max([len(c) for c in chunks] + [0])
            pr = -(-rows // 2)
            out = np.zeros((pr, 32, W, 2), dtype=np.float64)
            for lane, el in enumerate(chunks):
                for w, col in enumerate(cols):
                    v = col[el]
                    full = len(v) // 2
                    out[:full, lane, w, 0] = v[0:2 * full:2]
                    out[:full, lane, w, 1] = v[1:2 * full:2]
                    if len(v) & 1:
                        out[full, lane, w, 0] = v[-1]
            return out

        base_row = 0
        leftovers: List[Tuple[int, np.ndarray]] = []  # (owner lane, element ids)
        for s in range(KS):
            chunks = []
            for lane in range(32):
                r = 32 * s + lane
                if r >= G:
                    chunks.append(np.zeros(0, dtype=np.int64))
                    continue
                g = int(rank[r])
                for v, b in enumerate(lay.vec_bases):
                    perm[(v * KS + s) * 32 + lane] = b + g
                el = order[starts[g]: starts[g + 1]]
                if s == 0 and len(el) > L0:
                    leftovers.append((lane, el[L0:]))
                    el = el[:L0]
                own_len[s * 32 + lane] = len(el)
                chunks.append(el)
            recs = pair_rows(chunks)
            sl_base.append(base_row)
            base_row += recs.shape[0]
            rec_rows.append(recs)
        pieces: List[Tuple[int, np.ndarray]] = []
        for lane, el in leftovers:
            off = 0
            for ln in _pieces(len(el), Lf):
                pieces.append((lane, el[off: off + ln]))
                off += ln
        pieces.sort(key=lambda p: (-len(p[1]), p[0]))
        NF = -(-len(pieces) // 32)
        for f in range(NF):
            # within a foreign slice the pieces of one group sit on neighbouring lanes: their sums are added up by
            # a segmented shuffle reduction (log2 steps), then the owner pulls the segment head's total
            sl = sorted(pieces[f * 32: (f + 1) * 32], key=lambda p: p[0])
            chunks = [p[1] for p in sl] + [np.zeros(0, dtype=np.int64)] * (32 - len(sl))
            meta = [[0, 0, 0, -1] for _ in range(32)]  # rows, owner lane, segment mask, (owner lanes) head lane of its segment
            lane = 0
            while lane < len(sl):
                owner = sl[lane][0]
                m = 1
                while lane + m < len(sl) and sl[lane + m][0] == owner:
                    m += 1
                steps = 0
                while (1 << steps) < m:
                    steps += 1
                Q = max(Q, steps)
                for k in range(m):
                    mask = 0
                    for st in range(steps):
                        if k + (1 << st) < m:
                            mask |= 1 << st
                    meta[lane + k][0] = len(sl[lane + k][1])
                    meta[lane + k][1] = owner
                    meta[lane + k][2] = mask
                meta[owner][3] = lane
                lane += m
            fmeta.append(meta)
            recs = pair_rows(chunks)
            sl_base.append(base_row)
            base_row += recs.shape[0]
            rec_rows.append(recs)
        sl_base.append(base_row)  # sentinel: pair rows of slice s = sl_base[s + 1] - sl_base[s]

    # ---- int section: perm | own_len | sl_base + sentinel (padded to 4) | foreign meta int4 per lane
    sec: List[int] = list(perm) + list(own_len)
    if not sl_base:
        sl_base = [0]
    sec += sl_base + [0] * ((-len(sl_base)) % 4)
    for meta in fmeta:
        for row in meta:
            sec += row
    sec += [0] * ((-len(sec)) % 4)
    planf = np.asarray(ir.planf, dtype=np.float64)
    fpad = (-planf.size) % 4
    frec = np.concatenate([r.reshape(-1) for r in rec_rows]) if rec_rows else np.zeros(0)
    frec = np.concatenate([frec, np.zeros((-frec.size) % 4)])
    f_off = planf.size + fpad
    i_off = hdr_at + HDR_INTS
    hdr = [1, i_off, len(sec), f_off, int(frec.size), NF, Q, n_term]
    ir.plani = np.asarray(plani + hdr + sec, dtype=np.int32)
    ir.planf = np.concatenate([planf, np.zeros(fpad), frec])
    ir.owned_stats = dict(L0=int(L0) if lay.term >= 0 else 0, Lf=int(Lf) if lay.term >= 0 else 0, NF=NF, Q=Q,
                          pair_rows=[int(r.shape[0]) for r in rec_rows])
    return lay
