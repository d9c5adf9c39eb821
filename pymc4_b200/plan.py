"""Device execution plans: how one chain's log-density pass is spread over a TEAM of warps.

A chain is owned by a team of ``T`` warps (``TL = 32 T`` lanes; ``ModelIR.team_warps``).  Every
D-vector of the chain is spread over the team, lane ``tl`` holding elements ``tl, tl + TL, ...``.

Gather terms.  A likelihood such as ``Normal(y | a[idx] + b[idx] * x, eps)`` (the radon model,
/root/reference/notebooks/radon_hierarchical.ipynb:72-95; the reference evaluates it as
``tf.gather`` + broadcasted elementwise ops and differentiates through the gather with a
scatter-add) needs, per chain and per gradient evaluation, a gather from the chain's state for
every observation and a segment-sum of adjoints back into it.  The plan removes both:
observations are regrouped by gather target, groups are cut into chunks of at most ``cap``
observations, chunks are sorted by length and dealt ``TL`` at a time ("slices") to the lanes of
the team (sliced ELL).  Inside a slice each lane walks ONE chunk: the gathered values live in
registers, adjoints accumulate in registers, observation records are read with conflict-free
shared-memory loads.  Every chunk then parks its adjoint sums in the chain's ``part`` array.

The ``part`` array is the single hand-over point between the lanes that *produce* adjoints and the
lane that *owns* a gradient element.  Element ``j`` owns the K fixed entries ``part[K j .. K j + K)`` (K = 4;
1 for very large states), read with ONE (vector) load after one team barrier; its contributions (chunk sums of every gather op
that targets it, then -- for parameters read as scalars -- one warp-level partial sum per warp of
the team) fill them in order, entries nobody writes stay zero (the array is zeroed once per
launch).  An element with more than K contributions (a group cut into many chunks) continues in an
overflow block of ``ocnt_j`` 4-entry vectors starting at ``ofirst_j`` that its owner adds in a
loop: ``elem_tab[j] = ofirst_j << 12 | ocnt_j`` (0 for almost every element).  No atomics, no second
pass, a fixed summation order (same seed => same trace).  ``spart[k * T + w]`` is the entry warp w
parks the k-th scalar parameter's adjoint at (``ModelIR.scalar_pos`` order); row ``n_scalar`` is the
log-density (T consecutive entries, 16-byte aligned).

The plan depends on the index *values*, so it lives in two run-time pools (``planf`` records,
``plani`` metadata); the generated code only depends on which ops are planned and on ``T``.  The CPU
oracle ignores plans and evaluates the original tape, which is what makes the parity tests a check
of this layout.

Per-plan metadata (int32, 16-byte aligned):
  slices      int4 per slice: shortest chunk (0 if a lane is empty), longest chunk = record rows of
              the slice, first record row, 0
  slots       SW int4 per (slice, lane), right after the slices: gather target (-1 = empty lane),
              chunk length, then the ``part`` index of the chunk for every gather op of the plan
Records (reals): row-major [row][lane][width]; row r of slice s is ``first_row + r``.
Model tables at the end of the metadata pool: ``elem_tab[DP]``, ``spart[(n_scalar + 1) T]``, and one
int4 header per plan ``(n_slices, plani offset of the slices, planf offset of the records, n)``.
"""
from __future__ import annotations

import os
from collections import Counter
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import ir as I

ROW_UNROLL = 4
MAX_GATHER_OPS = 6
ELEM_CNT_BITS = 12
TEAM_SIZES = (1, 2, 4, 8, 16, 32)


def choose_team_warps(ir: "I.ModelIR") -> int:
    """Warps per chain: enough lanes that a lane walks only a handful of observations and holds a
    couple of state elements, as few as possible beyond that (every extra warp repeats the scalar
    sampler logic).  PM4_TEAM_WARPS overrides (experiments)."""
    env = os.environ.get("PM4_TEAM_WARPS")
    if env:
        t = int(env)
        if t not in TEAM_SIZES:
            raise ValueError("PM4_TEAM_WARPS must be one of %s" % (TEAM_SIZES,))
        return t
    work = sum(t.n for t in ir.terms if t.glm is None)  # dense GLM terms are not evaluated by the teams
    # small models: ONE warp per chain (csrc/pm4_warp.cuh) -- the state fits in 8 registers per lane and vector,
    # reductions are warp shuffles, and nothing of the sampler's scalar logic is repeated by a second warp
    if ir.D <= 32 * 8 and work <= 4096:
        return 1
    # ~640 term elements per warp; at most ~20 state elements per lane (theta, momentum and gradient
    # stay in registers); never more than 16 warps (512 threads keep 128 registers per thread)
    want = max(work / 640.0, ir.D / (32.0 * 20.0))
    for t in TEAM_SIZES:
        if t >= want:
            return min(t, 16)
    return 16 if ir.D <= 32 * 16 * 24 else 32


def _chunks_for_cap(sizes: np.ndarray, cap: int) -> List[Tuple[int, int]]:
    """[(group, length)] with every group cut into ceil(size/cap) near-equal pieces."""
    out = []
    for g, s in enumerate(sizes):
        if s == 0:
            continue
        k = -(-int(s) // cap)
        base, extra = divmod(int(s), k)
        out.extend((g, base + (1 if c < extra else 0)) for c in range(k))
    return out


def _cost(chunks: List[Tuple[int, int]], TL: int) -> float:
    """Warp-instructions per warp of one evaluation: every warp pays for the longest chunk of each
    slice, for the slice bookkeeping, and (the unluckiest owner lane) for overflow vectors."""
    lens = sorted((l for _, l in chunks), reverse=True)
    n_slices = -(-len(lens) // TL)
    cost = 0.0
    for s in range(n_slices):
        cost += 7.0 * lens[s * TL] + 10.0 + 30.0
    per_group = Counter(g for g, _ in chunks)
    over = max(0, -(-(max(per_group.values()) - 4) // 4))
    return cost + 8.0 * over


class _TermPlan:
    """Chunking of one term before ``part`` indices are known."""

    def __init__(self, plan: I.Plan, chunk_elems, targets, cols, gather_bases):
        self.plan = plan
        self.chunk_elems = chunk_elems      # [(group ordinal, element ids)] in slot order
        self.targets = targets              # group ordinal -> gather target value
        self.cols = cols                    # record columns (float64 arrays over the term's elements)
        self.gather_bases = gather_bases    # theta offset of every gather op


def _chunk_term(t_idx: int, term: I.Term, datai: np.ndarray, dataf: np.ndarray, TL: int) -> Optional[_TermPlan]:
    ld_x, ld_data = I.OP["ld_x"], I.OP["ld_data"]
    gathers = [o for o in term.ops if o.opcode == ld_x and o.mode == I.MODE_GATHER]
    if not gathers or term.n < 64:
        return None
    if any(o.opcode in (ld_x, I.OP["ld_z"]) and o.mode == I.MODE_ELEM for o in term.ops):
        return None  # element-wise reads of a free variable would have to be permuted too
    blob = Counter(o.blob for o in gathers).most_common(1)[0][0]
    if any(o.blob != blob for o in gathers):
        return None  # a second index array: keep the generic path
    if len(gathers) > MAX_GATHER_OPS:
        return None
    rec_ops = [o for o in term.ops if o.opcode == ld_data and o.mode == I.MODE_ELEM]
    if len(rec_ops) > 4:
        return None
    width = 1 if len(rec_ops) <= 1 else (2 if len(rec_ops) == 2 else 4)

    n = term.n
    idx = np.asarray(datai[blob: blob + n], dtype=np.int64)
    targets, inv = np.unique(idx, return_inverse=True)
    G = len(targets)
    sizes = np.bincount(inv, minlength=G)
    order = np.argsort(inv, kind="stable")  # element ids grouped by target, original order inside
    starts = np.concatenate([[0], np.cumsum(sizes)])

    best = None
    max_chunks_per_group = (1 << ELEM_CNT_BITS) - 64
    caps = sorted({2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 16, 18, 20, 24, 28, 32, 40, 48, 64, 96, 128,
                   int(sizes.max())})
    for cap in caps:
        if -(-int(sizes.max()) // cap) > max_chunks_per_group:
            continue
        ch = _chunks_for_cap(sizes, cap)
        c = _cost(ch, TL)
        if best is None or c < best[0]:
            best = (c, cap, ch)
    _, cap, chunks = best

    # element ids of every chunk, longest chunks first (ties: by group, so a group's chunks are neighbours)
    chunk_elems: List[Tuple[int, np.ndarray]] = []
    cursor: Dict[int, int] = {}
    for g, length in chunks:
        off = cursor.get(g, 0)
        chunk_elems.append((g, order[starts[g] + off: starts[g] + off + length]))
        cursor[g] = off + length
    chunk_elems.sort(key=lambda ge: (-len(ge[1]), ge[0]))
    cols = [np.asarray(dataf[o.blob: o.blob + n], dtype=np.float64) for o in rec_ops]
    plan = I.Plan(term=t_idx, index_blob=int(blob), gather_ops=[o.dst for o in gathers],
                  record_ops=[o.dst for o in rec_ops], width=width, n_slices=-(-len(chunk_elems) // TL),
                  n_groups=G, n_chunks=len(chunk_elems), chunk_cap=cap)
    return _TermPlan(plan, chunk_elems, targets, cols, [o.base for o in gathers])


def scalar_positions(ir: "I.ModelIR") -> List[int]:
    """theta positions read as a chain-uniform scalar by any term, in first-use order."""
    out: List[int] = []
    for t in ir.terms:
        for o in t.ops:
            if o.opcode in (I.OP["ld_x"], I.OP["ld_z"]) and o.mode == I.MODE_SCALAR and o.base not in out:
                out.append(o.base)
    return out


def attach_plans(ir: "I.ModelIR", team_warps: Optional[int] = None) -> "I.ModelIR":
    """Plan every eligible term of ``ir`` in place, lay out the ``part`` array and fill the pools."""
    for term in ir.terms:
        term.plan = None
    T = int(team_warps) if team_warps else choose_team_warps(ir)
    if T not in TEAM_SIZES:
        raise ValueError("team_warps must be one of %s" % (TEAM_SIZES,))
    TL = 32 * T
    ir.team_warps = T
    R = max(1, -(-ir.D // TL))
    DPt = R * TL
    ir.scalar_pos = scalar_positions(ir)

    tplans: List[_TermPlan] = []
    for t_idx, term in enumerate(ir.terms):
        tp = _chunk_term(t_idx, term, ir.datai, ir.dataf, TL)
        if tp is not None:
            tplans.append(tp)
            term.plan = tp.plan

    # ---- layout of the part array: K fixed entries per element, then overflow blocks, then lp ----
    K = 4 if DPt <= 4096 else 1  # large states cannot afford four shared-memory entries per element
    ir.part_k = K
    n_chunk_contrib = np.zeros(ir.D, dtype=np.int64)
    for tp in tplans:
        per_group = Counter(g for g, _ in tp.chunk_elems)
        for base in tp.gather_bases:
            for g, c in per_group.items():
                n_chunk_contrib[base + int(tp.targets[g])] += c
    cnt = n_chunk_contrib.copy()
    for pos in ir.scalar_pos:
        cnt[pos] += T
    ocnt = -(-np.maximum(cnt - K, 0) // 4)  # overflow vectors of 4 entries
    if ocnt.max(initial=0) >= (1 << ELEM_CNT_BITS):
        raise NotImplementedError("an element of theta receives more than %d partial sums" % ((1 << ELEM_CNT_BITS) - 1))
    ofirst = K * DPt + 4 * np.concatenate([[0], np.cumsum(ocnt)])[:-1]
    lp_first = K * DPt + 4 * int(ocnt.sum())  # read with one vector load: 16-byte aligned
    part_reals = lp_first + T
    if part_reals >= (1 << (31 - ELEM_CNT_BITS)):
        raise NotImplementedError("part array too large for the packed element table")
    elem_tab = np.zeros(DPt, dtype=np.int32)
    elem_tab[: ir.D] = np.where(ocnt > 0, (ofirst << ELEM_CNT_BITS) | ocnt, 0)

    def loc(j: int, o: int) -> int:  # part entry of the o-th contribution to element j
        return K * j + o if o < K else int(ofirst[j]) + (o - K)

    spart = []
    for pos in ir.scalar_pos:
        spart.extend(loc(pos, int(n_chunk_contrib[pos]) + w) for w in range(T))
    spart.extend(lp_first + w for w in range(T))
    used = np.zeros(ir.D, dtype=np.int64)  # chunk entries handed out so far, per element

    fparts, iparts, flen, ilen = [], [], 0, 0
    for tp in tplans:
        pl = tp.plan
        n_g = len(tp.gather_bases)
        SW = -(-(2 + n_g) // 4)
        n_slices = pl.n_slices
        slices = np.zeros((n_slices, 4), dtype=np.int32)
        slots = np.zeros((n_slices * TL, 4 * SW), dtype=np.int32)
        slots[:, 0] = -1
        rec_rows = 0
        for s in range(n_slices):
            lens = [len(e) for _, e in tp.chunk_elems[s * TL:(s + 1) * TL]]
            # every lane walks exactly its own chunk (per-lane trip count, the hardware masks the lanes
            # that are done): a slice needs as many record rows as its longest chunk
            full = lens[-1] if len(lens) == TL else 0
            rows = lens[0]
            slices[s] = (full, rows, rec_rows, 0)
            rec_rows += rows
        recs = np.zeros((rec_rows, TL, pl.width), dtype=np.float64)
        for slot, (g, elems) in enumerate(tp.chunk_elems):
            s, lane = divmod(slot, TL)
            tgt = int(tp.targets[g])
            row = [tgt, len(elems)]
            for base in tp.gather_bases:
                j = base + tgt
                row.append(loc(j, int(used[j])))
                used[j] += 1
            slots[slot, : len(row)] = row
            for w, col in enumerate(tp.cols):
                recs[slices[s, 2]: slices[s, 2] + len(elems), lane, w] = col[elems]
        meta = np.concatenate([slices.reshape(-1), slots.reshape(-1)]).astype(np.int32)
        fpad, ipad = (-flen) % 4, (-ilen) % 4
        if fpad:
            fparts.append(np.zeros(fpad)); flen += fpad
        if ipad:
            iparts.append(np.zeros(ipad, dtype=np.int32)); ilen += ipad
        pl.meta_off, pl.recs_off = ilen, flen
        fparts.append(recs.reshape(-1)); flen += recs.size
        iparts.append(meta); ilen += len(meta)
    assert np.array_equal(used, n_chunk_contrib)

    # model-level tables at the end of the metadata pool
    ipad = (-ilen) % 4
    if ipad:
        iparts.append(np.zeros(ipad, dtype=np.int32)); ilen += ipad
    ir.elem_off = ilen
    iparts.append(elem_tab); ilen += len(elem_tab)
    ir.spart_off = ilen
    sp = np.zeros(-(-len(spart) // 4) * 4, dtype=np.int32)
    sp[: len(spart)] = spart
    iparts.append(sp); ilen += len(sp)
    ir.phdr_off = ilen
    hdr = np.zeros((max(len(tplans), 1), 4), dtype=np.int32)
    for k, tp in enumerate(tplans):
        hdr[k] = (tp.plan.n_slices, tp.plan.meta_off, tp.plan.recs_off, ir.terms[tp.plan.term].n)
    iparts.append(hdr.reshape(-1)); ilen += hdr.size
    ir.part_reals = int(-(-part_reals // 4) * 4)
    # pools are copied with 16-byte bulk transfers: pad to a multiple of 4 elements
    if flen % 4:
        fparts.append(np.zeros(4 - flen % 4))
    ir.planf = np.concatenate(fparts) if fparts else np.zeros(0)
    ir.plani = np.concatenate(iparts).astype(np.int32)
    return ir
