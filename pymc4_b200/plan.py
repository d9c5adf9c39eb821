"""Device execution plans for terms that gather from free variables (sliced-ELL over groups).

A likelihood such as ``Normal(y | a[idx] + b[idx] * x, eps)`` (the radon model,
/root/reference/notebooks/radon_hierarchical.ipynb:72-95; the reference evaluates it as
``tf.gather`` + broadcasted elementwise ops and differentiates through the gather with a
scatter-add) needs, per chain and per gradient evaluation, a gather from the chain's state for
every observation and a segment-sum of adjoints back into it.  With one warp per chain the plan
below removes both: observations are regrouped by gather target, groups are cut into chunks of at
most ``cap`` observations, chunks are sorted by length and dealt 32 at a time ("slices") to the
lanes.  Inside a slice each lane walks ONE chunk: the gathered values live in registers, adjoints
accumulate in registers, observation records are read with conflict-free shared-memory loads.
Chunk partial sums are combined per group in a fixed order (deterministic, no atomics).

The plan depends on the index *values*, so it lives in two run-time pools (``planf`` records,
``plani`` metadata); the generated code only depends on which ops are planned (``Plan.gather_ops``
/ ``record_ops`` / ``width``).  The CPU oracle ignores plans and evaluates the original tape, which
is what makes the parity tests a check of this layout.
"""
from __future__ import annotations

from collections import Counter
from typing import List, Optional, Tuple

import numpy as np

from . import ir as I

HEADER = 12  # ints: n_slices, n_groups, n_slots, n_chunks, o_maxlen, o_recbase, o_ctarget, o_clen,
#                   o_gtarget, o_gptr, o_gslots, total_len


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


def _cost(chunks: List[Tuple[int, int]]) -> Tuple[float, int, int]:
    lens = sorted((l for _, l in chunks), reverse=True)
    n_slices = -(-len(lens) // 32)
    iters = sum(lens[s * 32] for s in range(n_slices))
    # ~7 issue slots per walked element row, ~14 per slice (prologue/epilogue), ~2 per chunk in the group pass
    return 7.0 * iters + 14.0 * n_slices + 2.0 * len(lens) / 32.0, n_slices, iters


def plan_term(t_idx: int, term: I.Term, datai: np.ndarray, dataf: np.ndarray) -> Optional[Tuple[I.Plan, np.ndarray, np.ndarray]]:
    ld_x, ld_data = I.OP["ld_x"], I.OP["ld_data"]
    gathers = [o for o in term.ops if o.opcode == ld_x and o.mode == I.MODE_GATHER]
    if not gathers or term.n < 64:
        return None
    if any(o.opcode in (ld_x, I.OP["ld_z"]) and o.mode == I.MODE_ELEM for o in term.ops):
        return None  # element-wise reads of a free variable would have to be permuted too
    blob = Counter(o.blob for o in gathers).most_common(1)[0][0]
    if any(o.blob != blob for o in gathers):
        return None  # a second index array: keep the generic path
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
    for cap in sorted({2, 3, 4, 6, 8, 10, 12, 16, 20, 24, 32, 48, 64, int(sizes.max())}):
        ch = _chunks_for_cap(sizes, cap)
        c, n_slices, iters = _cost(ch)
        if best is None or c < best[0]:
            best = (c, cap, ch, n_slices, iters)
    _, cap, chunks, n_slices, _ = best

    # element ids of every chunk
    chunk_elems: List[Tuple[int, np.ndarray]] = []
    cursor = {}
    for g, length in chunks:
        off = cursor.get(g, 0)
        chunk_elems.append((g, order[starts[g] + off: starts[g] + off + length]))
        cursor[g] = off + length
    chunk_elems.sort(key=lambda ge: -len(ge[1]))  # stable: ties keep group order

    n_slots = n_slices * 32
    slice_maxlen = np.zeros(n_slices, dtype=np.int32)
    slice_recbase = np.zeros(n_slices, dtype=np.int32)
    ctarget = np.full(n_slots, -1, dtype=np.int32)
    clen = np.zeros(n_slots, dtype=np.int32)
    rec_rows = 0
    for s in range(n_slices):
        slice_maxlen[s] = len(chunk_elems[s * 32][1])
        slice_recbase[s] = rec_rows
        rec_rows += int(slice_maxlen[s])
    recs = np.zeros((rec_rows, 32, width), dtype=np.float64)
    cols = [np.asarray(dataf[o.blob: o.blob + n], dtype=np.float64) for o in rec_ops]
    group_slots: List[List[int]] = [[] for _ in range(G)]
    for slot, (g, elems) in enumerate(chunk_elems):
        s, lane = divmod(slot, 32)
        ctarget[slot] = targets[g]
        clen[slot] = len(elems)
        group_slots[g].append(slot)
        for w, col in enumerate(cols):
            recs[slice_recbase[s]: slice_recbase[s] + len(elems), lane, w] = col[elems]
    gptr = np.zeros(G + 1, dtype=np.int32)
    gslots: List[int] = []
    for g in range(G):
        gslots.extend(sorted(group_slots[g]))
        gptr[g + 1] = len(gslots)

    body = [slice_maxlen, slice_recbase, ctarget, clen, targets.astype(np.int32), gptr,
            np.asarray(gslots, dtype=np.int32)]
    offs, cur = [], HEADER
    for arr in body:
        offs.append(cur)
        cur += len(arr)
    header = np.array([n_slices, G, n_slots, len(chunk_elems)] + offs + [cur], dtype=np.int32)
    meta = np.concatenate([header] + body).astype(np.int32)
    plan = I.Plan(term=t_idx, index_blob=int(blob), gather_ops=[o.dst for o in gathers],
                  record_ops=[o.dst for o in rec_ops], width=width, n_slices=n_slices, n_groups=G,
                  n_chunks=len(chunk_elems), chunk_cap=cap)
    return plan, meta, recs.reshape(-1)


def attach_plans(ir: I.ModelIR) -> I.ModelIR:
    """Plan every eligible term of ``ir`` in place and fill the plan pools."""
    fparts, iparts, flen, ilen = [], [], 0, 0
    for t_idx, term in enumerate(ir.terms):
        term.plan = None
        res = plan_term(t_idx, term, ir.datai, ir.dataf)
        if res is None:
            continue
        plan, meta, recs = res
        fpad, ipad = (-flen) % 4, (-ilen) % 4
        if fpad:
            fparts.append(np.zeros(fpad)); flen += fpad
        if ipad:
            iparts.append(np.zeros(ipad, dtype=np.int32)); ilen += ipad
        plan.meta_off, plan.recs_off = ilen, flen
        fparts.append(recs); flen += len(recs)
        iparts.append(meta); ilen += len(meta)
        term.plan = plan
    # pools are copied with 16-byte bulk transfers: pad to a multiple of 4 elements
    if flen % 4:
        fparts.append(np.zeros(4 - flen % 4))
    if ilen % 4:
        iparts.append(np.zeros(4 - ilen % 4, dtype=np.int32))
    ir.planf = np.concatenate(fparts) if fparts else np.zeros(0)
    ir.plani = np.concatenate(iparts).astype(np.int32) if iparts else np.zeros(0, dtype=np.int32)
    return ir
