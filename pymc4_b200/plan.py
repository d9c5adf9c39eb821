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
A chunk that is a whole group adds its sums straight into the chain's gradient; the chunks of a
split group park their partial sums in a small scratch array and are combined per group in a
fixed order afterwards (deterministic, no atomics).

The plan depends on the index *values*, so it lives in two run-time pools (``planf`` records,
``plani`` metadata); the generated code only depends on which ops are planned (``Plan.gather_ops``
/ ``record_ops`` / ``width``).  The CPU oracle ignores plans and evaluates the original tape, which
is what makes the parity tests a check of this layout.

Metadata layout (int32, all offsets relative to the start of the plan's metadata, every table
16-byte aligned):
  header[8]   n_slices, n_split_groups, n_part (scratch reals per gather), o_slices, o_slots, o_split, 0, 0
  slices      int4 per slice: rows every lane walks unmasked (min length rounded down to a multiple
              of 4), rows in the slice (max length rounded up to a multiple of 4), first record row, 0
  slots       int4 per (slice, lane): gather target (-1 = empty lane), chunk length,
              scratch index for chunks of split groups (-1 = whole group), 0
  split       int4 per split group: gather target, first scratch index, number of chunks, 0
Records (reals): row-major [row][lane][width]; row r of slice s is ``first_row + r``.
"""
from __future__ import annotations

from collections import Counter
from typing import List, Optional, Tuple

import numpy as np

from . import ir as I

HEADER = 8
ROW_UNROLL = 4


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


def _cost(chunks: List[Tuple[int, int]]) -> float:
    lens = sorted((l for _, l in chunks), reverse=True)
    n_slices = -(-len(lens) // 32)
    cost = 0.0
    for s in range(n_slices):
        sl = lens[s * 32:(s + 1) * 32]
        full = (sl[-1] if len(sl) == 32 else 0) // ROW_UNROLL * ROW_UNROLL
        rows = -(-sl[0] // ROW_UNROLL) * ROW_UNROLL
        cost += 6.5 * full + 8.5 * (rows - full) + 30.0  # unpredicated rows, masked rows, slice overhead
    n_groups_split = len({g for g, _ in chunks}) - len({g for g, c in Counter(g for g, _ in chunks).items() if c == 1})
    return cost + 25.0 * (-(-n_groups_split // 32))


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
    for cap in sorted({2, 3, 4, 6, 8, 10, 12, 14, 16, 20, 24, 32, 48, 64, int(sizes.max())}):
        ch = _chunks_for_cap(sizes, cap)
        c = _cost(ch)
        if best is None or c < best[0]:
            best = (c, cap, ch)
    _, cap, chunks = best

    # element ids of every chunk, longest chunks first (ties: by group, so a group's chunks are neighbours)
    chunk_elems: List[Tuple[int, np.ndarray]] = []
    cursor = {}
    for g, length in chunks:
        off = cursor.get(g, 0)
        chunk_elems.append((g, order[starts[g] + off: starts[g] + off + length]))
        cursor[g] = off + length
    chunk_elems.sort(key=lambda ge: (-len(ge[1]), ge[0]))
    n_chunks_of = Counter(g for g, _ in chunk_elems)

    n_slices = -(-len(chunk_elems) // 32)
    n_slots = n_slices * 32
    slices = np.zeros((n_slices, 4), dtype=np.int32)
    slots = np.zeros((n_slots, 4), dtype=np.int32)
    slots[:, 0] = -1
    slots[:, 2] = -1
    rec_rows = 0
    for s in range(n_slices):
        lens = [len(e) for _, e in chunk_elems[s * 32:(s + 1) * 32]]
        # both loops of the device walk are unrolled by ROW_UNROLL without remainder: the unmasked
        # prefix is rounded down, the slice's row count up (padding rows are zero records, masked)
        full = (lens[-1] if len(lens) == 32 else 0) // ROW_UNROLL * ROW_UNROLL
        rows = -(-lens[0] // ROW_UNROLL) * ROW_UNROLL
        slices[s] = (full, rows, rec_rows, 0)
        rec_rows += rows
    recs = np.zeros((rec_rows, 32, width), dtype=np.float64)
    cols = [np.asarray(dataf[o.blob: o.blob + n], dtype=np.float64) for o in rec_ops]

    # scratch indices: the chunks of a split group get consecutive slots
    split_groups = sorted(g for g, c in n_chunks_of.items() if c > 1)
    part_first, nxt = {}, 0
    for g in split_groups:
        part_first[g] = nxt
        nxt += n_chunks_of[g]
    n_part = nxt
    part_used = {g: 0 for g in split_groups}
    for slot, (g, elems) in enumerate(chunk_elems):
        s, lane = divmod(slot, 32)
        pidx = -1
        if g in part_first:
            pidx = part_first[g] + part_used[g]
            part_used[g] += 1
        slots[slot] = (targets[g], len(elems), pidx, 0)
        for w, col in enumerate(cols):
            recs[slices[s, 2]: slices[s, 2] + len(elems), lane, w] = col[elems]
    split = np.zeros((max(len(split_groups), 1), 4), dtype=np.int32)
    for q, g in enumerate(split_groups):
        split[q] = (targets[g], part_first[g], n_chunks_of[g], 0)

    o_slices = HEADER
    o_slots = o_slices + slices.size
    o_split = o_slots + slots.size
    header = np.array([n_slices, len(split_groups), max(n_part, 1), o_slices, o_slots, o_split, 0, 0], dtype=np.int32)
    meta = np.concatenate([header, slices.reshape(-1), slots.reshape(-1), split.reshape(-1)]).astype(np.int32)
    plan = I.Plan(term=t_idx, index_blob=int(blob), gather_ops=[o.dst for o in gathers],
                  record_ops=[o.dst for o in rec_ops], width=width, n_slices=n_slices, n_groups=G,
                  n_chunks=len(chunk_elems), chunk_cap=cap, n_part=max(n_part, 1))
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
