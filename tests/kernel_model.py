"""Executable model (plain Python, sequential loops) of the ORDER-INDEPENDENT CLOSED FORM the CUDA kernels
implement, phase by phase.  TEST INFRASTRUCTURE: it exists so the algorithm can be
property-tested against the sequential oracle on a CPU-only box before (and independently of) the device code.
It models the MATH of each phase (explicit per-record arrays, plain pointer jumping); the device code computes the same
quantities with bitmaps + popcount prefixes, splitter list ranking, segment trees and a hash table.

Phases (same letters as peritext_b200/csrc/merge_kernel.cuh):
  A  id table        T[K(ctr,actor)] -> insert record index,   K = (ctr-1)*R + actor
  B  parents/deletes p[i] = T[K(ref)], childCount, deleted[]
  C  runs            log-contiguous only-child chains
  D  run tree        parentRun, sibling order by descending K
  E  euler ranking   weighted list ranking of enter/exit nodes -> run start offsets
  F  sequence        pos[i]; tokens in sequence order; visible ranks
  G  marks           element-index intervals, boundary bitmap -> segments, stabbing-max per LWW type
  H  comments        per-id elementary pieces -> visible presence pieces -> head flags
  I  spans           head flags -> spans, comment lists, digest
form="visible" swaps G-I for the warp kernel's cut (peritext_b200/csrc/warp_kernel.cuh): intervals in visible space, ops that
cover no visible element dropped, marks evaluated per visible position (`spans_visible_form`).
"""
from __future__ import annotations

import numpy as np

from peritext_b200.packing import (ATTR_NONE, RESULT_DT, SPAN_DT, MergedBatch, PackedBatch, output_layout)

MASK64 = (1 << 64) - 1


def _mix64(z):
    z &= MASK64
    z ^= z >> 30; z = (z * 0xBF58476D1CE4E5B9) & MASK64
    z ^= z >> 27; z = (z * 0x94D049BB133111EB) & MASK64
    z ^= z >> 31
    return z


def term_text(i, tok): return _mix64(((i << 32) | tok) + 0x9E3779B97F4A7C15)
def term_span(j, start, flags, link): return _mix64(_mix64(((j << 32) | start) ^ 0xA5A5A5A55A5A5A5A) + ((flags << 32) | link))
def term_comment(j, k, cid): return _mix64(_mix64(((j << 32) | k) ^ 0x5BD1E9955BD1E995) + cid)
def term_counts(nv, ns): return _mix64(((nv << 32) | ns) ^ 0xC3C3C3C33C3C3C3C)
def term_hi(t): return ((t << 23) | (t >> 41)) & MASK64


EMPTY = -1
ROOT = -2


def spans_visible_form(mk, iv, vis_rank, nvis, mark_keys):
    """Phases G-I the way the WARP kernel computes them (peritext_b200/csrc/warp_kernel.cuh): mark intervals are mapped to
    VISIBLE space, ops that cover no visible element are dropped, and marks / link / comment-id set are evaluated per
    visible position; a span starts where any of them differs from the position before.  (The kernel's segment form
    evaluates the same thing once per elementary segment of the survivors' boundaries.)"""
    surv = []                                  # (va, vb, op index)
    for k, x in enumerate(iv):
        if x:
            va, vb = vis_rank[x[0]], vis_rank[x[1]]
            if va < vb:
                surv.append((va, vb, k))
    per_pos = []
    for x in range(nvis):
        cover = [k for (va, vb, k) in surv if va <= x < vb]
        flags, link = 0, ATTR_NONE
        for t, bit in ((0, 1), (1, 2), (3, 4)):
            ops = [k for k in cover if ((int(mk[k]["kind"]) >> 1) & 3) == t]
            if ops:
                w = max(ops, key=lambda k: mark_keys[k])            # LWW by opId (peritext.ts:304-313)
                if (int(mk[w]["kind"]) & 1) == 0:
                    flags |= bit
                    if t == 3:
                        link = int(mk[w]["attr"])
        cops = [k for k in cover if ((int(mk[k]["kind"]) >> 1) & 3) == 2]
        if cops:
            flags |= 8                                              # quirk Q3: the key exists as soon as any comment op covers
        ids = set()
        for cid in set(int(mk[k]["attr"]) for k in cops):
            last = max(k for k in cops if int(mk[k]["attr"]) == cid)    # arrival order (quirk Q4)
            if (int(mk[last]["kind"]) & 1) == 0:
                ids.add(cid)
        per_pos.append((flags, link, tuple(sorted(ids))))
    spans = []
    for x in range(nvis):
        if x == 0 or per_pos[x] != per_pos[x - 1]:
            fl, ln, cl = per_pos[x]
            spans.append((x, fl | (len(cl) << 8), ln, cl))
    return spans


def merge_log(ins, mk, n_actors, max_ctr, form="element"):
    """Returns dict(status, n_elems, n_visible, tokens, spans=[(start, flags, link, (comment ids...))], digest).
    form: "element" = phases G-I in element space as the CTA kernel does them; "visible" = as the warp kernel does them."""
    n, m, R = len(ins), len(mk), int(n_actors)
    C = int(max_ctr)
    out = dict(status=0, n_elems=0, n_visible=0, tokens=[], spans=[], digest=(0, 0))

    def K(ctr, actor):
        return (int(ctr) - 1) * R + int(actor)

    def bad_id(ctr, actor):
        return not (1 <= int(ctr) <= C and int(actor) < R)

    kind = [int(r["payload"]) >> 30 for r in ins]
    tok = [int(r["payload"]) & 0x3FFFFFFF for r in ins]
    for k in kind:
        if k > 1:
            out["status"] = 3; return out

    # ---- A: id table -------------------------------------------------------------------------------------------
    T = [EMPTY] * (C * R)
    for i, r in enumerate(ins):
        if bad_id(r["ctr"], r["actor"]):
            out["status"] = 2; return out
        if kind[i] == 0:
            k = K(r["ctr"], r["actor"])
            if T[k] != EMPTY:
                out["status"] = 2; return out
            T[k] = i

    def lookup(ctr, actor):
        if bad_id(ctr, actor):
            return EMPTY
        return T[K(ctr, actor)]

    # ---- B: parents, child counts, deletes -------------------------------------------------------------------------
    p = [EMPTY] * n
    child_count = [0] * (n + 1)          # slot n = HEAD
    deleted = [False] * n
    for i, r in enumerate(ins):
        if kind[i] == 0:
            if int(r["ref_ctr"]) == 0:
                p[i] = ROOT; child_count[n] += 1
            else:
                j = lookup(r["ref_ctr"], r["ref_actor"])
                if j == EMPTY:
                    out["status"] = 1; return out
                if K(r["ref_ctr"], r["ref_actor"]) >= K(r["ctr"], r["actor"]):
                    out["status"] = 5; return out
                p[i] = j; child_count[j] += 1
        else:
            j = lookup(r["ref_ctr"], r["ref_actor"]) if int(r["ref_ctr"]) else EMPTY
            if j == EMPTY:
                out["status"] = 1; return out
            deleted[j] = True

    # ---- C: runs -------------------------------------------------------------------------------------------------
    head = [False] * n
    for i in range(n):
        if kind[i] != 0:
            continue
        cont = i > 0 and kind[i - 1] == 0 and p[i] == i - 1 and child_count[i - 1] == 1
        head[i] = not cont
    run_id = [EMPTY] * n
    run_head, run_len = [], []
    for i in range(n):
        if kind[i] != 0:
            continue
        if head[i]:
            run_head.append(i); run_len.append(0)
        run_id[i] = len(run_head) - 1
        run_len[-1] += 1
    M = len(run_head)
    N = sum(run_len)

    # ---- D: run tree, sibling order (descending K of the run head) -------------------------------------------------
    prun = [M if p[h] == ROOT else run_id[p[h]] for h in run_head]       # node M = HEAD
    key = [K(ins[h]["ctr"], ins[h]["actor"]) for h in run_head]
    groups = [[] for _ in range(M + 1)]
    for r in range(M):
        groups[prun[r]].append(r)
    first_child = [EMPTY] * (M + 1)
    next_sib = [EMPTY] * M
    for q in range(M + 1):
        g = groups[q]
        if not g:
            continue
        ranked = [None] * len(g)
        for r in g:
            rank = sum(1 for s in g if key[s] > key[r])
            ranked[rank] = r
        first_child[q] = ranked[0]
        for a, b in zip(ranked, ranked[1:]):
            next_sib[a] = b

    # ---- E: Euler tour + weighted list ranking ------------------------------------------------------------------------
    # nodes: enter(r) = r, exit(r) = (M+1) + r, for r in 0..M (M = HEAD); END = 2(M+1)
    E = 2 * (M + 1)
    END = E
    nxt = [END] * (E + 1)
    w = [0] * (E + 1)
    for r in range(M + 1):
        ent, ext = r, (M + 1) + r
        nxt[ent] = first_child[r] if first_child[r] != EMPTY else ext
        w[ent] = run_len[r] if r < M else 0
        if r == M:
            nxt[ext] = END
        else:
            nxt[ext] = next_sib[r] if next_sib[r] != EMPTY else (M + 1) + prun[r]
    d = list(w)
    rounds = 0
    while (1 << rounds) < E + 1:
        nd = [d[x] + (d[nxt[x]] if nxt[x] != END else 0) if x != END else 0 for x in range(E + 1)]
        nn = [nxt[nxt[x]] if x != END and nxt[x] != END else END for x in range(E + 1)]
        d, nxt = nd, nn
        rounds += 1
    total = d[M]
    assert total == N
    run_start = [total - d[r] for r in range(M)]

    # ---- F: sequence positions, tokens, visible ranks -----------------------------------------------------------------
    pos = [EMPTY] * n
    seq_tok = [0] * N
    seq_del = [False] * N
    for i in range(n):
        if kind[i] == 0:
            r = run_id[i]
            pos[i] = run_start[r] + (i - run_head[r])
            seq_tok[pos[i]] = tok[i]
            seq_del[pos[i]] = deleted[i]
    vis_rank = [0] * (N + 1)                     # exclusive prefix count of visible elements
    for x in range(N):
        vis_rank[x + 1] = vis_rank[x] + (0 if seq_del[x] else 1)
    nvis = vis_rank[N]
    tokens = [seq_tok[x] for x in range(N) if not seq_del[x]]
    out["n_elems"], out["n_visible"], out["tokens"] = N, nvis, tokens

    # ---- G: marks -> element-index intervals, segments, stabbing max --------------------------------------------------
    # mark rank by K (bitmap over K space + prefix popcount in the kernel)
    mark_keys = []
    for k, r in enumerate(mk):
        if bad_id(r["ctr"], r["actor"]):
            out["status"] = 2; return out
        mark_keys.append(K(r["ctr"], r["actor"]))
    order = sorted(range(m), key=lambda k: mark_keys[k])
    if len(set(mark_keys)) != m:
        out["status"] = 2; return out
    mrank = [0] * m
    for rk, k in enumerate(order):
        mrank[k] = rk

    def slot(btype, ctr, actor, is_start, arrival):
        if btype == 2:                      # startOfText: never matches a slot (peritext.ts:236)
            return None if is_start else 2 * N
        if btype == 3:                      # endOfText: never matches (runs to the end)
            return None if is_start else 2 * N
        j = lookup(ctr, actor)
        # unknown element, or one that is inserted LATER in this log: the walk at apply time never matches it (no throw
        # in the reference, peritext.ts:236-241): a missing start is a no-op, a missing end never ends
        if j == EMPTY or j >= int(arrival):
            return None if is_start else 2 * N
        return 2 * pos[j] + (1 if btype == 1 else 0)

    iv = []                                   # per mark op: (a, b) element interval or None
    for k, r in enumerate(mk):
        sb, eb = int(r["bounds"]) & 3, (int(r["bounds"]) >> 2) & 3
        ps = slot(sb, r["start_ctr"], r["start_actor"], True, r["arrival"])
        pe = slot(eb, r["end_ctr"], r["end_actor"], False, r["arrival"])
        if ps is None:
            iv.append(None); continue
        if pe == ps:                         # same slot: the start branch wins, op never ends (Q2)
            pe = 2 * N
        a, b = (ps + 1) >> 1, min((pe + 1) >> 1, N)
        iv.append((a, b) if a < b else None)

    bits = [0] * (N + 1)
    for x in iv:
        if x:
            bits[x[0]] = 1; bits[x[1]] = 1
    seg = [0] * (N + 1)                       # seg(x) = popcount(bits[0..x])
    acc = 0
    for x in range(N + 1):
        acc += bits[x]; seg[x] = acc
    S = acc + 1                               # segment ids 0..S-1 (+ one past for b == N)
    win = {t: [0] * (S + 1) for t in (0, 1, 3)}      # stabbing max of (mrank+1) per LWW type
    ccount = [0] * (S + 2)
    for k, r in enumerate(mk):
        if not iv[k]:
            continue
        t = (int(r["kind"]) >> 1) & 3
        lo, hi = seg[iv[k][0]], seg[iv[k][1]]
        if t == 2:
            ccount[lo] += 1; ccount[hi] -= 1
        else:
            for s in range(lo, hi):
                win[t][s] = max(win[t][s], mrank[k] + 1)
    seg_flags = [0] * (S + 1)
    seg_link = [ATTR_NONE] * (S + 1)
    acc = 0
    for s in range(S + 1):
        acc += ccount[s]
        f = 0
        for t, bit in ((0, 1), (1, 2), (3, 4)):
            wv = win[t][s]
            if wv:
                op = mk[order[wv - 1]]
                if (int(op["kind"]) & 1) == 0:
                    f |= bit
                    if t == 3:
                        seg_link[s] = int(op["attr"])
        if acc > 0:
            f |= 8
        seg_flags[s] = f

    # ---- H: comments -> presence pieces in visible space -> head flags ---------------------------------------------------
    cops = [k for k in range(m) if ((int(mk[k]["kind"]) >> 1) & 3) == 2 and iv[k]]
    pieces = []                                # (id, va, vb)
    for k in cops:
        cid = int(mk[k]["attr"])
        same = [j for j in cops if int(mk[j]["attr"]) == cid]
        for which in (0, 1):
            e = iv[k][which]
            # dedupe: among the group's endpoints with this value, only the first (op index, side) emits
            if any(iv[j][w] == e and (j, w) < (k, which) for j in same for w in (0, 1)):
                continue
            ends = [x for j in same for x in iv[j] if x > e]
            if not ends:
                continue
            e2 = min(ends)
            cover = [j for j in same if iv[j][0] <= e and e2 <= iv[j][1]]
            if not cover:
                continue
            # opsToMarks folds comment ops in Set order = ARRIVAL order (peritext.ts:314-322, no opId comparison): the
            # last-arrived covering op of this id decides (quirk Q4); mark records are stored in arrival order
            wj = max(cover)
            if (int(mk[wj]["kind"]) & 1) == 0:
                va, vb = vis_rank[e], vis_rank[e2]
                if va < vb:
                    pieces.append((cid, va, vb))
    chead = [False] * (nvis + 1)
    for (cid, va, vb) in pieces:
        if not any(c2 == cid and b2 == va for (c2, a2, b2) in pieces):
            chead[va] = True
        if not any(c2 == cid and a2 == vb for (c2, a2, b2) in pieces):
            chead[vb] = True

    # ---- I: spans ---------------------------------------------------------------------------------------------------------
    vis_seg = [seg[x] for x in range(N) if not seq_del[x]]
    heads = []
    for v in range(nvis):
        if v == 0:
            h = True
        else:
            s1, s2 = vis_seg[v - 1], vis_seg[v]
            h = chead[v] or (s1 != s2 and (seg_flags[s1] != seg_flags[s2] or seg_link[s1] != seg_link[s2]))
        if h:
            heads.append(v)
    spans = []
    for j, v in enumerate(heads):
        s = vis_seg[v]
        cl = sorted(cid for (cid, va, vb) in pieces if va <= v < vb)
        flags = seg_flags[s] | (len(cl) << 8)
        spans.append((v, flags, seg_link[s], tuple(cl)))
    if form == "visible":
        spans = spans_visible_form(mk, iv, vis_rank, nvis, mark_keys)
    d0 = d1 = 0

    def add(t):
        nonlocal d0, d1
        d0 = (d0 + t) & MASK64; d1 = d1 ^ term_hi(t)
    for i, t in enumerate(tokens):
        add(term_text(i, t))
    for j, (st, fl, ln, cl) in enumerate(spans):
        for k2, c in enumerate(cl):
            add(term_comment(j, k2, c))
        add(term_span(j, st, fl, ln))
    add(term_counts(nvis, len(spans)))
    out["spans"] = spans
    out["digest"] = (d0, d1)
    return out


def merge_batch(batch: PackedBatch, form="element") -> MergedBatch:
    desc = batch.desc
    text_off, span_off, n_text, n_span = output_layout(desc)
    results = np.zeros(len(desc), RESULT_DT)
    text = np.zeros(max(n_text, 1), np.uint32)
    spans = np.zeros(max(n_span, 1), SPAN_DT)
    pool = []
    for i in range(len(desc)):
        ins, mk = batch.log_slice(i)
        o = merge_log(ins, mk, desc[i]["n_actors"], desc[i]["max_ctr"], form=form)
        results[i]["status"] = o["status"]
        if o["status"]:
            continue
        results[i]["n_elems"] = o["n_elems"]; results[i]["n_visible"] = o["n_visible"]; results[i]["n_spans"] = len(o["spans"])
        results[i]["digest"] = o["digest"]
        to, so = int(text_off[i]), int(span_off[i])
        text[to: to + o["n_visible"]] = o["tokens"]
        for j, (st, fl, ln, cl) in enumerate(o["spans"]):
            spans[so + j] = (st, fl, ln, len(pool) if cl else 0)
            pool.extend(cl)
    return MergedBatch(results, text_off, span_off, text, spans, np.array(pool, np.uint32))
