"""Order-independent closed forms of the reference's Patch stream, host side.

The reference derives a `Patch[]` while it applies each op sequentially (reference src/micromerge.ts:659-671 insert,
:689-703 delete, :589-592 makeList; src/peritext.ts:175-220, 251-281 marks).  All of them turn out to be functions of
(a) the FINAL position of every element in the RGA sequence — which the engine materialises on the GPU — and (b) the
ARRIVAL TIME of every op at this replica:

* insert / delete index  = #{ e : pos(e) < pos(x), t_ins(e) < t, not (t_del(e) < t) }         (a dominance count)
  (relative order of elements never changes, so final positions order every past state); a delete emits a patch only
  if it is the element's first delete.
* insert `marks`         = opsToMarks({ mark op : t_op < t, p(start) <= 2*pos(y)+1 < p(end) }), y = the last element
  left of x that was present at time t (`getActiveMarksAtIndex`, src/peritext.ts:328-330, 405-436).
* mark patches: the walk of a mark op rewrites its start slot, every slot already DEFINED (start/end of an earlier op)
  inside its range, and its end slot; between two consecutive rewritten slots one patch is open iff adding the op changes
  the marks of the set at the left slot; it covers the visible indices of the text as it was at time t.

`tests/test_patch_closed_form.py` checks these against the oracle's patch stream on seeded fuzz sessions.  This module is
what the `Micromerge` facade uses to return patches; a batched device version (2-D dominance counting) is round-2 work.
"""
from __future__ import annotations

from .packing import js_key, parse_op_id

INF = float("inf")


def compareOpIds(id1: str, id2: str) -> int:
    """reference src/micromerge.ts:812-827"""
    if id1 == id2:
        return 0
    c1, a1 = parse_op_id(id1)
    c2, a2 = parse_op_id(id2)
    return -1 if (c1 < c2 or (c1 == c2 and js_key(a1) < js_key(a2))) else 1


def ops_to_marks(ops) -> dict:
    """reference src/peritext.ts:294-326; `ops` in Set insertion order (= arrival order of the covering ops)."""
    mark_map, op_id_map = {}, {}
    for op in ops:
        mt = op["markType"]
        if mt != "comment":
            if mt not in op_id_map or compareOpIds(op["opId"], op_id_map[mt]) == 1:
                op_id_map[mt] = op["opId"]
                if op["action"] == "addMark":
                    mark_map[mt] = op.get("attrs") or {"active": True}
                else:
                    mark_map.pop(mt, None)
        else:
            cur = mark_map.get(mt) or []
            if op["action"] == "addMark" and not any(c["id"] == op["attrs"]["id"] for c in cur):
                mark_map[mt] = sorted(cur + [op["attrs"]], key=lambda c: js_key(c["id"]))
            elif op["action"] == "removeMark":
                mark_map[mt] = [c for c in cur if c["id"] != op["attrs"]["id"]]
    return mark_map


class ArrivalHistory:
    """Arrival times of one replica's list ops: what the closed forms need besides final positions."""

    def __init__(self):
        self.t = 0
        self.t_ins: dict[str, int] = {}
        self.t_del: dict[str, int] = {}     # first delete only
        self.marks: list[tuple] = []        # (t, start boundary, end boundary, op), arrival order

    def record(self, op: dict) -> tuple[int, bool]:
        """Registers `op`; returns (its arrival time, whether it can emit a patch)."""
        t = self.t
        self.t += 1
        emits = True
        if op["action"] == "set" and op.get("insert"):
            self.t_ins[op["opId"]] = t
        elif op["action"] == "del" and op.get("key") is None:
            emits = op["elemId"] not in self.t_del
            self.t_del.setdefault(op["elemId"], t)
        elif op["action"] in ("addMark", "removeMark"):
            self.marks.append((t, op["start"], op["end"], op))
        return t, emits


def _slot(b: dict, pos: dict, hist: "ArrivalHistory | None" = None, t: int | None = None) -> float:
    """Slot of boundary `b` as seen by a mark op that arrived at time `t`: an element that is unknown, or that is
    inserted only later at this replica, never matches while walking (src/peritext.ts:236-241)."""
    if b["type"] in ("startOfText", "endOfText") or b["elemId"] not in pos:
        return INF
    if hist is not None and not (b["elemId"] in hist.t_ins and hist.t_ins[b["elemId"]] < t):
        return INF
    return 2 * pos[b["elemId"]] + (1 if b["type"] == "after" else 0)


def derive_patch(op: dict, t: int, pos: dict, hist: ArrivalHistory) -> list[dict]:
    """Patches of list op `op` that arrived at time `t` (already recorded in `hist`); `pos`: elemId -> final position."""
    def present(e):        # element exists at time t
        return e in hist.t_ins and hist.t_ins[e] < t

    def visible(e):
        return present(e) and not (e in hist.t_del and hist.t_del[e] < t)

    act = op["action"]
    if act == "set" and op.get("insert"):
        p = pos[op["opId"]]
        index = sum(1 for e, pe in pos.items() if pe < p and visible(e))
        before = [pe for e, pe in pos.items() if pe < p and present(e)]
        marks = {}
        if before:
            s = 2 * max(before) + 1
            cover = []
            for (tm, sb, eb, q) in hist.marks:
                if tm >= t:
                    break
                qs, qe = _slot(sb, pos, hist, tm), _slot(eb, pos, hist, tm)
                if qe == qs:
                    qe = INF
                if qs <= s < qe:
                    cover.append(q)
            marks = ops_to_marks(cover)
        return [{"path": ["text"], "action": "insert", "index": index, "values": [op["value"]], "marks": marks}]
    if act == "del":
        p = pos[op["elemId"]]
        index = sum(1 for e, pe in pos.items() if pe < p and visible(e))
        return [{"path": ["text"], "action": "delete", "index": index, "count": 1}]
    if act in ("addMark", "removeMark"):
        ps, pe_raw = _slot(op["start"], pos, hist, t), _slot(op["end"], pos, hist, t)
        pe = INF if pe_raw == ps else pe_raw                                   # same slot: start wins, never ends (Q2)
        if ps == INF or ps >= pe:
            return []
        vis_after = sorted(2 * pos[e] + 1 for e in pos if visible(e))
        length = len(vis_after)

        def vis(s):
            return length if s == INF else sum(1 for a in vis_after if a <= s)
        earlier = []
        defined = set()
        for (tm, sb, eb, q) in hist.marks:
            if tm >= t:
                break
            qs, qraw = _slot(sb, pos, hist, tm), _slot(eb, pos, hist, tm)
            qe = INF if qraw == qs else qraw
            earlier.append((qs, qe, q))
            if qs != INF and qs <= qe:
                defined.add(qs)
            if qraw != INF and qraw != qs:
                defined.add(qraw)
        bounds = sorted({ps} | {d for d in defined if ps < d < pe} | ({pe} if pe != INF else set()))
        out = []
        for j, b in enumerate(bounds):
            if b == pe:
                break
            nxt = bounds[j + 1] if j + 1 < len(bounds) else INF
            cover = [q for (qs, qe, q) in earlier if qs != INF and qs <= b < qe]
            if ops_to_marks(cover) == ops_to_marks(cover + [op]):
                continue
            start_i, end_i = vis(b), vis(nxt)
            if end_i > start_i and start_i < length:
                patch = {"action": act, "markType": op["markType"], "path": ["text"], "startIndex": start_i}
                if act == "addMark" and op["markType"] in ("link", "comment"):
                    patch["attrs"] = op["attrs"]
                patch["endIndex"] = min(end_i, length)
                out.append(patch)
        return out
    return []
