"""Corners where the closed form has to follow the reference's ARRIVAL-dependent behaviour exactly (round-1 advisor
findings): mark boundaries that name an element inserted LATER in the same log (reference src/peritext.ts:236-241: the
walk never matches them), and add/remove of ONE comment id by concurrent changes (src/peritext.ts:314-322 folds in Set
order = arrival order, no opId comparison — quirk Q4).  CPU: kernel model and packed oracle replay vs the oracle document."""
import itertools

import pytest

from oracle.oracle import Micromerge
from oracle.packed import replay_packed
from peritext_b200.packing import decode_spans, pack_logs
from tests import kernel_model
from tests.harness import generateDocs


def noncausal_logs():
    """Each case: (name, [changes]) for one replica; the crafted change applies a mark BEFORE inserting its boundary element."""
    cases = []
    for which in ("start", "end", "both", "end-then-remove"):
        docs, _, init = generateDocs(Micromerge, "abcd", 1)
        d = docs[0]
        # opIds: 1@doc1 makeList, 2..5@doc1 = a b c d
        new_elem = "7@doc1"
        start = {"type": "before", "elemId": new_elem if which in ("start", "both") else "3@doc1"}
        end = {"type": "after", "elemId": new_elem if which in ("end", "both", "end-then-remove") else "5@doc1"}
        ops = [{"opId": "6@doc1", "action": "addMark", "obj": "1@doc1", "start": start, "end": end, "markType": "strong"},
               {"opId": "7@doc1", "action": "set", "obj": "1@doc1", "elemId": "4@doc1", "insert": True, "value": "X"}]
        if which == "end-then-remove":
            ops.append({"opId": "8@doc1", "action": "removeMark", "obj": "1@doc1", "start": {"type": "before", "elemId": "7@doc1"},
                        "end": {"type": "after", "elemId": "5@doc1"}, "markType": "strong"})
        ch = {"actor": "doc1", "seq": 2, "deps": {"doc1": 1}, "startOp": 6, "ops": ops}
        # apply through a second replica (doc1 itself would reject its own seq)
        r = Micromerge("doc2")
        r.applyChange(init); r.applyChange(ch)
        cases.append((which, [init, ch], r.getTextWithFormatting()))
    return cases


@pytest.mark.parametrize("case", noncausal_logs(), ids=lambda c: c[0])
def test_mark_boundary_inserted_later_is_never_matched(case):
    name, log, spans = case
    b = pack_logs([log])
    ref, _ = replay_packed(b)
    assert decode_spans(b, ref, 0) == spans                       # packed replay == oracle document
    for form in ("element", "visible"):                            # the CTA kernel's and the warp kernel's cut of the mark overlay
        got = kernel_model.merge_batch(b, form=form)
        assert got.canonical(0) == ref.canonical(0)
        assert decode_spans(b, got, 0) == spans
    if name == "start":
        assert spans == [{"marks": {}, "text": "abcXd"}]           # the op never enters DURING: a no-op
    if name == "end":
        assert spans == [{"marks": {}, "text": "a"}, {"marks": {"strong": {"active": True}}, "text": "bcXd"}]   # never ends


def q4_logs():
    """add (5@C... here 7@doc3) and remove (7@doc2) of one comment id by concurrent changes, both arrival orders."""
    out = []
    docs, _, init = generateDocs(Micromerge, "abc", 3)
    add = docs[2].change([{"path": ["text"], "action": "addMark", "startIndex": 0, "endIndex": 3, "markType": "comment", "attrs": {"id": "k"}}])["change"]
    rem = docs[1].change([{"path": ["text"], "action": "removeMark", "startIndex": 0, "endIndex": 2, "markType": "comment", "attrs": {"id": "k"}}])["change"]
    for order in itertools.permutations([add, rem]):
        r = Micromerge("reader")
        r.applyChange(init)
        for ch in order:
            r.applyChange(ch)
        out.append(([init, *order], r.getTextWithFormatting()))
    return out


@pytest.mark.parametrize("idx", [0, 1])
def test_concurrent_add_remove_of_one_comment_id_follows_arrival_order(idx):
    log, spans = q4_logs()[idx]
    b = pack_logs([log])
    ref, _ = replay_packed(b)
    assert decode_spans(b, ref, 0) == spans
    for form in ("element", "visible"):
        got = kernel_model.merge_batch(b, form=form)
        assert got.canonical(0) == ref.canonical(0)
        assert decode_spans(b, got, 0) == spans


def test_q4_orders_really_differ_in_the_reference():
    (l0, s0), (l1, s1) = q4_logs()
    assert s0 != s1      # the reference itself does not converge here; the engine reproduces each replica's own result
