"""Property tests: the order-independent closed form (tests/kernel_model.py — the executable model of what the CUDA
kernels compute) equals the sequential oracle on seeded fuzz sessions, on the reference KATs and on the quirk corners
of SURVEY.md §9.3.  Runs on CPU; this is the algorithm's proof obligation, independent of the device code."""
import pytest

from oracle.oracle import Micromerge
from oracle.packed import replay_packed
from peritext_b200.packing import decode_spans, pack_logs
from tests import kernel_model
from tests.harness import fuzz_session, generateDocs, load_kats, run_concurrent


def check_equal(logs, expect_spans=None):
    b = pack_logs(logs)
    ref, _ = replay_packed(b)
    # both cuts of the mark overlay: element space (CTA kernel) and visible space (warp kernel)
    for form in ("visible", "element"):
        got = kernel_model.merge_batch(b, form=form)
        for i in range(b.n_logs):
            assert got.canonical(i) == ref.canonical(i), f"log {i} ({form} form)"
            if expect_spans is not None:
                assert decode_spans(b, got, i) == expect_spans[i]
    return b, got


@pytest.mark.parametrize("kat", [k for k in load_kats() if k["kind"] == "concurrent"], ids=lambda k: f"L{k['line']}")
def test_kats(kat):
    rec = []
    run_concurrent(Micromerge, kat, record=rec)
    check_equal(rec, [kat["expectedResult"]] * 2)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_sessions(seed):
    docs, logs, _ = fuzz_session(Micromerge, seed, 150)
    spans = [d.getTextWithFormatting() for d in docs]
    assert spans[0] == spans[1] == spans[2]
    check_equal(logs, spans)


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_partial_sync(seed):
    # replicas that have NOT converged: each log is still a causally valid prefix-closed set of changes
    docs, logs, _ = fuzz_session(Micromerge, 1000 + seed, 120, sync_prob=0.3, full_sync_at_end=False)
    check_equal(logs, [d.getTextWithFormatting() for d in docs])


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_typing_runs(seed):
    # longer inserts -> real typing chains (run contraction path), 2 replicas
    docs, logs, _ = fuzz_session(Micromerge, 2000 + seed, 100, replicas=2, max_chars=6, initial="The Peritext editor")
    check_equal(logs, [d.getTextWithFormatting() for d in docs])


def test_quirk_zero_width_inclusive_mark_runs_to_end():
    # Q2: inclusive mark with startIndex == endIndex: start and end are the same slot, the start branch wins
    # (reference src/peritext.ts:236-241), so the op never ends.
    docs, logs, _ = fuzz_session(Micromerge, 7, 0, replicas=2, initial="abcdef")
    r = docs[0].change([{"path": ["text"], "action": "addMark", "startIndex": 2, "endIndex": 2, "markType": "strong"}])
    logs[0].append(r["change"])
    spans = docs[0].getTextWithFormatting()
    assert spans == [{"marks": {}, "text": "ab"}, {"marks": {"strong": {"active": True}}, "text": "cdef"}]
    check_equal([logs[0]], [spans])


@pytest.mark.parametrize("seed", range(10))
def test_fuzz_with_zero_width_marks(seed):
    docs, logs, _ = fuzz_session(Micromerge, 3000 + seed, 120, zero_width_prob=0.3)
    check_equal(logs, [d.getTextWithFormatting() for d in docs])


def test_quirk_comment_remove_only_gives_empty_array_key():
    # Q3: a region covered only by a comment removeMark has {comment: []}, which is not isEqual to {}
    docs, _, init = generateDocs(Micromerge, "abcdef", 1)
    d = docs[0]
    r = d.change([{"path": ["text"], "action": "removeMark", "startIndex": 1, "endIndex": 3, "markType": "comment", "attrs": {"id": "x"}}])
    spans = d.getTextWithFormatting()
    assert spans == [{"marks": {}, "text": "a"}, {"marks": {"comment": []}, "text": "bc"}, {"marks": {}, "text": "def"}]
    check_equal([[init, r["change"]]], [spans])


def test_comment_add_then_remove_and_readd():
    docs, _, init = generateDocs(Micromerge, "abcdefgh", 1)
    d = docs[0]
    chs = [init]
    for op in [dict(action="addMark", startIndex=0, endIndex=6, attrs={"id": "c1"}),
               dict(action="removeMark", startIndex=2, endIndex=4, attrs={"id": "c1"}),
               dict(action="addMark", startIndex=3, endIndex=8, attrs={"id": "c1"}),
               dict(action="addMark", startIndex=1, endIndex=5, attrs={"id": "c0"}),
               dict(action="removeMark", startIndex=0, endIndex=8, attrs={"id": "zz"})]:
        chs.append(d.change([{"path": ["text"], "markType": "comment", **op}])["change"])
    check_equal([chs], [d.getTextWithFormatting()])


def test_tombstones_and_empty_docs():
    docs, _, init = generateDocs(Micromerge, "abc", 1)
    d = docs[0]
    c1 = d.change([{"path": ["text"], "action": "addMark", "startIndex": 0, "endIndex": 3, "markType": "em"}])["change"]
    c2 = d.change([{"path": ["text"], "action": "delete", "index": 0, "count": 3}])["change"]
    assert d.getTextWithFormatting() == []
    check_equal([[init, c1, c2]], [[]])
    # an entirely empty list
    e = Micromerge("doc1")
    c0 = e.change([{"path": [], "action": "makeList", "key": "text"}])["change"]
    check_equal([[c0]], [[]])


def test_multi_character_values_and_unicode():
    docs, _, init = generateDocs(Micromerge, "ab", 1)
    d = docs[0]
    c1 = d.change([{"path": ["text"], "action": "insert", "index": 1, "values": [" is great!", "é", "\U0001F600"]}])["change"]
    check_equal([[init, c1]], [d.getTextWithFormatting()])


def test_error_statuses():
    import numpy as np
    docs, logs, _ = fuzz_session(Micromerge, 5, 30)
    b = pack_logs(logs[:1])
    # delete of an element that is not in the log -> "List element not found" (reference src/micromerge.ts:752)
    bad = b.select([0])
    bad.insdel = bad.insdel.copy()
    k = int(np.nonzero((bad.insdel["payload"] >> 30) == 1)[0][0])
    bad.insdel[k]["ref_ctr"] = bad.desc[0]["max_ctr"]
    bad.insdel[k]["ref_actor"] = bad.desc[0]["n_actors"] - 1
    if kernel_model.merge_batch(bad).results[0]["status"] == 0:   # that id happened to exist; use an unused one
        pytest.skip("id exists")
    ref, _ = replay_packed(bad)
    assert ref.results[0]["status"] == 1
    assert kernel_model.merge_batch(bad).results[0]["status"] == 1


def test_sparse_counters_are_reranked_densely():
    """A peer may pick any startOp (reference src/micromerge.ts:511 only takes the max): counters far beyond the op count
    are re-ranked by the packer; order — all that compareOpIds looks at — is preserved."""
    docs, _, init = generateDocs(Micromerge, "abc", 2)
    d1, d2 = docs
    big = {"actor": "doc2", "seq": 1, "deps": {"doc1": 1}, "startOp": 5_000_000, "ops": [
        {"opId": "5000000@doc2", "action": "set", "obj": "1@doc1", "elemId": "2@doc1", "insert": True, "value": "X"},
        {"opId": "5000001@doc2", "action": "set", "obj": "1@doc1", "elemId": "5000000@doc2", "insert": True, "value": "Y"},
        {"opId": "5000002@doc2", "action": "addMark", "obj": "1@doc1", "start": {"type": "before", "elemId": "5000000@doc2"},
         "end": {"type": "after", "elemId": "3@doc1"}, "markType": "link", "attrs": {"url": "u"}}]}
    d1.applyChange(big)
    c = d1.change([{"path": ["text"], "action": "insert", "index": 1, "values": ["Z"]}])["change"]   # 5000003@doc1, after 'a'
    logs = [[init, big, c]]
    b = pack_logs(logs)
    assert int(b.desc[0]["max_ctr"]) < 20 and b.log_counters[0] is not None
    check_equal(logs, [d1.getTextWithFormatting()])
