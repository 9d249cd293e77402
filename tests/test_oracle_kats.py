"""Pins the CPU oracle against the reference's own known-answer tests (tests/golden/kats.json, transcribed from
reference test/micromerge.ts by tests/golden/make_kats.py).  This is what makes the oracle trustworthy."""
import pytest

from oracle.oracle import Micromerge, RangeError, compareOpIds
from tests.harness import accumulatePatches, generateDocs, load_kats, run_concurrent

KATS = load_kats()


def test_kat_count():
    assert len(KATS) == 46
    assert sum(k["kind"] == "concurrent" for k in KATS) == 31


@pytest.mark.parametrize("kat", [k for k in KATS if k["kind"] == "concurrent"], ids=lambda k: f"L{k['line']}")
def test_concurrent_writes(kat):
    docs, patch_lists = run_concurrent(Micromerge, kat)
    expected = kat["expectedResult"]
    # batch codepath (reference test/micromerge.ts:78-79)
    assert docs[0].getTextWithFormatting(["text"]) == expected
    assert docs[1].getTextWithFormatting(["text"]) == expected
    # incremental patches converge to the same state (:84-85)
    assert accumulatePatches(patch_lists[0]) == expected
    assert accumulatePatches(patch_lists[1]) == expected


@pytest.mark.parametrize("kat", [k for k in KATS if k["kind"] == "script"], ids=lambda k: f"L{k['line']}")
def test_scripted(kat):
    docs, _, _ = generateDocs(Micromerge, kat["initialText"])
    saved = {}
    for st in kat["steps"]:
        doc = docs[st["doc"] - 1]
        do = st["do"]
        if do == "change":
            r = doc.change(st["ops"])
            if "save" in st:
                saved[st["save"]] = r["change"]
        elif do == "applyChange":
            patches = doc.applyChange(saved[st["change"]])
            if "expectPatches" in st:
                assert patches == st["expectPatches"]
        elif do == "expectRootText":
            assert doc.root["text"] == st["value"]
        elif do == "expectRootTextJoined":
            assert "".join(doc.root["text"]) == st["value"]
        elif do == "expectSpans":
            assert doc.getTextWithFormatting(["text"]) == st["value"]
        elif do == "getCursor":
            saved[st["save"]] = doc.getCursor(["text"], st["index"])
        elif do == "resolveCursor":
            assert doc.resolveCursor(saved[st["cursor"]]) == st["expect"]
        else:
            raise AssertionError(do)


def test_compare_op_ids():
    # reference src/micromerge.ts:812-827
    assert compareOpIds("3@a", "3@a") == 0
    assert compareOpIds("2@z", "10@a") == -1          # numeric, not lexicographic, counter
    assert compareOpIds("10@a", "2@z") == 1
    assert compareOpIds("5@doc1", "5@doc2") == -1     # tie -> actor string order
    assert compareOpIds("5@doc2", "5@doc1") == 1
    assert compareOpIds("5@doc10", "5@doc2") == -1    # JS string order: "doc10" < "doc2"
    assert compareOpIds("5@Z", "5@a") == -1           # code-unit order: upper case first


def test_admission_errors():
    # reference src/micromerge.ts:501-509: RangeError before any mutation, retry-safe
    docs, _, _ = generateDocs(Micromerge, "abc")
    d1, d2 = docs
    c1 = d1.change([{"path": ["text"], "action": "insert", "index": 3, "values": ["d"]}])["change"]
    c2 = d1.change([{"path": ["text"], "action": "insert", "index": 4, "values": ["e"]}])["change"]
    with pytest.raises(RangeError, match="Expected sequence number 2, got 3"):
        d2.applyChange(c2)
    d3 = Micromerge("doc3")
    with pytest.raises(RangeError, match="Missing dependency"):
        d3.applyChange(d2.change([{"path": ["text"], "action": "insert", "index": 0, "values": ["x"]}])["change"])
    d2.applyChange(c1)
    d2.applyChange(c2)
    assert "".join(d2.root["text"]) == "xabcde"


def test_links_minimal_trace():
    """BASELINE config 1: two replicas merge traces/links-minimal.json in different causal orders.
    The trace's queues are committed as tests/golden/links_minimal_queues.json (inputs only; the outputs recorded
    in the reference's trace file are the diverged result of a since-fixed bug)."""
    import json, os
    from tests.harness import GOLDEN
    q = json.load(open(os.path.join(GOLDEN, "links_minimal_queues.json")))["queues"]
    a = Micromerge("replicaA"); b = Micromerge("replicaB")
    for ch in [q["doc0"][0], q["doc0"][1], q["doc1"][0], q["doc2"][0]]:
        a.applyChange(ch)
    for ch in [q["doc0"][0], q["doc2"][0], q["doc1"][0], q["doc0"][1]]:
        b.applyChange(ch)
    expected = [{"marks": {"link": {"url": "https://inkandswitch.com/pushpin"}}, "text": "ABC9ee09150DE"}]
    assert a.getTextWithFormatting(["text"]) == expected
    assert b.getTextWithFormatting(["text"]) == expected
