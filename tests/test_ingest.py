"""Native wire-format ingest (csrc/ingest.cpp, C-ABI pt_ingest_*) against the Python packer `packing.pack_logs`, which is
the readable specification: same packed records, descriptors, change table and pools on the reference's KAT logs, fuzz
sessions, the links-minimal trace (Symbol fields lost in JSON), unicode / multi-character values and sparse counters.
Host code only: runs without a GPU."""
import json
import os

import numpy as np
import pytest

from oracle.oracle import Micromerge
from peritext_b200.engine import pack_logs_native
from peritext_b200.packing import pack_logs
from tests.harness import GOLDEN, fuzz_session, generateDocs, load_kats, run_concurrent


def assert_same(logs):
    ref = pack_logs(logs, with_changes=True)
    got = pack_logs_native([json.dumps(l) for l in logs])
    assert got.desc.tobytes() == ref.desc.tobytes()
    assert got.insdel.tobytes() == ref.insdel.tobytes()
    assert got.marks.tobytes() == ref.marks.tobytes()
    assert got.values == ref.values
    assert got.link_attrs == ref.link_attrs
    assert got.comment_ids == ref.comment_ids
    assert got.log_actors == ref.log_actors
    for a, b in zip(got.log_counters, ref.log_counters):
        assert (a is None) == (b is None) and (a is None or a.tolist() == b.tolist())
    for name in ("desc", "changes", "deps"):
        assert getattr(got.changes, name).tobytes() == getattr(ref.changes, name).tobytes(), name
    return got


def test_kat_logs():
    logs = []
    for kat in [k for k in load_kats() if k["kind"] == "concurrent"]:
        rec = []
        run_concurrent(Micromerge, kat, record=rec)
        logs += rec
    assert_same(logs)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_sessions(seed):
    _, logs, _ = fuzz_session(Micromerge, 500 + seed, 150)
    assert_same(logs)


def test_links_minimal_trace_without_symbol_fields():
    q = json.load(open(os.path.join(GOLDEN, "links_minimal_queues.json")))["queues"]
    a = [q["doc0"][0], q["doc0"][1], q["doc1"][0], q["doc2"][0]]
    b = [q["doc0"][0], q["doc2"][0], q["doc1"][0], q["doc0"][1]]
    assert_same([a, b])


def test_unicode_values_comments_and_actor_order():
    docs, _, init = generateDocs(Micromerge, "ab", 1)
    d = docs[0]
    c1 = d.change([{"path": ["text"], "action": "insert", "index": 1, "values": [" is great!", "é", "\U0001F600", "", "中"]}])["change"]
    c2 = d.change([{"path": ["text"], "action": "addMark", "startIndex": 0, "endIndex": 3, "markType": "comment", "attrs": {"id": "zé"}},
                   {"path": ["text"], "action": "addMark", "startIndex": 1, "endIndex": 4, "markType": "comment", "attrs": {"id": "a\U0001F600"}},
                   {"path": ["text"], "action": "addMark", "startIndex": 0, "endIndex": 2, "markType": "link", "attrs": {"url": "https://x.y/?q=\"1\"&r=\\"}}])["change"]
    got = assert_same([[init, c1, c2]])
    assert got.values == [" is great!", ""]
    # actor ids ranked in UTF-16 code-unit order (JS string <), not code-point order
    odd = [dict(init, actor="￿"), ]
    logs = [[{"actor": "\U00010000", "seq": 1, "deps": {}, "startOp": 1, "ops": [
        {"opId": "1@\U00010000", "action": "makeList", "obj": "_root", "key": "text"},
        {"opId": "2@\U00010000", "action": "set", "obj": "1@\U00010000", "elemId": "_head", "insert": True, "value": "x"}]},
        {"actor": "￿", "seq": 1, "deps": {"\U00010000": 1}, "startOp": 3, "ops": [
            {"opId": "3@￿", "action": "set", "obj": "1@\U00010000", "elemId": "2@\U00010000", "insert": True, "value": "y"}]}]]
    got = assert_same(logs)
    assert got.log_actors[0] == ["\U00010000", "￿"]      # surrogate pair D800.. sorts before FFFF


def test_sparse_counters_and_errors():
    docs, _, init = generateDocs(Micromerge, "abc", 2)
    big = {"actor": "doc2", "seq": 1, "deps": {"doc1": 1}, "startOp": 5_000_000, "ops": [
        {"opId": "5000000@doc2", "action": "set", "obj": "1@doc1", "elemId": "2@doc1", "insert": True, "value": "X"},
        {"opId": "5000002@doc2", "action": "addMark", "obj": "1@doc1", "start": {"type": "before", "elemId": "5000000@doc2"},
         "end": {"type": "after", "elemId": "3@doc1"}, "markType": "link", "attrs": {"url": "u"}}]}
    got = assert_same([[init, big]])
    assert got.log_counters[0] is not None
    with pytest.raises(ValueError):
        pack_logs_native(['[{"actor": "a", "seq": 1, "ops": [{"opId": "bad", "action": "makeList", "obj": "_root", "key": "text"}]}]'])
    with pytest.raises(ValueError):
        pack_logs_native(['[{"actor": "a", "seq": 1, "ops": ['])
