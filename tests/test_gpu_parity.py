"""GPU parity tests (run on the B200 box): the CUDA engine, called through the C-ABI, against the CPU oracle on the
same inputs — bit-exact on every output array (tokens, spans, comment lists, digests), and against the reference's
own expected spans for the KATs.  /root/reference is not needed at run time."""
import numpy as np
import pytest

from oracle.oracle import Micromerge as OracleMicromerge
from oracle.packed import replay_packed
from peritext_b200.packing import decode_spans, pack_logs
from tests.harness import fuzz_session, generateDocs, load_kats, run_concurrent

pytestmark = pytest.mark.gpu


def assert_batch_equal(batch, got, ref):
    assert got.results["status"].tolist() == ref.results["status"].tolist()
    for i in range(batch.n_logs):
        assert got.canonical(i) == ref.canonical(i), f"log {i}"


def test_kats_through_engine(engine):
    logs, expected = [], []
    for kat in [k for k in load_kats() if k["kind"] == "concurrent"]:
        rec = []
        run_concurrent(OracleMicromerge, kat, record=rec)
        logs += rec
        expected += [kat["expectedResult"]] * 2
    batch = pack_logs(logs)
    got = engine.run(batch)
    ref, _ = replay_packed(batch)
    assert_batch_equal(batch, got, ref)
    for i, exp in enumerate(expected):
        assert decode_spans(batch, got, i) == exp


def test_links_minimal_trace(engine):
    import json, os
    from tests.harness import GOLDEN
    q = json.load(open(os.path.join(GOLDEN, "links_minimal_queues.json")))["queues"]
    a = [q["doc0"][0], q["doc0"][1], q["doc1"][0], q["doc2"][0]]
    b = [q["doc0"][0], q["doc2"][0], q["doc1"][0], q["doc0"][1]]
    batch = pack_logs([a, b])
    got = engine.run(batch)
    expected = [{"marks": {"link": {"url": "https://inkandswitch.com/pushpin"}}, "text": "ABC9ee09150DE"}]
    assert decode_spans(batch, got, 0) == expected
    assert decode_spans(batch, got, 1) == expected
    assert got.results[0]["digest"].tolist() == got.results[1]["digest"].tolist()


def test_fuzz_sessions(engine):
    logs, spans = [], []
    for seed in range(60):
        docs, lg, _ = fuzz_session(OracleMicromerge, seed, 150)
        logs += lg
        spans += [d.getTextWithFormatting() for d in docs]
    for seed in range(20):
        docs, lg, _ = fuzz_session(OracleMicromerge, 2000 + seed, 150, replicas=2, max_chars=6, initial="The Peritext editor")
        logs += lg
        spans += [d.getTextWithFormatting() for d in docs]
    for seed in range(20):
        docs, lg, _ = fuzz_session(OracleMicromerge, 1000 + seed, 120, sync_prob=0.3, full_sync_at_end=False)
        logs += lg
        spans += [d.getTextWithFormatting() for d in docs]
    for seed in range(10):
        docs, lg, _ = fuzz_session(OracleMicromerge, 3000 + seed, 120, zero_width_prob=0.3)
        logs += lg
        spans += [d.getTextWithFormatting() for d in docs]
    batch = pack_logs(logs)
    got = engine.run(batch)
    ref, _ = replay_packed(batch, threads=4)
    assert_batch_equal(batch, got, ref)
    for i in range(0, batch.n_logs, 7):
        assert decode_spans(batch, got, i) == spans[i]


def test_quirks_and_edges(engine):
    logs = []
    # Q3 comment remove only; empty list; fully deleted; multi-char values
    docs, _, init = generateDocs(OracleMicromerge, "abcdef", 1)
    r = docs[0].change([{"path": ["text"], "action": "removeMark", "startIndex": 1, "endIndex": 3, "markType": "comment", "attrs": {"id": "x"}}])
    logs.append([init, r["change"]])
    e = OracleMicromerge("doc1")
    logs.append([e.change([{"path": [], "action": "makeList", "key": "text"}])["change"]])
    docs, _, init = generateDocs(OracleMicromerge, "abc", 1)
    c1 = docs[0].change([{"path": ["text"], "action": "addMark", "startIndex": 0, "endIndex": 3, "markType": "em"}])["change"]
    c2 = docs[0].change([{"path": ["text"], "action": "delete", "index": 0, "count": 3}])["change"]
    logs.append([init, c1, c2])
    docs, _, init = generateDocs(OracleMicromerge, "ab", 1)
    c1 = docs[0].change([{"path": ["text"], "action": "insert", "index": 1, "values": [" is great!", "é", "\U0001F600"]}])["change"]
    logs.append([init, c1])
    # Q2 zero-width inclusive mark runs to the end of the text
    docs, _, init = generateDocs(OracleMicromerge, "abcdef", 1)
    c1 = docs[0].change([{"path": ["text"], "action": "addMark", "startIndex": 2, "endIndex": 2, "markType": "strong"}])["change"]
    logs.append([init, c1])
    batch = pack_logs(logs)
    got = engine.run(batch)
    ref, _ = replay_packed(batch)
    assert_batch_equal(batch, got, ref)
    assert decode_spans(batch, got, 0) == [{"marks": {}, "text": "a"}, {"marks": {"comment": []}, "text": "bc"}, {"marks": {}, "text": "def"}]
    assert decode_spans(batch, got, 1) == []
    assert decode_spans(batch, got, 2) == []
    assert decode_spans(batch, got, 4) == [{"marks": {}, "text": "ab"}, {"marks": {"strong": {"active": True}}, "text": "cdef"}]


def test_error_status_element_not_found(engine):
    docs, logs, _ = fuzz_session(OracleMicromerge, 5, 40)
    batch = pack_logs(logs)
    bad = batch.select([0, 1])
    bad.insdel = bad.insdel.copy()
    ins0, _ = bad.log_slice(0)
    k = int(np.nonzero((ins0["payload"] >> 30) == 1)[0][0])
    used = {(int(r["ctr"]), int(r["actor"])) for r in ins0 if (int(r["payload"]) >> 30) == 0}
    free = next((c, a) for c in range(int(bad.desc[0]["max_ctr"]), 0, -1) for a in range(int(bad.desc[0]["n_actors"])) if (c, a) not in used)
    bad.insdel[k]["ref_ctr"], bad.insdel[k]["ref_actor"] = free
    got = engine.run(bad)
    ref, _ = replay_packed(bad)
    assert got.results["status"].tolist() == [1, 0] == ref.results["status"].tolist()   # src/micromerge.ts:752
    assert got.canonical(1) == ref.canonical(1)


def test_cuda_graph_path_on_a_user_stream():
    """On a non-default stream the engine replays its launch sequence as a CUDA graph; results must not change,
    also across re-uploads of different batches on one handle."""
    import torch
    from peritext_b200.engine import BatchEngine
    s = torch.cuda.Stream()
    eng = BatchEngine(0, stream=s.cuda_stream)
    for seed0 in (500, 600):
        logs = []
        for seed in range(6):
            _, lg, _ = fuzz_session(OracleMicromerge, seed0 + seed, 100)
            logs += lg
        batch = pack_logs(logs)
        eng.upload(batch)
        for _ in range(3):
            eng.merge()
        got = eng.download()
        ref, _ = replay_packed(batch, threads=2)
        assert_batch_equal(batch, got, ref)
    eng.close()


def test_pipelined_engine_equals_single_handle():
    from peritext_b200 import workload
    from peritext_b200.engine import BatchEngine, PipelinedEngine
    batch = workload.generate("c3", n_docs=24, ops_per_doc=2000)
    eng = BatchEngine(0)
    whole = eng.run(batch)
    pipe = PipelinedEngine(0, chunks=4)
    parts = pipe.run(batch, copy=True)
    assert sum(p.results.shape[0] for p in parts) == batch.n_logs
    k = 0
    for p in parts:
        for i in range(p.results.shape[0]):
            assert p.canonical(i) == whole.canonical(k), k
            k += 1
    pipe.close(); eng.close()


def test_run_compressed_upload_gives_identical_results():
    from peritext_b200 import workload
    from peritext_b200.engine import BatchEngine, PipelinedEngine, compress_runs
    for cfg in ("c2", "c3", "c4"):
        batch = workload.generate(cfg, n_docs=12, ops_per_doc=1500)
        eng = BatchEngine(0)
        whole = eng.run(batch)
        runs = compress_runs(batch)
        eng.upload_runs(runs); eng.merge()
        viaruns = eng.download()
        for i in range(batch.n_logs):
            assert viaruns.canonical(i) == whole.canonical(i), (cfg, i)
        pipe = PipelinedEngine(0, chunks=3)
        parts = pipe.run(runs, copy=True)
        k = 0
        for p in parts:
            for i in range(p.results.shape[0]):
                assert p.canonical(i) == whole.canonical(k), (cfg, k)
                k += 1
        assert k == batch.n_logs
        pipe.close(); eng.close()
