"""GPU parity on the BASELINE.json workload shapes: generated batches through the engine vs the oracle replay
(bit-exact on every array), and — at sizes the O(N^2) oracle cannot replay in full — size-independent properties:
every replica of a document yields the same digest (the fuzz harness's convergence assertion, reference
test/fuzz.ts:277-278), statuses are clean, and a deterministic sample of logs matches the oracle exactly."""
import numpy as np
import pytest

from oracle.packed import replay_packed
from peritext_b200 import workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from peritext_b200.engine import BatchEngine
    e = BatchEngine(0)
    yield e
    e.close()


def full_compare(engine, batch, threads=8):
    got = engine.run(batch)
    ref, _ = replay_packed(batch, threads=threads)
    assert got.results["status"].tolist() == ref.results["status"].tolist()
    assert (got.results["status"] == 0).all()
    for name in ("n_elems", "n_visible", "n_spans"):
        assert got.results[name].tolist() == ref.results[name].tolist(), name
    assert got.results["digest"].tolist() == ref.results["digest"].tolist()
    for i in range(batch.n_logs):
        assert got.canonical(i) == ref.canonical(i), f"log {i}"
    return got


@pytest.mark.parametrize("cfg,n_docs,ops", [("c2", 24, 2500), ("c3", 24, 2500), ("c4", 300, 1000), ("c5", 2, 6000),
                                            ("c2", 3, 10000), ("c3", 3, 10000)])
def test_generated_workloads_match_oracle(engine, cfg, n_docs, ops):
    batch = workload.generate(cfg, n_docs=n_docs, ops_per_doc=ops, n_marks=800)
    got = full_compare(engine, batch)
    R = batch.meta["replicas"]
    dig = got.results["digest"].reshape(n_docs, R, 2)
    assert (dig == dig[:, :1, :]).all()


def test_large_log_uses_wide_index_path(engine):
    # > 32000 records in one log: the u32-index instantiation, arena spills to the global slab
    batch = workload.generate("c2", n_docs=1, ops_per_doc=40000)
    assert int(batch.desc["n_insdel"].max()) > 32000
    got = engine.run(batch)
    ref, _ = replay_packed(batch, first=0, count=1)
    assert got.canonical(0) == ref.canonical(0)
    assert got.results[1]["status"] == 0 and got.results[0]["digest"].tolist() == got.results[1]["digest"].tolist()


@pytest.mark.parametrize("cfg", ["c2", "c3"])
def test_full_size_properties(engine, cfg):
    batch = workload.generate(cfg, n_docs=300)            # 600 logs x 10K ops
    got = engine.run(batch)
    assert (got.results["status"] == 0).all()
    R = batch.meta["replicas"]
    dig = got.results["digest"].reshape(-1, R, 2)
    assert (dig == dig[:, :1, :]).all()                    # replicas converged
    assert (got.results["n_visible"] <= got.results["n_elems"]).all()
    # deterministic sample of logs against the oracle
    idx = list(range(0, batch.n_logs, 75))
    sub = batch.select(idx)
    ref, _ = replay_packed(sub, threads=8)
    for k, i in enumerate(idx):
        a, b = got.canonical(i), ref.canonical(k)
        assert a[:4] == b[:4] and a[4] == b[4] and a[6] == b[6], f"log {i}"
        assert [s[:3] + (s[3],) for s in a[5]] == [s[:3] + (s[3],) for s in b[5]]


def test_merge_is_idempotent_and_order_of_logs_irrelevant(engine):
    batch = workload.generate("c3", n_docs=16, ops_per_doc=3000)
    a = engine.run(batch)
    b = engine.run(batch)
    assert a.results.tobytes() == b.results.tobytes()
    perm = list(reversed(range(batch.n_logs)))
    c = engine.run(batch.select(perm))
    for k, i in enumerate(perm):
        assert c.canonical(k) == a.canonical(i)


def test_empty_batch_and_empty_logs(engine):
    import numpy as np
    from peritext_b200.packing import DESC_DT, INSDEL_DT, MARK_DT, PackedBatch
    empty = PackedBatch(np.zeros(0, DESC_DT), np.zeros(0, INSDEL_DT), np.zeros(0, MARK_DT))
    got = engine.run(empty)
    assert got.results.shape[0] == 0
    # logs without any record between real ones
    batch = workload.generate("c3", n_docs=2, ops_per_doc=300)
    d = np.zeros(batch.n_logs + 2, DESC_DT)
    d[0] = (0, 0, 0, 0, 1, 1)
    d[1:-1] = batch.desc
    d[-1] = (len(batch.insdel), len(batch.marks), 0, 0, 1, 1)
    mixed = PackedBatch(d, batch.insdel, batch.marks, meta=dict(batch.meta))
    got = engine.run(mixed)
    ref, _ = replay_packed(mixed)
    for i in range(mixed.n_logs):
        assert got.canonical(i) == ref.canonical(i), i
    assert got.results[0]["n_spans"] == 0 and got.results[-1]["n_visible"] == 0


@pytest.mark.parametrize("ops", [31900, 32100])
def test_index_width_boundary(engine, ops):
    # just below / above the u16 -> u32 index switch (32000 records per log); the O(N^2) oracle replays one replica,
    # the other one is checked through the convergence digest
    batch = workload.generate("c2", n_docs=1, ops_per_doc=ops)
    got = engine.run(batch)
    ref, _ = replay_packed(batch, first=0, count=1)
    assert got.canonical(0) == ref.canonical(0)
    assert got.results[1]["status"] == 0 and got.results[0]["digest"].tolist() == got.results[1]["digest"].tolist()
