"""GPU: batched index -> element resolution (pt_batch_query_elements) — the device form of getListElementId (reference
src/micromerge.ts:762-805), which `Micromerge.change` uses to turn visible indices into elemIds (with lookAfterTombstones for
insert positions) and `getCursor` uses for cursors — against the reference function evaluated on the oracle's element
sequence, for every visible index of every replica of seeded fuzz sessions, and on the reference's cursor KATs."""
import numpy as np
import pytest

from oracle.oracle import Micromerge as O
from peritext_b200.micromerge import Micromerge as Facade
from peritext_b200.packing import RangeError, pack_logs
from tests.harness import fuzz_session

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sengine():
    from peritext_b200.engine import BatchEngine
    e = BatchEngine(0, emit_sequence=True)
    yield e
    e.close()


def elem_id(batch, i, rec):
    ins, _ = batch.log_slice(i)
    r = ins[int(rec)]
    cmap = batch.log_counters[i] if batch.log_counters else None
    ctr = int(r["ctr"]) if cmap is None else int(cmap[int(r["ctr"])])
    return f"{ctr}@{batch.log_actors[i][int(r['actor'])]}"


@pytest.mark.parametrize("seed", range(6))
def test_every_visible_index_of_every_replica(sengine, seed):
    docs, logs, _ = fuzz_session(O, 4100 + seed, 140, sync_prob=0.5, full_sync_at_end=bool(seed % 2), remove_comments=True)
    batch = pack_logs(logs)
    merged = sengine.run(batch)
    assert (merged.results["status"] == 0).all()
    qlog, qidx, qflag, want = [], [], [], []
    saw_shift = False
    for i, log in enumerate(logs):
        fresh = O("observer")
        for ch in log:
            fresh.applyChange(ch)
        meta = [[e["elemId"], e["deleted"], e["after"], None] for e in fresh.elements()]
        n_vis = sum(1 for e in meta if not e[1])
        assert n_vis == int(merged.results[i]["n_visible"])
        for k in range(n_vis + 1):
            for flag in (False, True):
                qlog.append(i); qidx.append(k); qflag.append(1 if flag else 0)
                try:
                    want.append(Facade._getListElementId(meta, k, flag))
                except RangeError:
                    want.append(None)
            if k < n_vis and want[-1] != want[-2]:
                saw_shift = True
    got = sengine.query_elements(np.array(qlog, np.uint32), np.array(qidx, np.uint32), np.array(qflag, np.uint32))
    for k in range(len(want)):
        g = None if int(got[k]) == 0xFFFFFFFF else elem_id(batch, qlog[k], got[k])
        assert g == want[k], (qlog[k], qidx[k], qflag[k])
    if seed == 0:
        assert saw_shift        # lookAfterTombstones really moved some answers (tombstones with a defined after-slot exist)


def test_cursor_kats_resolve_through_the_device_query(sengine):
    from tests.harness import generateDocs, load_kats
    for kat in [k for k in load_kats() if k["kind"] == "script" and any(st["do"] == "getCursor" for st in k["steps"])]:
        docs, _, init = generateDocs(O, kat["initialText"])
        logs = [[init], [init]]
        saved = {}
        for st in kat["steps"]:
            d = st["doc"] - 1
            if st["do"] == "change":
                ch = docs[d].change(st["ops"])["change"]; logs[d].append(ch)
                if "save" in st:
                    saved[st["save"]] = ch
            elif st["do"] == "applyChange":
                docs[d].applyChange(saved[st["change"]]); logs[d].append(saved[st["change"]])
            elif st["do"] == "getCursor":
                batch = pack_logs([logs[d]])
                sengine.run(batch)
                rec = sengine.query_elements(np.array([0], np.uint32), np.array([st["index"]], np.uint32))
                assert elem_id(batch, 0, rec[0]) == docs[d].getCursor(["text"], st["index"])["elemId"], kat["name"]
