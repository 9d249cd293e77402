"""GPU: the admission pre-pass (pt_batch_upload_changes + admit_kernel) against the oracle's applyChange, which throws
RangeError "Expected sequence number" / "Missing dependency" before mutating (reference src/micromerge.ts:501-509): valid
logs are untouched, tampered logs report status 6 / 7 and the index of the change the reference would have thrown at."""
import copy
import json

import numpy as np
import pytest

from oracle.oracle import Micromerge as O
from oracle.packed import replay_packed
from peritext_b200.engine import pack_logs_native
from peritext_b200.packing import decode_spans, pack_logs
from tests.harness import fuzz_session

pytestmark = pytest.mark.gpu


def oracle_admission(log):
    """(status, index of the rejected change) the way the reference reports it: apply in order, stop at the first throw."""
    d = O("reader")
    for k, ch in enumerate(log):
        try:
            d.applyChange(ch)
        except Exception as e:       # the oracle raises its RangeError equivalents
            msg = str(e)
            if "Expected sequence number" in msg:
                return 6, k
            if "Missing dependency" in msg:
                return 7, k
            raise
    return 0, None


def tampered_logs():
    logs = []
    for seed in range(6):
        _, ls, _ = fuzz_session(O, 900 + seed, 80)
        logs += ls
    out = [("valid-%d" % i, l) for i, l in enumerate(logs[:6])]
    rng = np.random.default_rng(5)
    for i, l in enumerate(logs[6:18]):
        l = copy.deepcopy(l)
        kind = i % 4
        k = int(rng.integers(1, len(l) - 1))
        if kind == 0:
            del l[k]                                   # a dropped change: seq gap or missing dependency, whichever comes first
        elif kind == 1:
            l[k], l[k + 1] = l[k + 1], l[k]            # swapped neighbours (may still be causally fine)
        elif kind == 2:
            l.insert(k, l[k])                          # a change delivered twice
        else:
            l[k]["deps"] = dict(l[k].get("deps") or {}, ghost=1)     # a dependency on an actor nobody has heard of
        out.append(("tampered-%d-%d" % (kind, i), l))
    return out


def test_admission_statuses_match_the_oracle(engine):
    cases = tampered_logs()
    want = [oracle_admission(l) for _, l in cases]
    assert any(w[0] == 6 for w in want) and any(w[0] == 7 for w in want) and any(w[0] == 0 for w in want)
    for pack in (lambda ls: pack_logs(ls, with_changes=True), lambda ls: pack_logs_native([json.dumps(l) for l in ls])):
        batch = pack([l for _, l in cases])
        got = engine.run(batch)
        ref, _ = replay_packed(batch)
        for i, ((name, log), (st, idx)) in enumerate(zip(cases, want)):
            r = got.results[i]
            if st:
                assert (int(r["status"]), int(r["n_elems"])) == (st, idx), name
                assert int(r["n_visible"]) == 0 and int(r["n_spans"]) == 0
            else:
                assert int(r["status"]) == int(ref.results[i]["status"]), name
                assert got.canonical(i) == ref.canonical(i), name


def test_valid_logs_are_unchanged_by_the_pre_pass(engine):
    _, logs, _ = fuzz_session(O, 77, 150)
    a = engine.run(pack_logs(logs))
    b = engine.run(pack_logs(logs, with_changes=True))
    for i in range(len(logs)):
        assert int(b.results[i]["status"]) == 0
        assert a.canonical(i)[1:] == b.canonical(i)[1:]
        assert decode_spans(pack_logs(logs), a, i) == decode_spans(pack_logs(logs, with_changes=True), b, i)
