"""the Patch stream of `applyChange` (reference src/micromerge.ts:661-671 insert, :689-703
delete; src/peritext.ts:198-220, 251-281 marks) has order-independent closed forms over (final sequence position,
arrival time) — `peritext_b200/patches.py`.  This CPU test checks them against the oracle's patch stream on seeded
fuzz sessions (converged and unconverged replicas, zero-width marks): every patch, in order, deep-equal."""
import pytest

from oracle.oracle import Micromerge
from peritext_b200.patches import ArrivalHistory, derive_patch, ops_to_marks
from tests.harness import fuzz_session


def closed_form_patches(log, final_elements, list_id):
    pos = {e["elemId"]: k for k, e in enumerate(final_elements)}
    hist = ArrivalHistory()
    out = []
    for ch in log:
        for op in ch["ops"]:
            if op.get("obj") != list_id:
                continue
            t, emits = hist.record(op)
            if emits:
                out += derive_patch(op, t, pos, hist)
    return out


@pytest.mark.parametrize("seed", range(16))
def test_patch_stream_equals_closed_form(seed):
    _, logs, _ = fuzz_session(Micromerge, 7000 + seed, 120, sync_prob=0.6 if seed % 3 else 1.0,
                              zero_width_prob=0.1 if seed % 2 else 0.0, full_sync_at_end=bool(seed % 4))
    for r, log in enumerate(logs):
        fresh = Micromerge(f"observer{r}")
        got = []
        for ch in log:
            got += [p for p in fresh.applyChange(ch) if p["action"] != "makeList"]
        want = closed_form_patches(log, fresh.elements(), "1@doc1")
        assert got == want, f"replica {r}"


def test_ops_to_marks_matches_reference_rules():
    a = {"opId": "5@a", "action": "addMark", "markType": "strong"}
    r = {"opId": "6@a", "action": "removeMark", "markType": "strong"}
    assert ops_to_marks([a, r]) == {} and ops_to_marks([r, a]) == {}            # LWW by opId, not by order
    c1 = {"opId": "7@a", "action": "addMark", "markType": "comment", "attrs": {"id": "b"}}
    c2 = {"opId": "8@a", "action": "addMark", "markType": "comment", "attrs": {"id": "a"}}
    rm = {"opId": "9@a", "action": "removeMark", "markType": "comment", "attrs": {"id": "b"}}
    assert ops_to_marks([c1, c2]) == {"comment": [{"id": "a"}, {"id": "b"}]}
    assert ops_to_marks([c1, c2, rm]) == {"comment": [{"id": "a"}]}
    assert ops_to_marks([rm]) == {"comment": []}                                  # quirk Q3


@pytest.mark.parametrize("seed", range(8))
def test_patch_stream_typing_runs_two_replicas(seed):
    # longer inserts (typing chains) and a longer initial text: more tombstone-boundary and overlap cases per log
    _, logs, _ = fuzz_session(Micromerge, 7500 + seed, 150, replicas=2, max_chars=6, initial="The Peritext editor", sync_prob=0.5)
    for r, log in enumerate(logs):
        fresh = Micromerge(f"observer{r}")
        got = []
        for ch in log:
            got += [p for p in fresh.applyChange(ch) if p["action"] != "makeList"]
        assert got == closed_form_patches(log, fresh.elements(), "1@doc1"), f"replica {r}"


def test_patch_kats_closed_form():
    """The reference's four exact Patch KATs (test/micromerge.ts:915-1029) through the closed forms."""
    from tests.harness import generateDocs, load_kats
    for kat in [k for k in load_kats() if k["kind"] == "script" and any("expectPatches" in st for st in k["steps"])]:
        docs, _, init = generateDocs(Micromerge, kat["initialText"])
        logs = [[init], [init]]
        saved = {}
        for st in kat["steps"]:
            d = st["doc"] - 1
            if st["do"] == "change":
                ch = docs[d].change(st["ops"])["change"]
                logs[d].append(ch)
                if "save" in st:
                    saved[st["save"]] = ch
            elif st["do"] == "applyChange":
                docs[d].applyChange(saved[st["change"]])
                logs[d].append(saved[st["change"]])
                allp = closed_form_patches(logs[d], docs[d].elements(), "1@doc1")
                n = len(st["expectPatches"])
                assert allp[-n:] == st["expectPatches"], kat["name"]
