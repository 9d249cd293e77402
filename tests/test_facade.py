"""`peritext_b200.Micromerge` facade: the reference's admission behaviour on CPU, and (gpu) the reference's own test
scenarios driven through it."""
import pytest

from oracle.oracle import Micromerge as OracleMicromerge
from peritext_b200 import RangeError
from peritext_b200.micromerge import Micromerge
from tests.harness import generateDocs, load_kats, run_concurrent


def test_admission_errors_before_mutation():
    # reference src/micromerge.ts:501-509 and test/merge.ts:12-17 (retry relies on no mutation)
    docs, _, init = generateDocs(OracleMicromerge, "abc")
    c1 = docs[0].change([{"path": ["text"], "action": "insert", "index": 3, "values": ["d"]}])["change"]
    c2 = docs[0].change([{"path": ["text"], "action": "insert", "index": 4, "values": ["e"]}])["change"]
    m = Micromerge("replica")
    with pytest.raises(RangeError, match="Missing dependency: change 1 by actor doc1"):
        m.applyChange(docs[1].change([{"path": ["text"], "action": "insert", "index": 0, "values": ["x"]}])["change"])
    m.applyChange(init)
    with pytest.raises(RangeError, match="Expected sequence number 2, got 3"):
        m.applyChange(c2)
    assert m.clock == {"doc1": 1}
    m.applyChange(c1); m.applyChange(c2)
    assert m.clock == {"doc1": 3}
    with pytest.raises(RangeError, match="Object does not exist"):
        m.applyChange({"actor": "doc9", "seq": 1, "deps": {}, "startOp": 50, "ops": [
            {"opId": "50@doc9", "action": "set", "obj": "7@nobody", "elemId": "_head", "insert": True, "value": "q"}]})


@pytest.mark.gpu
def test_kats_through_facade():
    for kat in [k for k in load_kats() if k["kind"] == "concurrent"]:
        rec = []
        run_concurrent(OracleMicromerge, kat, record=rec)
        for log in rec:
            m = Micromerge("replica")
            for ch in log:
                assert m.applyChange(ch) == []
            assert m.getTextWithFormatting(["text"]) == kat["expectedResult"]
            assert "".join(m.root["text"]) == "".join(s["text"] for s in kat["expectedResult"])
