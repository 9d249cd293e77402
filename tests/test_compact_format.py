"""The compact wire format (pt_compact_ops: 8-byte ins/del and 16-byte mark records, include/peritext_b200.h): the host
converter is elementwise and lossless for representable logs — checked on CPU by expanding in numpy exactly like the device
kernels do — and refuses logs it cannot represent."""
import ctypes

import numpy as np
import pytest

from peritext_b200 import workload
from peritext_b200.engine import INSDEL_C8_DT, MARK_C16_DT, EngineError, _PackedOps, _check, load_library
from peritext_b200.packing import INSDEL_DT, MARK_DT


def convert(batch, threads=3):
    L = load_library()
    desc = np.ascontiguousarray(batch.desc); ins = np.ascontiguousarray(batch.insdel); mk = np.ascontiguousarray(batch.marks)
    ci = np.zeros(max(1, len(ins)), INSDEL_C8_DT); cm = np.zeros(max(1, len(mk)), MARK_C16_DT)
    ops = _PackedOps(len(desc), desc.ctypes.data, ins.ctypes.data, len(ins), mk.ctypes.data, len(mk))
    _check(L.pt_compact_ops(ctypes.byref(ops), ci.ctypes.data, cm.ctypes.data, threads), "pt_compact_ops")
    return ci[: len(ins)], cm[: len(mk)]


def expand(ci, cm):
    ins = np.zeros(len(ci), INSDEL_DT); mk = np.zeros(len(cm), MARK_DT)
    w = ci["w"]; tok = w >> 10
    ins["ctr"], ins["ref_ctr"], ins["actor"], ins["ref_actor"] = ci["ctr"], ci["ref_ctr"], w & 0xF, (w >> 4) & 0xF
    ins["payload"] = (((w >> 8) & 3) << 30) | np.where(tok & 0x200000, 0x20000000, 0) | (tok & 0x1FFFFF)
    w = cm["w"]
    mk["ctr"], mk["start_ctr"], mk["end_ctr"], mk["arrival"], mk["attr"] = cm["ctr"], cm["start_ctr"], cm["end_ctr"], cm["arrival"], cm["attr"]
    mk["actor"], mk["start_actor"], mk["end_actor"], mk["kind"], mk["bounds"] = w & 0xF, (w >> 4) & 0xF, (w >> 8) & 0xF, (w >> 12) & 7, (w >> 15) & 0xF
    return ins, mk


@pytest.mark.parametrize("cfg,n_docs,ops", [("c4", 40, 1000), ("c3", 4, 4000), ("c2", 4, 4000)])
def test_round_trip(cfg, n_docs, ops):
    b = workload.generate(cfg, n_docs=n_docs, ops_per_doc=ops)
    ins, mk = expand(*convert(b))
    assert ins.tobytes() == b.insdel.tobytes()
    assert mk.tobytes() == b.marks.tobytes()


def test_pooled_values_and_unicode_tokens():
    from oracle.oracle import Micromerge
    from peritext_b200.packing import pack_logs
    from tests.harness import generateDocs
    docs, _, init = generateDocs(Micromerge, "ab", 1)
    c1 = docs[0].change([{"path": ["text"], "action": "insert", "index": 1, "values": [" is great!", "é", "\\U0001F600", "中"]}])["change"]
    b = pack_logs([[init, c1]])
    ins, mk = expand(*convert(b))
    assert ins.tobytes() == b.insdel.tobytes()


def test_unrepresentable_logs_are_refused():
    b = workload.generate("c2", n_docs=1, ops_per_doc=2000)
    b.desc["max_ctr"][0] = 70000
    with pytest.raises(EngineError):
        convert(b)
