"""Seeded synthetic workloads for the BASELINE.json configs — binding of csrc/workload.cpp.

``generate("c2", n_docs=...)`` returns a ``PackedBatch`` whose log ``d * R + r`` is what replica ``r`` of document ``d``
applied, in its own arrival order.  Comment attrs are synthetic integers (``doc * 4096 + k``), already in rank order;
link attrs are url ids 0..25 (``A.com``..``Z.com``, reference test/fuzz.ts:28).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .packing import DESC_DT, INSDEL_DT, MARK_DT, PackedBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpt_workload.so")
_lib = None

# name -> (kind, default docs, ops per doc, replicas, n_marks)   (BASELINE.json configs[1..4])
CONFIGS = {
    "c2": dict(kind=2, n_docs=1000, ops_per_doc=10000, replicas=2, n_marks=0,
               label="1K docs x 10K ops, insert/delete only, 2 replicas"),
    "c3": dict(kind=3, n_docs=1000, ops_per_doc=10000, replicas=2, n_marks=0,
               label="1K docs x 10K ops with bold/italic/link/comment marks, 2 replicas"),
    "c4": dict(kind=4, n_docs=100000, ops_per_doc=1000, replicas=3, n_marks=0,
               label="100K docs x 1K ops fuzz-generated, 3 concurrent replicas"),
    "c5": dict(kind=5, n_docs=10000, ops_per_doc=100000, replicas=2, n_marks=10000,
               label="10K docs x 100K-char long-form, dense overlapping marks"),
}


class _Config(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("n_docs", ctypes.c_uint32), ("doc_first", ctypes.c_uint32),
                ("ops_per_doc", ctypes.c_uint32), ("replicas", ctypes.c_uint32), ("n_marks", ctypes.c_uint32),
                ("seed", ctypes.c_uint64), ("threads", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class _Batch(ctypes.Structure):
    _fields_ = [("n_logs", ctypes.c_uint32), ("desc", ctypes.c_void_p), ("insdel", ctypes.c_void_p),
                ("n_insdel", ctypes.c_uint64), ("marks", ctypes.c_void_p), ("n_marks", ctypes.c_uint64),
                ("unique_ops", ctypes.c_uint64)]


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run __graft_entry__.build()")
        L = ctypes.CDLL(LIB_PATH)
        L.ptw_generate.argtypes = [ctypes.POINTER(_Config), ctypes.POINTER(ctypes.POINTER(_Batch))]
        L.ptw_generate.restype = ctypes.c_int
        L.ptw_free.argtypes = [ctypes.POINTER(_Batch)]
        _lib = L
    return _lib


class _SyntheticComments:
    def __getitem__(self, i):
        return {"id": "comment-%010d" % int(i)}


def generate(config: str, *, n_docs: int | None = None, ops_per_doc: int | None = None, replicas: int | None = None,
             n_marks: int | None = None, doc_first: int = 0, seed: int | None = None, threads: int = 0) -> PackedBatch:
    cfg = dict(CONFIGS[config])
    if n_docs is not None: cfg["n_docs"] = n_docs
    if ops_per_doc is not None: cfg["ops_per_doc"] = ops_per_doc
    if replicas is not None: cfg["replicas"] = replicas
    if n_marks is not None: cfg["n_marks"] = n_marks
    seed = (0x5EED0000 + cfg["kind"]) if seed is None else seed
    L = _load()
    c = _Config(cfg["kind"], cfg["n_docs"], doc_first, cfg["ops_per_doc"], cfg["replicas"], cfg["n_marks"], seed, threads, 0)
    out = ctypes.POINTER(_Batch)()
    rc = L.ptw_generate(ctypes.byref(c), ctypes.byref(out))
    if rc != 0:
        raise RuntimeError(f"ptw_generate failed: {rc}")
    b = out.contents

    def arr(ptr, count, dt):
        if not count:
            return np.zeros(0, dt)
        buf = (ctypes.c_char * (count * dt.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt, count=count).copy()

    batch = PackedBatch(arr(b.desc, b.n_logs, DESC_DT), arr(b.insdel, b.n_insdel, INSDEL_DT), arr(b.marks, b.n_marks, MARK_DT),
                        values=[], link_attrs=[{"url": f"{ch}.com"} for ch in "ABCDEFGHIJKLMNOPQRSTUVWXYZ"],
                        comment_ids=_SyntheticComments(), other_attrs=[],
                        meta=dict(config=config, label=cfg["label"], n_docs=cfg["n_docs"], replicas=cfg["replicas"],
                                  ops_per_doc=cfg["ops_per_doc"], unique_ops=int(b.unique_ops), seed=seed, doc_first=doc_first))
    L.ptw_free(out)
    return batch
