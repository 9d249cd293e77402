"""TEST INFRASTRUCTURE — drives the oracle's packed replay (po_replay_packed in oracle/peritext_oracle.cpp):
the reference's sequential algorithm applied to packed logs, emitting the engine's binary result format so that
bulk parity checks are exact array comparisons.  Also the timed body of bench.py's cpu_baseline leg."""
from __future__ import annotations

import ctypes
import time

import numpy as np

from peritext_b200.packing import (RESULT_DT, SPAN_DT, MergedBatch, PackedBatch, comment_pool_capacity, output_layout)

from .oracle import lib


class _PackedOps(ctypes.Structure):
    _fields_ = [("n_logs", ctypes.c_uint32), ("logs", ctypes.c_void_p), ("insdel", ctypes.c_void_p),
                ("n_insdel_total", ctypes.c_uint64), ("marks", ctypes.c_void_p), ("n_mark_total", ctypes.c_uint64)]


def replay_packed(batch: PackedBatch, *, first: int = 0, count: int | None = None, threads: int = 1,
                  flatten: bool = True) -> tuple[MergedBatch, float]:
    """Returns (results, seconds spent inside the replay)."""
    L = lib()
    L.po_replay_packed.restype = ctypes.c_int
    L.po_replay_packed.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.c_uint64, ctypes.c_void_p]
    n = batch.n_logs
    count = n - first if count is None else count
    desc = np.ascontiguousarray(batch.desc)
    insdel = np.ascontiguousarray(batch.insdel)
    marks = np.ascontiguousarray(batch.marks)
    text_off, span_off, n_text, n_span = output_layout(desc)
    results = np.zeros(n, RESULT_DT)
    text = np.zeros(max(n_text, 1), np.uint32)
    spans = np.zeros(max(n_span, 1), SPAN_DT)
    cap = comment_pool_capacity(batch)
    pool = np.zeros(cap, np.uint32)
    used = ctypes.c_uint64(0)
    ops = _PackedOps(n, desc.ctypes.data, insdel.ctypes.data, len(insdel), marks.ctypes.data, len(marks))
    t0 = time.perf_counter()
    rc = L.po_replay_packed(ctypes.byref(ops), first, count, threads, 1 if flatten else 0, results.ctypes.data,
                            text_off.ctypes.data, span_off.ctypes.data, text.ctypes.data, spans.ctypes.data,
                            pool.ctypes.data, cap, ctypes.byref(used))
    dt = time.perf_counter() - t0
    assert rc == 0
    return MergedBatch(results, text_off, span_off, text, spans, pool[: used.value].copy()), dt
