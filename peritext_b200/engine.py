"""ctypes binding of the engine's C-ABI (include/peritext_b200.h, built as peritext_b200/libperitext_b200.so).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable, every entry point raises
``EngineError``.  torch is not needed here; pass ``torch.cuda.current_stream().cuda_stream`` as ``stream`` to enqueue
on a torch stream.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .packing import (CDESC_DT, CHANGE_DT, DEP_DT, DESC_DT, INSDEL_DT, MARK_DT, RESULT_DT, SPAN_DT, ChangeTable, MergedBatch, PackedBatch)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libperitext_b200.so")
_lib = None

EXPORTS = ["pt_batch_create", "pt_batch_upload", "pt_batch_upload_runs", "pt_compress_runs", "pt_compact_ops", "pt_batch_upload_compact", "pt_batch_adopt_device", "pt_batch_upload_changes",
           "pt_ingest_create", "pt_ingest_parse", "pt_ingest_packed", "pt_ingest_pool", "pt_ingest_error", "pt_ingest_destroy", "pt_batch_merge", "pt_batch_sync",
           "pt_batch_download", "pt_batch_download_begin", "pt_batch_download_results", "pt_batch_device_results", "pt_batch_launch_count", "pt_batch_stats",
           "pt_batch_last_merge_ms", "pt_batch_set_comment_pool", "pt_batch_download_patches", "pt_batch_set_patch_pool", "pt_batch_query_elements", "pt_batch_destroy", "pt_strerror", "pt_last_error", "pt_version"]


class EngineError(RuntimeError):
    pass


class _PackedOps(ctypes.Structure):
    _fields_ = [("n_logs", ctypes.c_uint32), ("logs", ctypes.c_void_p), ("insdel", ctypes.c_void_p),
                ("n_insdel_total", ctypes.c_uint64), ("marks", ctypes.c_void_p), ("n_mark_total", ctypes.c_uint64)]


class _PackedRuns(ctypes.Structure):
    _fields_ = [("n_logs", ctypes.c_uint32), ("logs", ctypes.c_void_p), ("run_off", ctypes.c_void_p), ("tok_off", ctypes.c_void_p),
                ("runs", ctypes.c_void_p), ("tokens", ctypes.c_void_p), ("marks", ctypes.c_void_p),
                ("n_insdel_total", ctypes.c_uint64), ("n_mark_total", ctypes.c_uint64)]


INSDEL_C8_DT = np.dtype([("ctr", "<u2"), ("ref_ctr", "<u2"), ("w", "<u4")])
MARK_C16_DT = np.dtype([("ctr", "<u2"), ("start_ctr", "<u2"), ("end_ctr", "<u2"), ("arrival", "<u2"), ("attr", "<u4"), ("w", "<u4")])
RUN_DT = np.dtype([("ctr0", "<u4"), ("ref_ctr", "<u4"), ("actor", "<u2"), ("ref_actor", "<u2"), ("kind_count", "<u4")])


class PackedRuns:
    """Run-compressed wire form of a PackedBatch (include/peritext_b200.h pt_packed_runs): typing runs and consecutive
    deletes collapse to one 16-byte run record (+ 4 bytes per inserted value)."""

    def __init__(self, desc, run_off, tok_off, runs, tokens, marks, n_insdel_total):
        self.desc, self.run_off, self.tok_off, self.runs, self.tokens, self.marks = desc, run_off, tok_off, runs, tokens, marks
        self.n_insdel_total = int(n_insdel_total)

    @property
    def n_logs(self) -> int:
        return int(self.desc.shape[0])

    @property
    def nbytes(self) -> int:
        return int(self.desc.nbytes + self.run_off.nbytes + self.tok_off.nbytes + self.runs.nbytes + self.tokens.nbytes + self.marks.nbytes)

    def slice_logs(self, a: int, b: int) -> "PackedRuns":
        """Logs [a, b) as views (pinned memory stays pinned); offsets re-based (small copies of the offset arrays)."""
        d = self.desc[a:b].copy()
        ro = self.run_off[a:b + 1].copy(); to = self.tok_off[a:b + 1].copy()
        r0, r1, t0, t1 = int(ro[0]), int(ro[-1]), int(to[0]), int(to[-1])
        if len(d):
            i0, m0 = int(d[0]["insdel_off"]), int(d[0]["mark_off"])
            m1 = int(d[-1]["mark_off"]) + int(d[-1]["n_mark"]); i1 = int(d[-1]["insdel_off"]) + int(d[-1]["n_insdel"])
            d["insdel_off"] -= i0; d["mark_off"] -= m0
        else:
            i0 = i1 = m0 = m1 = 0
        return PackedRuns(d, ro - ro[0], to - to[0], self.runs[r0:r1], self.tokens[t0:t1], self.marks[m0:m1], i1 - i0)


def compress_runs(batch: PackedBatch, pin=None) -> PackedRuns:
    """Host-side run compression (pt_compress_runs).  `pin(nbytes_array) -> array` may place the big arrays in pinned memory."""
    L = load_library()
    desc = np.ascontiguousarray(batch.desc)
    insdel = np.ascontiguousarray(batch.insdel)
    marks = np.ascontiguousarray(batch.marks)
    ops = _PackedOps(len(desc), desc.ctypes.data, insdel.ctypes.data, len(insdel), marks.ctypes.data, len(marks))
    n = len(desc)
    run_off = np.zeros(n + 1, np.uint64); tok_off = np.zeros(n + 1, np.uint64)
    nr, nt = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _check(L.pt_compress_runs(ctypes.byref(ops), run_off.ctypes.data, tok_off.ctypes.data, None, None, ctypes.byref(nr), ctypes.byref(nt)), "pt_compress_runs")
    alloc = pin or (lambda a: a)
    runs = alloc(np.zeros(max(1, nr.value), RUN_DT)); tokens = alloc(np.zeros(max(1, nt.value), np.uint32))
    _check(L.pt_compress_runs(ctypes.byref(ops), run_off.ctypes.data, tok_off.ctypes.data, runs.ctypes.data, tokens.ctypes.data, ctypes.byref(nr), ctypes.byref(nt)), "pt_compress_runs")
    return PackedRuns(desc, alloc(run_off), alloc(tok_off), runs[: nr.value], tokens[: nt.value], alloc(marks) if pin else marks, len(insdel))


class _ChangeTable(ctypes.Structure):
    _fields_ = [("n_logs", ctypes.c_uint32), ("logs", ctypes.c_void_p), ("changes", ctypes.c_void_p), ("n_changes_total", ctypes.c_uint64),
                ("deps", ctypes.c_void_p), ("n_deps_total", ctypes.c_uint64)]


class _SpansView(ctypes.Structure):
    _fields_ = [("n_logs", ctypes.c_uint32), ("results", ctypes.c_void_p), ("text_off", ctypes.c_void_p),
                ("span_off", ctypes.c_void_p), ("text", ctypes.c_void_p), ("spans", ctypes.c_void_p),
                ("comment_pool", ctypes.c_void_p), ("comment_pool_used", ctypes.c_uint64), ("seq", ctypes.c_void_p),
                ("seq_off", ctypes.c_void_p), ("comment_pool_needed", ctypes.c_uint64)]


class _Limits(ctypes.Structure):
    _fields_ = [("comment_pool_entries", ctypes.c_uint64), ("flags", ctypes.c_uint32), ("patch_pool_items", ctypes.c_uint32),
                ("reserved", ctypes.c_uint32 * 4)]


class _PatchView(ctypes.Structure):
    _fields_ = [("recs", ctypes.c_void_p), ("items", ctypes.c_void_p), ("n_items", ctypes.c_uint64), ("n_items_needed", ctypes.c_uint64),
                ("status", ctypes.c_void_p)]


QUERY_DT = np.dtype([("log", "<u4"), ("index", "<u4"), ("flags", "<u4"), ("reserved", "<u4")])
PATCH_REC_DT = np.dtype([("index", "<u4"), ("flags", "<u4"), ("link_attr", "<u4"), ("reserved", "<u4")])
PATCH_ITEM_DT = np.dtype([("log", "<u4"), ("tag", "<u4"), ("a", "<u4"), ("b", "<u4")])
FLAG_EMIT_SEQUENCE = 1
FLAG_EMIT_PATCHES = 2


def load_library() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"(make -C peritext_b200/csrc). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32, u64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64
    L.pt_batch_create.argtypes = [ctypes.c_int, vp, vp, ctypes.POINTER(vp)]
    L.pt_batch_upload.argtypes = [vp, vp]
    L.pt_batch_adopt_device.argtypes = [vp, vp]
    L.pt_batch_upload_runs.argtypes = [vp, vp]
    L.pt_compress_runs.argtypes = [vp, vp, vp, vp, vp, ctypes.POINTER(u64), ctypes.POINTER(u64)]
    L.pt_batch_upload_changes.argtypes = [vp, vp]
    L.pt_compact_ops.argtypes = [vp, vp, vp, ctypes.c_int]
    L.pt_batch_upload_compact.argtypes = [vp, vp]
    L.pt_ingest_create.argtypes = [ctypes.POINTER(vp)]
    L.pt_ingest_parse.argtypes = [vp, vp, vp, u32, ctypes.c_int]
    L.pt_ingest_packed.argtypes = [vp, vp, vp]
    L.pt_ingest_pool.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(u64), ctypes.POINTER(vp)]
    L.pt_ingest_error.argtypes = [vp]; L.pt_ingest_error.restype = ctypes.c_char_p
    L.pt_ingest_destroy.argtypes = [vp]; L.pt_ingest_destroy.restype = None
    L.pt_batch_merge.argtypes = [vp]
    L.pt_batch_sync.argtypes = [vp]
    L.pt_batch_download.argtypes = [vp, vp]
    L.pt_batch_download_begin.argtypes = [vp]
    L.pt_batch_download_results.argtypes = [vp, vp, u32]
    L.pt_batch_device_results.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(u32)]
    L.pt_batch_launch_count.argtypes = [vp]; L.pt_batch_launch_count.restype = u64
    L.pt_batch_stats.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64 * 4)]
    L.pt_batch_last_merge_ms.argtypes = [vp]; L.pt_batch_last_merge_ms.restype = ctypes.c_float
    L.pt_batch_set_comment_pool.argtypes = [vp, u64]
    L.pt_batch_download_patches.argtypes = [vp, vp]
    L.pt_batch_set_patch_pool.argtypes = [vp, u64]
    L.pt_batch_query_elements.argtypes = [vp, vp, u32, vp]
    L.pt_batch_destroy.argtypes = [vp]; L.pt_batch_destroy.restype = None
    L.pt_strerror.argtypes = [ctypes.c_int]; L.pt_strerror.restype = ctypes.c_char_p
    L.pt_last_error.restype = ctypes.c_char_p
    L.pt_version.restype = ctypes.c_char_p
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        L = load_library()
        raise EngineError(f"{what}: {L.pt_strerror(rc).decode()} ({L.pt_last_error().decode()})")


class BatchEngine:
    """One handle per (GPU, batch).  ``upload`` -> ``merge`` -> ``download``."""

    def __init__(self, device: int = 0, stream: int | None = None, comment_pool_entries: int = 0, emit_sequence: bool = False,
                 emit_patches: bool = False):
        L = load_library()
        self._L = L
        self._h = ctypes.c_void_p()
        self.emit_patches = emit_patches
        flags = (FLAG_EMIT_SEQUENCE if (emit_sequence or emit_patches) else 0) | (FLAG_EMIT_PATCHES if emit_patches else 0)
        lim = _Limits(comment_pool_entries, flags, 0, (ctypes.c_uint32 * 4)())
        _check(L.pt_batch_create(device, ctypes.byref(lim), ctypes.c_void_p(stream or 0), ctypes.byref(self._h)), "pt_batch_create")
        self._keep = None
        self.n_logs = 0

    def _ops_struct(self, desc, insdel_ptr, n_insdel, marks_ptr, n_mark):
        return _PackedOps(len(desc), desc.ctypes.data, insdel_ptr, n_insdel, marks_ptr, n_mark)

    def upload(self, batch: PackedBatch):
        desc = np.ascontiguousarray(batch.desc)
        insdel = np.ascontiguousarray(batch.insdel)
        marks = np.ascontiguousarray(batch.marks)
        ops = self._ops_struct(desc, insdel.ctypes.data, len(insdel), marks.ctypes.data, len(marks))
        _check(self._L.pt_batch_upload(self._h, ctypes.byref(ops)), "pt_batch_upload")
        self.n_logs = len(desc)
        self._n_insdel = len(insdel)

    def upload_changes(self, table: ChangeTable):
        """Attach the batch's change table: the next merge runs the admission pre-pass (seq / deps checks of
        Micromerge.applyChange, reference src/micromerge.ts:501-509) and rejected logs report status 6 / 7."""
        d, c, p = np.ascontiguousarray(table.desc), np.ascontiguousarray(table.changes), np.ascontiguousarray(table.deps)
        t = _ChangeTable(len(d), d.ctypes.data, c.ctypes.data if len(c) else 0, len(c), p.ctypes.data if len(p) else 0, len(p))
        _check(self._L.pt_batch_upload_changes(self._h, ctypes.byref(t)), "pt_batch_upload_changes")

    def upload_compact(self, batch: PackedBatch, cins: np.ndarray | None = None, cmarks: np.ndarray | None = None, threads: int = 0):
        """Upload in the compact wire format (8-byte ins/del, 16-byte mark records; expanded on the device): the conversion
        (pt_compact_ops, multithreaded) writes into `cins` / `cmarks` (INSDEL_C8_DT / MARK_C16_DT arrays, ideally pinned)."""
        desc = np.ascontiguousarray(batch.desc)
        insdel = np.ascontiguousarray(batch.insdel); marks = np.ascontiguousarray(batch.marks)
        if cins is None:
            cins = np.zeros(max(1, len(insdel)), INSDEL_C8_DT)
        if cmarks is None:
            cmarks = np.zeros(max(1, len(marks)), MARK_C16_DT)
        ops = self._ops_struct(desc, insdel.ctypes.data, len(insdel), marks.ctypes.data, len(marks))
        _check(self._L.pt_compact_ops(ctypes.byref(ops), cins.ctypes.data, cmarks.ctypes.data, threads), "pt_compact_ops")
        cc = _PackedOps(len(desc), desc.ctypes.data, cins.ctypes.data, len(insdel), cmarks.ctypes.data, len(marks))     # same field layout as pt_packed_compact
        self._keep = (desc, cins, cmarks)
        _check(self._L.pt_batch_upload_compact(self._h, ctypes.byref(cc)), "pt_batch_upload_compact")
        self.n_logs = len(desc)
        self._n_insdel = len(insdel)

    def upload_runs(self, r: PackedRuns):
        desc = np.ascontiguousarray(r.desc)
        st = _PackedRuns(len(desc), desc.ctypes.data, r.run_off.ctypes.data, r.tok_off.ctypes.data, r.runs.ctypes.data if len(r.runs) else 0,
                         r.tokens.ctypes.data if len(r.tokens) else 0, r.marks.ctypes.data if len(r.marks) else 0, r.n_insdel_total, len(r.marks))
        self._keep = (desc, r)
        _check(self._L.pt_batch_upload_runs(self._h, ctypes.byref(st)), "pt_batch_upload_runs")
        self.n_logs = len(desc)

    def adopt_device(self, desc: np.ndarray, insdel_dev_ptr: int, n_insdel: int, marks_dev_ptr: int, n_mark: int):
        """Use op arrays already resident in device memory (e.g. ``tensor.data_ptr()``); caller keeps them alive."""
        desc = np.ascontiguousarray(desc)
        ops = self._ops_struct(desc, insdel_dev_ptr, n_insdel, marks_dev_ptr, n_mark)
        _check(self._L.pt_batch_adopt_device(self._h, ctypes.byref(ops)), "pt_batch_adopt_device")
        self.n_logs = len(desc)

    def merge(self):
        _check(self._L.pt_batch_merge(self._h), "pt_batch_merge")

    def sync(self):
        _check(self._L.pt_batch_sync(self._h), "pt_batch_sync")

    @property
    def last_merge_ms(self) -> float:
        return float(self._L.pt_batch_last_merge_ms(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.pt_batch_launch_count(self._h))

    def stats(self) -> dict:
        out = (ctypes.c_uint64 * 4)()
        _check(self._L.pt_batch_stats(self._h, ctypes.byref(out)), "pt_batch_stats")
        return {"logs_shared_only": int(out[0]), "logs_spill_path": int(out[1]), "logs_deferred_to_big_bin": int(out[2]),
                "comment_pool_needed": int(out[3])}

    def device_results_ptr(self) -> int:
        p = ctypes.c_void_p(); n = ctypes.c_uint32()
        _check(self._L.pt_batch_device_results(self._h, ctypes.byref(p), ctypes.byref(n)), "pt_batch_device_results")
        return int(p.value or 0)

    def results(self) -> np.ndarray:
        out = np.zeros(self.n_logs, RESULT_DT)
        _check(self._L.pt_batch_download_results(self._h, out.ctypes.data, self.n_logs), "pt_batch_download_results")
        return out

    def download_begin(self):
        _check(self._L.pt_batch_download_begin(self._h), "pt_batch_download_begin")

    def download(self, copy: bool = True) -> MergedBatch:
        v = _SpansView()
        _check(self._L.pt_batch_download(self._h, ctypes.byref(v)), "pt_batch_download")
        n = v.n_logs

        def arr(ptr, count, dt):
            if not count:
                return np.zeros(0, dt)
            buf = (ctypes.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr)
            a = np.frombuffer(buf, dtype=dt, count=count)
            return a.copy() if copy else a

        results = arr(v.results, n, RESULT_DT)
        text_off = arr(v.text_off, n + 1, np.uint64)      # packed on the device: offsets are the scan of the counts
        span_off = arr(v.span_off, n + 1, np.uint64)
        n_text, n_span = int(text_off[-1]), int(span_off[-1])
        seq_off = arr(v.seq_off, n, np.uint64) if (n and v.seq) else None
        n_seq = (int(seq_off[-1]) + int(results[-1]["n_elems"])) if seq_off is not None else 0
        self.comment_pool_needed = int(v.comment_pool_needed)
        self.comment_pool_used = int(v.comment_pool_used)
        return MergedBatch(results, text_off, span_off, arr(v.text, n_text, np.uint32), arr(v.spans, n_span, SPAN_DT),
                           arr(v.comment_pool, int(v.comment_pool_used), np.uint32),
                           arr(v.seq, n_seq, np.uint32) if v.seq else None, seq_off)

    def download_patches(self):
        """PT_FLAG_EMIT_PATCHES: (patch records per ins/del record, pool items, per-log status, items needed) of the last merge."""
        v = _PatchView()
        _check(self._L.pt_batch_download_patches(self._h, ctypes.byref(v)), "pt_batch_download_patches")

        def arr(ptr, count, dt):
            if not count or not ptr:
                return np.zeros(0, dt)
            buf = (ctypes.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt, count=count).copy()
        return arr(v.recs, self._n_insdel, PATCH_REC_DT), arr(v.items, int(v.n_items), PATCH_ITEM_DT), arr(v.status, self.n_logs, np.uint32), int(v.n_items_needed)

    def run_with_patches(self, batch: PackedBatch):
        """upload -> merge (+ device Patch stream) -> download; returns (MergedBatch, DevicePatches)."""
        out = self.run(batch)
        recs, items, status, needed = self.download_patches()
        if needed > len(items):
            _check(self._L.pt_batch_set_patch_pool(self._h, needed + 16), "pt_batch_set_patch_pool")
            self.merge(); out = self.download()
            recs, items, status, needed = self.download_patches()
        from .packing import DevicePatches
        return out, DevicePatches(recs, items, status)

    def query_elements(self, logs, indices, look_after_tombstones=False) -> np.ndarray:
        """Batched getListElementId on the device (reference src/micromerge.ts:762-805): for query k the index of the insert
        record of log `logs[k]`'s `indices[k]`-th visible element (with `look_after_tombstones`: moved to the last following
        tombstone whose after-slot is defined, the rule `change()` uses for insert positions); 0xFFFFFFFF = out of bounds.
        Needs emit_sequence and a completed merge."""
        q = np.zeros(len(logs), QUERY_DT)
        q["log"], q["index"] = logs, indices
        q["flags"] = np.asarray(look_after_tombstones, dtype=np.uint32) if not np.isscalar(look_after_tombstones) else (1 if look_after_tombstones else 0)
        out = np.zeros(len(q), np.uint32)
        _check(self._L.pt_batch_query_elements(self._h, q.ctypes.data, len(q), out.ctypes.data), "pt_batch_query_elements")
        return out

    def set_comment_pool(self, entries: int):
        _check(self._L.pt_batch_set_comment_pool(self._h, int(entries)), "pt_batch_set_comment_pool")

    def run(self, batch: PackedBatch) -> MergedBatch:
        """upload -> merge -> download.  The comment pool has a default capacity; a log whose comment lists do not fit
        reports status 4 without consuming pool space and the engine reports the batch's exact demand, so one re-merge
        with a pool of that size always succeeds (documents with many overlapping comments are valid input)."""
        self.upload(batch)
        if getattr(batch, "changes", None) is not None:
            self.upload_changes(batch.changes)
        self.merge(); out = self.download()
        if len(out.results) and (out.results["status"] == 4).any() and self.comment_pool_needed > self.comment_pool_used:
            self.set_comment_pool(self.comment_pool_needed + 16)
            self.merge(); out = self.download()
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.pt_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PipelinedEngine:
    """Host <-> device pipelining for one big batch: the batch is cut into `chunks` runs of logs, each with its own
    engine handle and CUDA stream, so chunk k+1's upload overlaps chunk k's merge and chunk k-1's download (separate
    copy engines).  Inputs should live in pinned host memory (otherwise the uploads synchronise)."""

    def __init__(self, device: int = 0, chunks: int = 4, streams=None):
        self.chunks = chunks
        self._streams = streams
        if streams is None:
            import torch
            self._streams = [torch.cuda.Stream(device=device) for _ in range(chunks)]
        self.engines = [BatchEngine(device, stream=s.cuda_stream) for s in self._streams]

    def _compact_buffers(self, k: int, n_ins: int, n_mk: int):
        """Pinned conversion targets of chunk k (allocated once, grown on demand)."""
        import torch
        if not hasattr(self, "_cbuf"):
            self._cbuf = {}
        cur = self._cbuf.get(k)
        if cur is None or cur[0].numel() < n_ins * 8 or cur[1].numel() < n_mk * 16:
            cur = (torch.empty(max(16, n_ins * 8 + n_ins), dtype=torch.uint8).pin_memory(), torch.empty(max(16, n_mk * 16 + n_mk), dtype=torch.uint8).pin_memory())
            self._cbuf[k] = cur
        return cur[0].numpy()[: n_ins * 8].view(INSDEL_C8_DT), cur[1].numpy()[: n_mk * 16].view(MARK_C16_DT)

    def run(self, batch, copy: bool = False, compact: bool = False, threads: int = 0) -> list[MergedBatch]:
        """`batch`: a PackedBatch, or a PackedRuns (run-compressed upload).  `compact`: convert every chunk to the compact wire
        format on the host (multithreaded, inside this call) and upload half the bytes; chunk k+1's conversion overlaps
        chunk k's transfer."""
        n = batch.n_logs
        # cut by records, not by log count, so the chunks carry similar work
        w = np.cumsum(batch.desc["n_insdel"].astype(np.int64) + 2 * batch.desc["n_mark"].astype(np.int64))
        cuts = [0] + [int(np.searchsorted(w, w[-1] * (k + 1) / self.chunks, side="left")) + 1 for k in range(self.chunks - 1)] + [n] if n else [0, 0]
        cuts = sorted(set(min(max(c, 0), n) for c in cuts))
        subs = [batch.slice_logs(a, b) for a, b in zip(cuts, cuts[1:]) if b > a]
        used = self.engines[: len(subs)]
        for e, sb in zip(used, subs):
            if isinstance(sb, PackedRuns):
                e.upload_runs(sb)
            elif compact:
                ci, cm = self._compact_buffers(used.index(e), len(sb.insdel), len(sb.marks))
                e.upload_compact(sb, ci, cm, threads)
            else:
                e.upload(sb)
            e.merge(); e.download_begin()
        return [e.download(copy=copy) for e in used]

    def close(self):
        for e in self.engines:
            e.close()


def pack_logs_native(logs_json, threads: int = 0) -> PackedBatch:
    """Native (C++, multithreaded) wire-format ingest: ``logs_json[i]`` = JSON text (str or bytes) of the Change objects
    replica i applied, in arrival order -> PackedBatch with its change table (csrc/ingest.cpp, pt_ingest_*).  Packs exactly
    like ``packing.pack_logs(..., with_changes=True)``."""
    import json
    L = load_library()
    blobs = [s.encode("utf-8", "surrogatepass") if isinstance(s, str) else bytes(s) for s in logs_json]
    n = len(blobs)
    ptrs = (ctypes.c_char_p * max(1, n))(*blobs) if n else (ctypes.c_char_p * 1)()
    lens = (ctypes.c_uint64 * max(1, n))(*[len(b) for b in blobs]) if n else (ctypes.c_uint64 * 1)()
    h = ctypes.c_void_p()
    _check(L.pt_ingest_create(ctypes.byref(h)), "pt_ingest_create")
    try:
        rc = L.pt_ingest_parse(h, ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(lens, ctypes.c_void_p), n, threads)
        if rc != 0:
            raise ValueError(L.pt_ingest_error(h).decode("utf-8", "replace"))
        ops, tab = _PackedOps(), _ChangeTable()
        _check(L.pt_ingest_packed(h, ctypes.byref(ops), ctypes.byref(tab)), "pt_ingest_packed")

        def arr(ptr, count, dt):
            if not count or not ptr:
                return np.zeros(0, dt)
            buf = (ctypes.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dt, count=count).copy()

        def pool(kind):
            data, off, cnt, first = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_void_p()
            _check(L.pt_ingest_pool(h, kind, ctypes.byref(data), ctypes.byref(off), ctypes.byref(cnt), ctypes.byref(first)), "pt_ingest_pool")
            o = arr(off.value, cnt.value + 1, np.uint64)
            raw = bytes(arr(data.value, int(o[-1]), np.uint8)) if cnt.value else b""
            items = [raw[int(o[k]): int(o[k + 1])] for k in range(cnt.value)]
            f = arr(first.value, n + 1, np.uint64) if first.value else None
            return items, f

        u16 = lambda b: b.decode("utf-16-le", "surrogatepass")
        values = [u16(b) for b in pool(0)[0]]
        link_attrs = [json.loads(b.decode("utf-8", "surrogatepass")) for b in pool(1)[0]]
        comment_attrs = [json.loads(b.decode("utf-8", "surrogatepass")) for b in pool(3)[0]]
        actors, afirst = pool(4)
        counters, cfirst = pool(5)
        log_actors = [[u16(x) for x in actors[int(afirst[i]): int(afirst[i + 1])]] for i in range(n)]
        log_counters = []
        for i in range(n):
            c = counters[int(cfirst[i]): int(cfirst[i + 1])]
            log_counters.append(np.array([int.from_bytes(x, "little") for x in c], dtype=np.uint64) if c else None)
        table = ChangeTable(arr(tab.logs, tab.n_logs, CDESC_DT), arr(tab.changes, tab.n_changes_total, CHANGE_DT), arr(tab.deps, tab.n_deps_total, DEP_DT))
        return PackedBatch(arr(ops.logs, ops.n_logs, DESC_DT), arr(ops.insdel, ops.n_insdel_total, INSDEL_DT), arr(ops.marks, ops.n_mark_total, MARK_DT),
                           values, link_attrs, comment_attrs, [], log_actors=log_actors, log_counters=log_counters, changes=table)
    finally:
        L.pt_ingest_destroy(h)
