"""peritext_b200 — B200-native batch CRDT-merge engine for Peritext's op-log apply + flatten hot path.

Public surface:
  * ``Micromerge``      — facade with the reference class surface (reference src/micromerge.ts:262)
  * ``BatchEngine``     — the batch entry (many logs per launch) over the C-ABI in include/peritext_b200.h
  * ``pack_logs`` / ``decode_spans`` — wire-format ingest / result decode
  * ``workload``        — seeded synthetic trace generator (BASELINE.json configs)
"""
from .packing import PackedBatch, MergedBatch, pack_logs, decode_spans, RangeError  # noqa: F401


def __getattr__(name):
    if name in ("BatchEngine", "EngineError", "load_library"):
        from . import engine
        return getattr(engine, name)
    if name == "Micromerge":
        from .micromerge import Micromerge
        return Micromerge
    raise AttributeError(name)
