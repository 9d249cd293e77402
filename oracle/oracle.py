"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (oracle/peritext_oracle.cpp).

``Micromerge`` here mirrors the reference class surface (reference src/micromerge.ts:262) so the parity tests
read like the reference's own tests: ``change`` (:308), ``applyChange`` (:499), ``getTextWithFormatting``
(:516), ``getCursor``/``resolveCursor`` (:465/:475), ``root`` (:290), ``clock`` (:273).

Parity pin: tests/golden/kats.json (transcribed from the reference's test/micromerge.ts).
"""
from __future__ import annotations

import ctypes
import json
import os
import subprocess
from typing import Any

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libperitext_oracle.so")
_lib = None


class RangeError(Exception):
    """JS RangeError thrown by the reference (e.g. src/micromerge.ts:503,507,752,804)."""


class JsError(Exception):
    """JS Error thrown by the reference."""


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "peritext_oracle.cpp")
    deps = [src, os.path.join(_HERE, "json_min.hpp"),
            os.path.join(_HERE, "..", "include", "peritext_b200.h"),
            os.path.join(_HERE, "..", "include", "pt_digest.h")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.exists(d) and os.path.getmtime(d) > os.path.getmtime(_LIB_PATH) for d in deps)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libperitext_oracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.po_doc_new.restype = ctypes.c_void_p
        L.po_doc_new.argtypes = [ctypes.c_char_p]
        L.po_doc_free.argtypes = [ctypes.c_void_p]
        for name, args in [("po_doc_change", [ctypes.c_void_p, ctypes.c_char_p]),
                           ("po_doc_apply_change", [ctypes.c_void_p, ctypes.c_char_p]),
                           ("po_doc_spans", [ctypes.c_void_p]), ("po_doc_root", [ctypes.c_void_p]),
                           ("po_doc_clock", [ctypes.c_void_p]), ("po_doc_elements", [ctypes.c_void_p]),
                           ("po_doc_get_cursor", [ctypes.c_void_p, ctypes.c_int64]),
                           ("po_doc_resolve_cursor", [ctypes.c_void_p, ctypes.c_char_p])]:
            fn = getattr(L, name)
            fn.restype = ctypes.c_void_p
            fn.argtypes = args
        L.po_free.argtypes = [ctypes.c_void_p]
        L.po_compare_op_ids.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        L.po_compare_op_ids.restype = ctypes.c_int
        L.po_version.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _take(ptr) -> str:
    s = ctypes.string_at(ptr).decode("utf-8")
    lib().po_free(ptr)
    if s.startswith("!"):
        kind, _, msg = s[1:].partition(":")
        raise (RangeError if kind == "RangeError" else JsError)(msg)
    return s


def compareOpIds(a: str, b: str) -> int:
    """reference src/micromerge.ts:812-827"""
    return lib().po_compare_op_ids(a.encode(), b.encode())


class Micromerge:
    contentKey = "text"  # src/micromerge.ts:264

    def __init__(self, actorId: str):
        self.actorId = actorId
        self._h = lib().po_doc_new(actorId.encode("utf-8"))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.po_doc_free(self._h)
            self._h = None

    def change(self, ops: list[dict]) -> dict[str, Any]:
        return json.loads(_take(lib().po_doc_change(self._h, json.dumps(ops).encode("utf-8"))))

    def applyChange(self, change: dict) -> list[dict]:
        return json.loads(_take(lib().po_doc_apply_change(self._h, json.dumps(change).encode("utf-8"))))

    def getTextWithFormatting(self, path=("text",)) -> list[dict]:
        assert list(path) == ["text"]
        return json.loads(_take(lib().po_doc_spans(self._h)))

    @property
    def root(self) -> dict:
        return json.loads(_take(lib().po_doc_root(self._h)))

    def getRoot(self) -> dict:
        return self.root

    @property
    def clock(self) -> dict:
        return json.loads(_take(lib().po_doc_clock(self._h)))

    def getCursor(self, path, index: int) -> dict:
        return json.loads(_take(lib().po_doc_get_cursor(self._h, index)))

    def resolveCursor(self, cursor: dict) -> int:
        return int(_take(lib().po_doc_resolve_cursor(self._h, json.dumps(cursor).encode("utf-8"))))

    def elements(self) -> list[dict]:
        """Element sequence incl. tombstones (private `metadata` in the reference, fuzz.ts:214 reaches in)."""
        return json.loads(_take(lib().po_doc_elements(self._h)))
