"""`Micromerge` facade over the batch engine — the reference's class surface (reference src/micromerge.ts:262) for the
hot path: `applyChange` (:499) buffers the change after the reference's causal admission checks, and
`getTextWithFormatting` (:516) materialises the document on the GPU through the C-ABI.

Not built in round 1 (SURVEY.md §8f "next" rows): the Patch stream returned by `applyChange` (an empty list is returned)
and local op generation `change()` / cursors, which need the materialised element order on the host.
"""
from __future__ import annotations

import copy

from .packing import RangeError, decode_spans, pack_logs, parse_op_id, token_str

_default_engine = None


def default_engine():
    global _default_engine
    if _default_engine is None:
        from .engine import BatchEngine
        _default_engine = BatchEngine(0)
    return _default_engine


class Micromerge:
    contentKey = "text"  # reference src/micromerge.ts:264

    def __init__(self, actorId: str, engine=None):
        self.actorId = actorId
        self.clock: dict[str, int] = {}       # :273
        self._maxOp = 0                       # :271
        self._applied: list[dict] = []        # arrival order == the packed log
        self._objects = {"_root"}             # object ids created so far (makeList / makeMap), :541-547
        self._engine = engine
        self._cache = None

    # -- reference src/micromerge.ts:499-514 --------------------------------------------------------------------------
    def applyChange(self, change: dict) -> list:
        lastSeq = self.clock.get(change["actor"], 0)
        if change["seq"] != lastSeq + 1:
            raise RangeError(f"Expected sequence number {lastSeq + 1}, got {change['seq']}")
        for actor, dep in (change.get("deps") or {}).items():
            if not self.clock.get(actor) or self.clock[actor] < dep:
                raise RangeError(f"Missing dependency: change {dep} by actor {actor}")
        for op in change["ops"]:                                     # :538-540, checked before buffering
            obj = op.get("obj") or "_root"
            if obj not in self._objects:
                raise RangeError(f"Object does not exist: {obj}")
            if op["action"] in ("makeList", "makeMap"):
                self._objects.add(op["opId"])
            parse_op_id(op["opId"])
        self.clock[change["actor"]] = change["seq"]
        self._maxOp = max(self._maxOp, change["startOp"] + len(change["ops"]) - 1)
        self._applied.append(copy.deepcopy(change))                  # Change objects passed in are not mutated
        self._cache = None
        return []   # Patch[]: SURVEY.md §8(f) row 1 — not derived by the batch engine yet

    def change(self, ops):
        raise NotImplementedError("change() (local op generation, reference src/micromerge.ts:308) is a SURVEY.md §8(f) "
                                  "'next' row; generate Change objects with the reference or the test oracle")

    def _materialise(self):
        if self._cache is None:
            batch = pack_logs([self._applied])
            eng = self._engine or default_engine()
            self._cache = (batch, eng.run(batch))
        return self._cache

    # -- reference src/micromerge.ts:516-529 -> src/peritext.ts:337 ----------------------------------------------------
    def getTextWithFormatting(self, path=("text",)) -> list[dict]:
        if list(path) != ["text"]:
            raise RangeError(f"No object at path {list(path)!r}")
        from .packing import _root_text_list
        if _root_text_list(self._applied) is None:
            raise KeyError("Child not found: text in _root")         # :458
        batch, merged = self._materialise()
        return decode_spans(batch, merged, 0)

    @property
    def root(self) -> dict:
        """`{text: [...visible values]}` (reference :290; fuzz.ts:33 and the tests read `root.text`)."""
        from .packing import _root_text_list
        if _root_text_list(self._applied) is None:
            return {}
        batch, merged = self._materialise()
        if int(merged.results[0]["status"]) != 0:
            raise RangeError("List element not found")
        return {"text": [token_str(int(t), batch.values) for t in merged.tokens(0)]}

    def getRoot(self) -> dict:
        return self.root
