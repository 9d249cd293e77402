"""`Micromerge` facade over the batch engine — the reference's class surface (reference src/micromerge.ts:262).

* `applyChange` (:499) runs the reference's causal admission checks and buffers the change (the op log).
* `getTextWithFormatting` (:516) / `root` (:290) materialise the document on the GPU through the C-ABI.
* `change` (:308), `getCursor` / `resolveCursor` (:465/:475) — SURVEY.md §8(f) rows 2 and 4 — are generated on the host
  from the element sequence the engine emits (`PT_FLAG_EMIT_SEQUENCE`): visible index -> elemId with
  `lookAfterTombstones` (:762-805) and `changeMark`'s boundary choice (reference src/peritext.ts:458-501).  Local ops are
  applied to the host mirror incrementally (their opIds are larger than everything known, so they land right after their
  reference element); remote changes invalidate the mirror and the next access re-materialises on the GPU.

* The Patch stream (`applyChange` / `applyChanges` return values) comes from the device: the engine
  is created with `PT_FLAG_EMIT_PATCHES` and the patch kernel (csrc/patch_kernel.cuh) derives the patches of every op of
  the log in the same pass that materialises the document.  `applyChanges(changes)` admits a whole list of changes (causal
  retry, as the reference's test/merge.ts:4-23) and derives all their patches in ONE device pass; `applyChange` is the
  batch of one.  Logs too large for the device patch kernel, and the ops `change()` generates locally (applied to the host
  mirror without touching the GPU), use the same closed forms on the host (`peritext_b200/patches.py`).

Deviations from the reference class, all at the admission boundary (documented, tested in tests/test_facade.py):
  * an insert whose opId is SMALLER than its reference element's (a hand-made change that chose a startOp below an element
    it depends on) is refused by `applyChange` with RangeError — the reference would merge it; the engine's closed form
    needs Lamport counters (the reference's own `change()` always produces them, src/micromerge.ts:487, 511);
  * `applyChange` validates the objects of ALL ops of a change before it bumps the clock (the reference bumps first and
    applies ops up to the failing one).
"""
from __future__ import annotations

import copy

from .packing import RangeError, _root_text_list, decode_spans, js_key, pack_logs, parse_op_id, patch_stream, token_str
from .patches import ArrivalHistory, derive_patch

_default_engine = None
INCLUSIVE = {"strong": True, "em": True, "comment": False, "link": False}   # markSpec.inclusive, reference src/schema.ts:45-96


def default_engine():
    global _default_engine
    if _default_engine is None:
        from .engine import BatchEngine
        _default_engine = BatchEngine(0, emit_sequence=True, emit_patches=True)
    return _default_engine


def compareOpIds(id1: str, id2: str) -> int:
    """reference src/micromerge.ts:812-827"""
    if id1 == id2:
        return 0
    c1, a1 = parse_op_id(id1)
    c2, a2 = parse_op_id(id2)
    return -1 if (c1 < c2 or (c1 == c2 and js_key(a1) < js_key(a2))) else 1


class JsError(Exception):
    """JS Error thrown by the reference (non-Range)."""


class Micromerge:
    contentKey = "text"  # reference src/micromerge.ts:264

    def __init__(self, actorId: str, engine=None, patches: bool = True):
        """`patches=False`: `applyChange` / `change` return `[]` and do not materialise (bulk ingest: buffer many changes,
        read the result once)."""
        self.actorId = actorId
        self._want_patches = patches
        self.clock: dict[str, int] = {}       # :273
        self._seq = 0                         # :269
        self._maxOp = 0                       # :271
        self._applied: list[dict] = []        # arrival order == the packed log
        self._objects = {"_root": "map"}      # object id -> kind (:275-281, 541-547)
        self._engine = engine
        self._cache = None                    # (batch, merged) of the last GPU materialisation
        self._mirror = None                   # host mirror of the text list: [[elemId, deleted, hasAfter, value], ...]
        self._mirror_list = None
        self._hist = ArrivalHistory()         # arrival times of the current text list's ops (for the Patch closed forms)
        self._hist_list = None
        self._root_keys: dict[str, str] = {}  # ROOT map: key -> opId that last won LWW (src/micromerge.ts:584-586)
        self._root_vals: dict[str, object] = {}   # ROOT map: primitive values / child placeholders of the keys other than "text"

    # -- reference src/micromerge.ts:499-514 --------------------------------------------------------------------------
    def applyChange(self, change: dict) -> list:
        snap = (dict(self.clock), self._maxOp, dict(self._objects), len(self._applied), dict(self._root_keys), dict(self._root_vals),
                copy.deepcopy(self._hist), self._hist_list)
        self._admit(change)
        try:
            return self._patches_for(change["ops"], local=False)
        except RangeError:
            # e.g. "List element not found": the reference drops the rest of the failing change; this facade drops the whole
            # change, so that the document stays usable (a log the engine rejects would poison every later materialisation)
            self.clock, self._maxOp, self._objects, n, self._root_keys, self._root_vals, self._hist, self._hist_list = snap
            del self._applied[n:]
            self._cache = None; self._mirror = None
            raise

    def applyChanges(self, changes: list) -> list:
        """Admit a list of changes with causal retry (a change whose dependencies are not there yet goes back to the end of
        the queue — the reference's test helper test/merge.ts:4-23) and return the concatenated Patch lists: ONE device pass
        materialises the document and derives every patch."""
        queue, order, iterations = list(changes), [], 0
        while queue:
            ch = queue.pop(0)
            try:
                self._admit(ch)
                order.append(ch)
            except RangeError:
                queue.append(ch)
            iterations += 1
            if iterations > 10000:
                raise RuntimeError("applyChanges did not converge")
        return self._patches_for([op for ch in order for op in ch["ops"]], local=False)

    def _admit(self, change: dict):
        lastSeq = self.clock.get(change["actor"], 0)
        if change["seq"] != lastSeq + 1:
            raise RangeError(f"Expected sequence number {lastSeq + 1}, got {change['seq']}")
        for actor, dep in (change.get("deps") or {}).items():
            if not self.clock.get(actor) or self.clock[actor] < dep:
                raise RangeError(f"Missing dependency: change {dep} by actor {actor}")
        created = {}
        for op in change["ops"]:                                     # :538-547, validated in op order before buffering
            obj = op.get("obj") or "_root"
            if obj not in self._objects and obj not in created:
                raise RangeError(f"Object does not exist: {obj}")
            parse_op_id(op["opId"])
            if op["action"] in ("makeList", "makeMap"):
                created[op["opId"]] = "list" if op["action"] == "makeList" else "map"
            if op["action"] == "set" and op.get("insert") and op.get("elemId") not in (None, "_head") and compareOpIds(op["opId"], op["elemId"]) <= 0:
                raise RangeError(f"insert {op['opId']} does not follow its reference element {op['elemId']} in Lamport order")
        self._objects.update(created)
        self.clock[change["actor"]] = change["seq"]
        self._maxOp = max(self._maxOp, change["startOp"] + len(change["ops"]) - 1)
        self._applied.append(copy.deepcopy(change))                  # Change objects passed in are not mutated
        self._cache = None
        self._mirror = None

    # -- Patch[] of a list of ops that were just appended to the log (closed forms, peritext_b200/patches.py) ------------
    def _root_op(self, op) -> list:
        """ROOT-map LWW bookkeeping in arrival order; returns the makeList patch if the op wins (src/micromerge.ts:584-592)."""
        key = op.get("key")
        if key is None or op["action"] in ("addMark", "removeMark"):
            return []
        cur = self._root_keys.get(key)
        if cur is None or compareOpIds(cur, op["opId"]) == -1:
            self._root_keys[key] = op["opId"]
            if key != "text":                                   # :586-603: the winner's value (child objects other than the text list are
                if op["action"] == "set":                       # not materialised by this engine: empty placeholders)
                    self._root_vals[key] = op.get("value")
                elif op["action"] == "del":
                    self._root_vals.pop(key, None)
                elif op["action"] in ("makeList", "makeMap"):
                    self._root_vals[key] = [] if op["action"] == "makeList" else {}
            if op["action"] == "makeList":
                if key == "text":
                    self._hist, self._hist_list = ArrivalHistory(), op["opId"]
                return [{**op, "path": ["text"]}]
        return []

    def _patches_for(self, ops, local: bool) -> list:
        if not self._want_patches:
            for op in ops:                      # keep the ROOT / arrival bookkeeping consistent, derive nothing
                obj = op.get("obj") or "_root"
                if obj == "_root":
                    self._root_op(op)
                elif obj == self._hist_list and (op["action"] in ("addMark", "removeMark") or op.get("key") is None):
                    self._hist.record(op)
            return []
        pending, out = [], []
        for op in ops:
            obj = op.get("obj") or "_root"
            if obj == "_root":
                out.append(("root", self._root_op(op)))
            elif obj == self._hist_list and (op["action"] in ("addMark", "removeMark") or op.get("key") is None):
                t, emits = self._hist.record(op)
                out.append(("list", (op, t, emits)))
                pending.append(op)
        if not pending:
            return [p for kind, ps in out if kind == "root" for p in ps]
        device = None
        if not local:
            # remote changes: the device derived the patches of every op of the log while materialising it
            batch, merged, dp = self._materialise()
            if int(merged.results[0]["status"]) != 0:
                from .packing import LOG_STATUS
                raise RangeError(LOG_STATUS.get(int(merged.results[0]["status"]), "engine status %d" % int(merged.results[0]["status"])))
            if dp is not None and int(dp.status[0]) == 0:
                lid = self._text_list_id()
                log_ops = [op for ch in self._applied for op in ch["ops"]
                           if op.get("obj") == lid and (op["action"] in ("addMark", "removeMark") or (op["action"] == "set" and op.get("insert")) or (op["action"] == "del" and op.get("key") is None))]
                device = patch_stream(batch, dp, 0, log_ops)[len(log_ops) - len(pending):]
        pos = None
        patches, k = [], 0
        for kind, item in out:
            if kind == "root":
                patches += item
            else:
                op, t, emits = item
                if device is not None:
                    patches += device[k]
                elif emits:
                    if pos is None:
                        pos = {e[0]: j for j, e in enumerate(self._meta())}      # final positions (GPU materialisation or local mirror)
                    patches += derive_patch(op, t, pos, self._hist)
                k += 1
        return patches

    # -- GPU materialisation --------------------------------------------------------------------------------------------
    def _text_list_id(self):
        return _root_text_list(self._applied)

    def _materialise(self):
        """(packed log, merged result, device patches or None) — one device pass, cached until the log changes."""
        if self._cache is None:
            batch = pack_logs([self._applied])
            eng = self._engine or default_engine()
            self.device_passes = getattr(self, "device_passes", 0) + 1
            if getattr(eng, "emit_patches", False) and self._want_patches:
                merged, dp = eng.run_with_patches(batch)
                self._cache = (batch, merged, dp)
            else:
                self._cache = (batch, eng.run(batch), None)
        return self._cache

    def _meta(self):
        """Host mirror of the reference's `metadata[textList]` (+ the visible values)."""
        lid = self._text_list_id()
        if self._mirror is not None and self._mirror_list == lid:
            return self._mirror
        batch, merged, _ = self._materialise()
        if int(merged.results[0]["status"]) != 0:
            raise RangeError("List element not found")
        if merged.seq is None:
            raise JsError("engine was created without emit_sequence; change()/cursors need the element sequence")
        ins, _ = batch.log_slice(0)
        actors = batch.log_actors[0]
        cmap = batch.log_counters[0] if batch.log_counters else None     # dense counter rank -> original counter
        mirror = []
        for e in merged.sequence(0):
            r = ins[int(e) & 0x3FFFFFFF]
            ctr = int(r["ctr"]) if cmap is None else int(cmap[int(r["ctr"])])
            eid = f"{ctr}@{actors[int(r['actor'])]}"
            tok = int(r["payload"]) & 0x3FFFFFFF
            # bit 30: the element's markOpsAfter slot is defined (some applied op starts / ends `after` it, peritext.ts:239-241)
            mirror.append([eid, bool(int(e) >> 31), bool((int(e) >> 30) & 1), token_str(tok, batch.values)])
        self._mirror, self._mirror_list = mirror, lid
        return mirror

    # -- reference src/micromerge.ts:762-805 ------------------------------------------------------------------------------
    @staticmethod
    def _getListElementId(meta, index: int, lookAfterTombstones: bool = False) -> str:
        visible = -1
        for metaIndex, element in enumerate(meta):
            if not element[1]:
                visible += 1
                if visible == index:
                    if lookAfterTombstones:
                        elemIndex, peek, latest = metaIndex, metaIndex + 1, 0
                        while peek < len(meta) and meta[peek][1]:
                            if meta[peek][2]:
                                latest = peek
                            peek += 1
                        if latest:
                            elemIndex = latest
                        return meta[elemIndex][0]
                    return element[0]
        raise RangeError(f"List index out of bounds: {index}")

    def _findListElement(self, meta, elemId: str):
        """reference src/micromerge.ts:731-755 -> (index, visible)"""
        visible = 0
        for index, element in enumerate(meta):
            if element[0] == elemId:
                return index, visible
            if not element[1]:
                visible += 1
        raise RangeError(f"List element not found: {elemId}")

    # -- local op application to the mirror (reference applyListInsert :614-672, applyListUpdate :677-724, marks) --------
    def _apply_local(self, meta, op):
        if op["action"] == "set":
            if op["elemId"] == "_head":
                index = 0
            else:
                index = self._findListElement(meta, op["elemId"])[0] + 1
            while index < len(meta) and compareOpIds(op["opId"], meta[index][0]) < 0:
                index += 1
            meta.insert(index, [op["opId"], False, False, op["value"]])
        elif op["action"] == "del":
            meta[self._findListElement(meta, op["elemId"])[0]][1] = True
        else:
            if op["end"]["type"] == "after":
                meta[self._findListElement(meta, op["end"]["elemId"])[0]][2] = True

    # -- reference src/micromerge.ts:308-441 ------------------------------------------------------------------------------
    def change(self, ops: list[dict]) -> dict:
        deps = dict(self.clock)
        self._seq += 1
        self.clock[self.actorId] = self._seq
        change = {"actor": self.actorId, "seq": self._seq, "deps": deps, "startOp": self._maxOp + 1, "ops": []}
        self._applied.append(change)       # ops are appended as they are generated (makeNewOp applies each op at once, :483-493)
        self._cache = None

        def make(op):
            self._maxOp += 1
            op = {"opId": f"{self._maxOp}@{self.actorId}", **op}
            change["ops"].append(op)
            return op

        for inputOp in ops:
            path = list(inputOp.get("path") or [])
            if path == []:
                objId, kind = "_root", "map"
            elif path == ["text"]:
                objId = self._text_list_id()
                if objId is None:
                    raise JsError("Child not found: text in _root")
                kind = self._objects.get(objId, "list")
            else:
                raise RangeError(f"No object at path {path!r}")
            action = inputOp["action"]
            if kind == "list":
                meta = self._meta()
                visible_len = sum(1 for e in meta if not e[1])
                if action == "insert":
                    elemId = "_head" if inputOp["index"] == 0 else self._getListElementId(meta, inputOp["index"] - 1, True)
                    for value in inputOp["values"]:
                        if not isinstance(value, str):
                            raise JsError("Expected value inserted into text to be a string")
                        op = make({"action": "set", "obj": objId, "elemId": elemId, "insert": True, "value": value})
                        self._apply_local(meta, op)
                        elemId = op["opId"]
                elif action == "delete":
                    for _ in range(inputOp["count"]):
                        op = make({"action": "del", "obj": objId, "elemId": self._getListElementId(meta, inputOp["index"])})
                        self._apply_local(meta, op)
                elif action in ("addMark", "removeMark"):
                    mt = inputOp["markType"]
                    start = {"type": "before", "elemId": self._getListElementId(meta, inputOp["startIndex"])}      # peritext.ts:488
                    if INCLUSIVE[mt] and inputOp["endIndex"] >= visible_len:
                        end = {"type": "endOfText"}                                                                 # :491-492
                    elif INCLUSIVE[mt]:
                        end = {"type": "before", "elemId": self._getListElementId(meta, inputOp["endIndex"])}       # :494
                    else:
                        end = {"type": "after", "elemId": self._getListElementId(meta, inputOp["endIndex"] - 1)}    # :496
                    body = {"action": action, "obj": objId, "start": start, "end": end, "markType": mt}
                    if inputOp.get("attrs"):
                        body["attrs"] = inputOp["attrs"]
                    op = make(body)
                    self._apply_local(meta, op)
                elif action == "del":
                    raise JsError("Use the remove action")
                else:
                    raise JsError("Unimplemented")
            else:
                if action in ("makeList", "makeMap", "del"):
                    op = make({"action": action, "obj": objId, "key": inputOp["key"]})
                    if action != "del":
                        self._objects[op["opId"]] = "list" if action == "makeList" else "map"
                    self._mirror = None
                elif action == "set":
                    make({"action": "set", "obj": objId, "key": inputOp["key"], "value": inputOp.get("value")})
                else:
                    raise JsError(f"Not a list: {path}")
        self._cache = None                 # the GPU materialisation (if any) predates the ops generated above
        return {"change": copy.deepcopy(change), "patches": self._patches_for(change["ops"], local=True)}

    # -- reference src/micromerge.ts:465-477 ------------------------------------------------------------------------------
    def getCursor(self, path, index: int) -> dict:
        return {"objectId": self._text_list_id(), "elemId": self._getListElementId(self._meta(), index)}

    def resolveCursor(self, cursor: dict) -> int:
        return self._findListElement(self._meta(), cursor["elemId"])[1]

    # -- reference src/micromerge.ts:516-529 -> src/peritext.ts:337 ----------------------------------------------------
    def getTextWithFormatting(self, path=("text",)) -> list[dict]:
        if list(path) != ["text"]:
            raise RangeError(f"No object at path {list(path)!r}")
        if self._text_list_id() is None:
            raise JsError("Child not found: text in _root")          # :458
        batch, merged, _ = self._materialise()
        return decode_spans(batch, merged, 0)

    @property
    def root(self) -> dict:
        """`{text: [...visible values]}` (reference :290; fuzz.ts:33 and the tests read `root.text`)."""
        if self._text_list_id() is None:
            return dict(self._root_vals)
        if self._mirror is not None and self._mirror_list == self._text_list_id():
            return {**self._root_vals, "text": [e[3] for e in self._mirror if not e[1]]}
        batch, merged, _ = self._materialise()
        if int(merged.results[0]["status"]) != 0:
            raise RangeError("List element not found")
        return {**self._root_vals, "text": [token_str(int(t), batch.values) for t in merged.tokens(0)]}

    def getRoot(self) -> dict:
        return self.root
