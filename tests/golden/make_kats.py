#!/usr/bin/env python3
"""Transcribe the reference's known-answer tests into tests/golden/kats.json.

Source: /root/reference/test/micromerge.ts (inkandswitch/peritext @89c162d3).  The reference is TypeScript and
cannot be executed in this image (no Node.js), so the test file is *parsed*, not run:

* every ``testConcurrentWrites({...})`` call (reference test/micromerge.ts:46-86 defines the harness) has its
  object-literal argument parsed by the small JS-literal parser below -> kind "concurrent";
* the 15 free-form ``it(...)`` cases (basic, patches, comment/link flatten, cursors) are transcribed by hand in
  ``SCRIPTED`` below; each entry names the ``it`` title and the script checks that this title really occurs at the
  cited line, so a drifted reference makes the generation fail instead of silently going stale.

Run (in the dev container, where /root/reference exists):  python tests/golden/make_kats.py
The output is committed; nothing at test time reads /root/reference.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = "/root/reference/test/micromerge.ts"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kats.json")


# ------------------------------------------------------------------------------------------------
# JS object-literal subset parser (objects, arrays, double-quoted strings, numbers, booleans,
# bare identifier keys, trailing commas, // comments).
# ------------------------------------------------------------------------------------------------
class JsLit:
    def __init__(self, s: str, pos: int):
        self.s, self.p = s, pos

    def ws(self):
        s = self.s
        while self.p < len(s):
            if s[self.p] in " \t\r\n":
                self.p += 1
            elif s.startswith("//", self.p):
                self.p = s.index("\n", self.p)
            elif s.startswith("/*", self.p):
                self.p = s.index("*/", self.p) + 2
            else:
                break

    def value(self):
        self.ws()
        c = self.s[self.p]
        if c == "{":
            self.p += 1
            out = {}
            while True:
                self.ws()
                if self.s[self.p] == "}":
                    self.p += 1
                    return out
                if self.s[self.p] == '"':
                    key = self.string()
                else:
                    m = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*").match(self.s, self.p)
                    assert m, f"bad key at {self.p}: {self.s[self.p:self.p+30]!r}"
                    key = m.group(0)
                    self.p = m.end()
                self.ws()
                assert self.s[self.p] == ":", f"expected ':' at {self.p}: {self.s[self.p:self.p+30]!r}"
                self.p += 1
                out[key] = self.value()
                self.ws()
                if self.s[self.p] == ",":
                    self.p += 1
        if c == "[":
            self.p += 1
            out = []
            while True:
                self.ws()
                if self.s[self.p] == "]":
                    self.p += 1
                    return out
                out.append(self.value())
                self.ws()
                if self.s[self.p] == ",":
                    self.p += 1
        if c == '"':
            v = self.string()
            if self.s.startswith('.split("")', self.p):   # e.g. values: "ara".split("")  (reference test/micromerge.ts:613,757)
                self.p += len('.split("")')
                return list(v)
            return v
        m = re.compile(r"-?[0-9]+(\.[0-9]+)?").match(self.s, self.p)
        if m:
            self.p = m.end()
            t = m.group(0)
            return float(t) if "." in t else int(t)
        for lit, val in (("true", True), ("false", False), ("null", None)):
            if self.s.startswith(lit, self.p):
                self.p += len(lit)
                return val
        raise AssertionError(f"unexpected token at {self.p}: {self.s[self.p:self.p+40]!r}")

    def string(self):
        assert self.s[self.p] == '"'
        m = re.compile(r'"((?:[^"\\]|\\.)*)"').match(self.s, self.p)
        self.p = m.end()
        return json.loads('"' + m.group(1) + '"')


def line_of(src: str, pos: int) -> int:
    return src.count("\n", 0, pos) + 1


DEFAULT = "The Peritext editor"  # reference test/micromerge.ts:9, test/generateDocs.ts:6
TEXT_CHARS = list(DEFAULT)


def P(op):  # add path like the tests do
    return {"path": ["text"], **op}


# Hand-transcribed free-form cases.  (title, reference line of the `it(`, steps)
SCRIPTED = [
    ("can insert and delete text", 89, dict(initialText="abcde", steps=[
        dict(do="change", doc=1, ops=[P(dict(action="delete", index=0, count=3))]),
        dict(do="expectRootTextJoined", doc=1, value="de"),                                   # :104
    ])),
    ("records local changes in the deps clock", 110, dict(initialText="a", steps=[
        dict(do="change", doc=2, ops=[P(dict(action="insert", index=1, values=["b"]))], save="change2"),
        dict(do="applyChange", doc=1, change="change2"),                                      # :119-121 doesNotThrow
        dict(do="expectRootText", doc=1, value=["a", "b"]),                                   # :123
        dict(do="expectRootText", doc=2, value=["a", "b"]),                                   # :124
    ])),
    ("produces the correct patch for applying a simple insertion", 915, dict(steps=[
        dict(do="change", doc=1, ops=[P(dict(action="insert", index=7, values=["a"]))], save="c"),
        dict(do="applyChange", doc=2, change="c",
             expectPatches=[dict(path=["text"], action="insert", index=7, values=["a"], marks={})]),   # :928-931
    ])),
    ("produces a patch with adjusted insertion index on concurrent inserts", 938, dict(steps=[
        dict(do="change", doc=1, ops=[P(dict(action="insert", index=1, values=["a", "b", "c"]))]),
        dict(do="change", doc=2, ops=[P(dict(action="insert", index=2, values=["b"]))], save="change2"),
        dict(do="applyChange", doc=1, change="change2",
             expectPatches=[dict(path=["text"], action="insert", index=5, values=["b"], marks={})]),   # :967-975
    ])),
    ("produces the correct patch for applying a simple deletion", 981, dict(steps=[
        dict(do="change", doc=1, ops=[P(dict(action="delete", index=5, count=1))], save="c"),
        dict(do="applyChange", doc=2, change="c",
             expectPatches=[dict(path=["text"], action="delete", index=5, count=1)]),                  # :995
    ])),
    ("turns a multi-char deletion into multiple single char deletions", 1001, dict(steps=[
        dict(do="change", doc=1, ops=[P(dict(action="delete", index=5, count=2))], save="c"),
        dict(do="applyChange", doc=2, change="c",
             expectPatches=[dict(path=["text"], action="delete", index=5, count=1),
                            dict(path=["text"], action="delete", index=5, count=1)]),                  # :1015-1028
    ])),
    ("returns a single comment in the flattened spans", 1033, dict(steps=[
        dict(do="change", doc=1, ops=[P(dict(action="addMark", startIndex=4, endIndex=12, markType="comment", attrs={"id": "abc-123"}))]),
        dict(do="expectRootText", doc=1, value=TEXT_CHARS),                                   # :1049
        dict(do="expectSpans", doc=1, value=[
            {"marks": {}, "text": "The "},
            {"marks": {"comment": [{"id": "abc-123"}]}, "text": "Peritext"},
            {"marks": {}, "text": " editor"}]),                                               # :1051-1058
    ])),
    ("correctly flattens two comments from the same user", 1061, dict(steps=[
        dict(do="change", doc=1, ops=[
            P(dict(action="addMark", startIndex=0, endIndex=12, markType="comment", attrs={"id": "abc-123"})),
            P(dict(action="addMark", startIndex=4, endIndex=19, markType="comment", attrs={"id": "def-789"}))]),
        dict(do="expectRootText", doc=1, value=TEXT_CHARS),                                   # :1086
        dict(do="expectSpans", doc=1, value=[
            {"marks": {"comment": [{"id": "abc-123"}]}, "text": "The "},
            {"marks": {"comment": [{"id": "abc-123"}, {"id": "def-789"}]}, "text": "Peritext"},
            {"marks": {"comment": [{"id": "def-789"}]}, "text": " editor"}]),                 # :1088-1100
    ])),
    ("returns a single link in the flattened spans", 1146, dict(steps=[
        dict(do="change", doc=1, ops=[P(dict(action="addMark", startIndex=4, endIndex=12, markType="link", attrs={"url": "https://inkandswitch.com"}))]),
        dict(do="expectRootText", doc=1, value=TEXT_CHARS),                                   # :1162
        dict(do="expectSpans", doc=1, value=[
            {"marks": {}, "text": "The "},
            {"marks": {"link": {"url": "https://inkandswitch.com"}}, "text": "Peritext"},
            {"marks": {}, "text": " editor"}]),                                               # :1164-1175
    ])),
    ("can resolve a cursor position", 1291, dict(steps=[
        dict(do="getCursor", doc=1, index=5, save="cursor"),
        dict(do="resolveCursor", doc=1, cursor="cursor", expect=5),                           # :1301
    ])),
    ("increments cursor position when insert happens before cursor", 1304, dict(steps=[
        dict(do="getCursor", doc=1, index=5, save="cursor"),
        dict(do="change", doc=1, ops=[P(dict(action="insert", index=0, values=["a", "b", "c"]))]),
        dict(do="resolveCursor", doc=1, cursor="cursor", expect=8),                           # :1324
    ])),
    ("does not move cursor position when insert happens after cursor", 1327, dict(steps=[
        dict(do="getCursor", doc=1, index=5, save="cursor"),
        dict(do="change", doc=1, ops=[P(dict(action="insert", index=7, values=["a", "b", "c"]))]),
        dict(do="resolveCursor", doc=1, cursor="cursor", expect=5),                           # :1347
    ])),
    ("moves cursor left if deletion happens before cursor", 1350, dict(steps=[
        dict(do="getCursor", doc=1, index=5, save="cursor"),
        dict(do="change", doc=1, ops=[P(dict(action="delete", index=0, count=3))]),
        dict(do="resolveCursor", doc=1, cursor="cursor", expect=2),                           # :1370
    ])),
    ("doesn't move cursor if deletion happens after cursor", 1373, dict(steps=[
        dict(do="getCursor", doc=1, index=5, save="cursor"),
        dict(do="change", doc=1, ops=[P(dict(action="delete", index=7, count=3))]),
        dict(do="resolveCursor", doc=1, cursor="cursor", expect=5),                           # :1393
    ])),
    ("returns index 0 if everything before the cursor is deleted", 1396, dict(steps=[
        dict(do="getCursor", doc=1, index=5, save="cursor"),
        dict(do="change", doc=1, ops=[P(dict(action="delete", index=0, count=7))]),
        dict(do="resolveCursor", doc=1, cursor="cursor", expect=0),                           # :1415
    ])),
]


def main() -> int:
    src = open(REF, encoding="utf-8").read()
    lines = src.split("\n")
    its = [(m.start(), m.group(1)) for m in re.finditer(r'\bit\(\s*"((?:[^"\\]|\\.)*)"', src)]
    assert len(its) == 46, f"expected 46 it() cases, found {len(its)}"

    kats = []
    seen_titles = set()
    # concurrent cases: each testConcurrentWrites( call inside an it()
    for m in re.finditer(r"\btestConcurrentWrites\(\s*\{", src):
        if line_of(src, m.start()) == 46:  # the definition itself is `const testConcurrentWrites = (args...` (no `({`)
            continue
        brace = src.index("{", m.start())
        spec = JsLit(src, brace).value()
        owner = max((it for it in its if it[0] < m.start()), key=lambda t: t[0])
        title = owner[1]
        seen_titles.add(title)
        assert set(spec) <= {"initialText", "preOps", "inputOps1", "inputOps2", "expectedResult"}, spec.keys()
        kats.append({
            "kind": "concurrent",
            "name": title,
            "line": line_of(src, m.start()),
            "initialText": spec.get("initialText", DEFAULT),
            "preOps": spec.get("preOps"),
            "inputOps1": spec.get("inputOps1", []),
            "inputOps2": spec.get("inputOps2", []),
            "expectedResult": spec["expectedResult"],
        })
    assert len(kats) == 31, f"expected 31 testConcurrentWrites calls, found {len(kats)}"

    for title, line, body in SCRIPTED:
        assert f'it("{title}"' in lines[line - 1], f"title mismatch at line {line}: {lines[line-1]!r}"
        seen_titles.add(title)
        kats.append({"kind": "script", "name": title, "line": line,
                     "initialText": body.get("initialText", DEFAULT), "steps": body["steps"]})

    missing = [t for _, t in its if t not in seen_titles]
    assert not missing, f"it() cases not transcribed: {missing}"
    kats.sort(key=lambda k: k["line"])
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump({"source": "inkandswitch/peritext@89c162d3 test/micromerge.ts", "count": len(kats), "kats": kats}, f,
                  indent=1, ensure_ascii=False)
        f.write("\n")
    print(f"wrote {len(kats)} KATs to {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
