// peritext_oracle.cpp — TEST INFRASTRUCTURE. CPU restatement of the reference algorithm.
//
//   *** This file is the ORACLE: a checker, not a product path. ***
//   Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
//   load it.  Nothing under peritext_b200/ imports, links or executes it.
//
// It is a sequential, line-by-line restatement of the reference TypeScript (inkandswitch/peritext
// @89c162d3), keeping the reference's data model (element array + optional per-element before/after
// op sets, immutable Sets shared by reference) and its asymptotics (linear findListElement, array
// splices, per-mark full walk).  Citations are reference `file:line`.
//
// Parity pin: the reference cannot be executed in this image (no Node.js); this oracle is pinned
// against the 46 known-answer tests of the reference's test/micromerge.ts, transcribed into
// tests/golden/kats.json by tests/golden/make_kats.py (see tests/test_oracle_kats.py).  Corners the
// reference tests do not pin ({comment: []} vs {}, non-ASCII actor ordering) are "parity unpinned".
//
// Deliberate deviations that make the CPU baseline OPTIMISTIC for the reference (stated in DESIGN.md):
//   * opIds are pre-parsed (ctr, actor) pairs; the reference re-parses two strings with a RegExp on
//     every compareOpIds call (src/micromerge.ts:815-825).
//   * applyAddRemoveMark walks the element array directly; the reference first allocates a 2N-entry
//     positions array (src/peritext.ts:167-171).

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/peritext_b200.h"
#include "../include/pt_digest.h"
#include "json_min.hpp"

using pjson::Value;

namespace po {

// ---------------------------------------------------------------------------------------------
// Errors: mirror `throw new RangeError(...)` / `throw new Error(...)`.
// ---------------------------------------------------------------------------------------------
struct JsError : std::runtime_error {
    std::string kind;
    JsError(const std::string& k, const std::string& m) : std::runtime_error(m), kind(k) {}
};
[[noreturn]] static void rangeError(const std::string& m) { throw JsError("RangeError", m); }
[[noreturn]] static void error(const std::string& m) { throw JsError("Error", m); }

// ---------------------------------------------------------------------------------------------
// Interning of actor ids and attribute values (process-wide; guarded for the JSON interface,
// read-only during the threaded packed replay).
// ---------------------------------------------------------------------------------------------
static std::vector<uint16_t> utf16_of(const std::string& s) {
    std::vector<uint16_t> out;
    size_t i = 0, n = s.size();
    while (i < n) {
        unsigned char c = (unsigned char)s[i];
        uint32_t cp; int len;
        if (c < 0x80) { cp = c; len = 1; }
        else if ((c >> 5) == 6) { cp = c & 0x1F; len = 2; }
        else if ((c >> 4) == 14) { cp = c & 0x0F; len = 3; }
        else { cp = c & 0x07; len = 4; }
        for (int k = 1; k < len && i + k < n; k++) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
        i += len;
        if (cp >= 0x10000) { cp -= 0x10000; out.push_back((uint16_t)(0xD800 + (cp >> 10))); out.push_back((uint16_t)(0xDC00 + (cp & 0x3FF))); }
        else out.push_back((uint16_t)cp);
    }
    return out;
}

struct Interner {
    std::mutex mu;
    std::vector<std::string> names;
    std::vector<std::vector<uint16_t>> u16;  // JS string order is UTF-16 code-unit order
    std::unordered_map<std::string, int> index;
    int intern(const std::string& s) {
        std::lock_guard<std::mutex> g(mu);
        auto it = index.find(s);
        if (it != index.end()) return it->second;
        int id = (int)names.size();
        names.push_back(s); u16.push_back(utf16_of(s)); index.emplace(s, id);
        return id;
    }
    // JS `a < b` on strings
    bool less(int a, int b) const { return a != b && u16[a] < u16[b]; }
    const std::string& name(int id) const { return names[id]; }
};
// deque-like stability is not needed: vectors are only appended under the mutex and the threaded
// replay pre-interns everything it needs before spawning threads.
static Interner g_actors;
static Interner g_strings;  // attrs canonical JSON, comment ids, mark values

// ---------------------------------------------------------------------------------------------
// Operation ids.  "ctr@actor" (src/micromerge.ts:488).  ROOT / HEAD are JS Symbols (:18-19).
// ---------------------------------------------------------------------------------------------
struct OpId {
    uint32_t ctr = 0;
    int32_t actor = -1;  // -1 ROOT, -2 HEAD, -3 undefined
    bool operator==(const OpId& o) const { return ctr == o.ctr && actor == o.actor; }
    bool operator!=(const OpId& o) const { return !(*this == o); }
    uint64_t key() const { return ((uint64_t)(uint32_t)actor << 32) | ctr; }
};
static const OpId ROOT{0, -1};
static const OpId HEAD{0, -2};
static const OpId UNDEF{0, -3};

static std::string opIdStr(const OpId& id) {
    if (id == ROOT) return "_root";
    if (id == HEAD) return "_head";
    if (id == UNDEF) return "undefined";
    return std::to_string(id.ctr) + "@" + g_actors.name(id.actor);
}
// regex ^([0-9]+)@(.*)$  (src/micromerge.ts:815)
static OpId parseOpId(const std::string& s) {
    size_t at = s.find('@');
    if (at == std::string::npos || at == 0) error("Invalid operation ID: " + s);
    uint64_t c = 0;
    for (size_t i = 0; i < at; i++) {
        if (s[i] < '0' || s[i] > '9') error("Invalid operation ID: " + s);
        c = c * 10 + (uint64_t)(s[i] - '0');
    }
    OpId id; id.ctr = (uint32_t)c; id.actor = g_actors.intern(s.substr(at + 1));
    return id;
}
// compareOpIds (src/micromerge.ts:812-827)
static int compareOpIds(const OpId& a, const OpId& b) {
    if (a == b) return 0;
    if (a.ctr < b.ctr || (a.ctr == b.ctr && g_actors.less(a.actor, b.actor))) return -1;
    return +1;
}

// ---------------------------------------------------------------------------------------------
// Mark operations (src/peritext.ts:11-65) and markSpec constants (src/schema.ts:45-96).
// ---------------------------------------------------------------------------------------------
enum MarkType { STRONG = 0, EM = 1, COMMENT = 2, LINK = 3 };  // ALL_MARKS order, src/schema.ts:125
static const char* kMarkNames[4] = {"strong", "em", "comment", "link"};
static const bool kInclusive[4] = {true, true, false, false};      // src/schema.ts:51,58,64,84
static const bool kAllowMultiple[4] = {false, false, true, false}; // src/schema.ts:50,57,77,85
static int markTypeOf(const std::string& s) {
    for (int i = 0; i < 4; i++) if (s == kMarkNames[i]) return i;
    error("Unknown mark type: " + s);
}

enum BoundType { B_BEFORE = 0, B_AFTER = 1, B_START = 2, B_END = 3 };
struct Boundary { int type = B_BEFORE; OpId elemId = UNDEF; };

struct MarkOp {
    OpId opId, obj;
    bool add = true;
    int markType = STRONG;
    Boundary start, end;
    bool hasAttrs = false;
    int attrsId = -1;   // interned canonical JSON of attrs (value stored in the MarkMap)
    int commentId = -1; // interned attrs.id (comment only)
    Value attrs;        // kept for JSON output
};
using OpSet = std::vector<const MarkOp*>;           // JS Set<MarkOperation>: insertion ordered, by identity
using OpSetRef = std::shared_ptr<const OpSet>;      // Sets are never mutated in place, only replaced

// MarkMap (src/peritext.ts:135-137) in interned form.  value ids: interned canonical JSON.
struct MarkMap {
    int single[4] = {-1, -1, -1, -1};         // strong / em / (unused) / link value id, -1 = key absent
    bool hasComment = false;                  // `comment` key present (possibly [])
    std::vector<std::pair<int, int>> comments;  // (comment id string id, attrs id), sorted by id
    bool operator==(const MarkMap& o) const {   // lodash isEqual on plain objects (key order ignored)
        for (int i = 0; i < 4; i++) if (single[i] != o.single[i]) return false;
        if (hasComment != o.hasComment) return false;
        if (comments.size() != o.comments.size()) return false;
        for (size_t i = 0; i < comments.size(); i++) if (comments[i].second != o.comments[i].second) return false;
        return true;
    }
};
static int g_activeTrue = -1;  // interned {"active":true}
static int activeTrueId() {
    if (g_activeTrue < 0) g_activeTrue = g_strings.intern("{\"active\":true}");
    return g_activeTrue;
}

// opsToMarks (src/peritext.ts:294-326)
static MarkMap opsToMarks(const OpSet& ops) {
    MarkMap markMap;
    OpId opIdMap[4] = {UNDEF, UNDEF, UNDEF, UNDEF};
    for (const MarkOp* op : ops) {
        const OpId& existingOpId = opIdMap[op->markType];
        if (!kAllowMultiple[op->markType]) {
            if (existingOpId == UNDEF || compareOpIds(op->opId, existingOpId) == 1) {   // :305
                opIdMap[op->markType] = op->opId;
                if (op->add) markMap.single[op->markType] = op->hasAttrs ? op->attrsId : activeTrueId();  // :308
                else markMap.single[op->markType] = -1;                                                   // :311
            }
        } else {
            bool found = false;
            for (auto& c : markMap.comments) if (c.first == op->commentId) { found = true; break; }
            if (op->add && !found) {                                                    // :315
                // sortBy([...existing, op.attrs], c => c.id): stable ascending by id (JS string order)
                markMap.comments.push_back({op->commentId, op->attrsId});
                std::stable_sort(markMap.comments.begin(), markMap.comments.end(),
                                 [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return g_strings.less(a.first, b.first); });
                markMap.hasComment = true;
            } else if (!op->add) {                                                      // :319-320
                std::vector<std::pair<int, int>> kept;
                for (auto& c : markMap.comments) if (c.first != op->commentId) kept.push_back(c);
                markMap.comments.swap(kept);
                markMap.hasComment = true;   // assigns an array even when empty (quirk Q3)
            }
        }
    }
    return markMap;
}

// ---------------------------------------------------------------------------------------------
// Objects and metadata (src/micromerge.ts:217-257, 275-281)
// ---------------------------------------------------------------------------------------------
struct Elem {  // ListItemMetadata (src/micromerge.ts:237-253)
    OpId elemId, valueId;
    bool deleted = false;
    OpSetRef markOpsBefore, markOpsAfter;  // null = undefined
};
struct MapEntry { bool isChild = false; OpId child = UNDEF; Value prim; };
struct Object {
    bool isList = false;
    // list
    std::vector<Elem> meta;           // this.metadata[objId] (array)
    std::vector<std::string> text;    // this.objects[objId]  (visible values)
    // map
    std::vector<std::pair<std::string, MapEntry>> fields;            // this.objects[objId] (insertion ordered)
    std::vector<std::pair<std::string, OpId>> keyMeta;               // metadata[key] = opId
    std::vector<std::pair<std::string, OpId>> children;              // metadata[CHILDREN][key]
    template <class V> static V* find(std::vector<std::pair<std::string, V>>& v, const std::string& k) {
        for (auto& kv : v) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    template <class V> static void put(std::vector<std::pair<std::string, V>>& v, const std::string& k, const V& val) {
        for (auto& kv : v) if (kv.first == k) { kv.second = val; return; }
        v.push_back({k, val});
    }
};

enum Action { A_MAKELIST, A_MAKEMAP, A_SET, A_DEL, A_ADDMARK, A_REMOVEMARK };
static const char* kActionNames[6] = {"makeList", "makeMap", "set", "del", "addMark", "removeMark"};

struct Op {  // Operation (src/micromerge.ts:143-212)
    OpId opId, obj = ROOT;
    Action action = A_SET;
    bool hasElemId = false; OpId elemId = UNDEF;
    bool insert = false;
    bool hasKey = false; std::string key;
    Value value;                       // set value
    std::shared_ptr<MarkOp> mark;      // add/removeMark payload (object identity = this pointer)
};

struct Patch {  // Patch (src/micromerge.ts:25-30)
    enum Kind { INSERT, DELETE, MARK, MAKELIST } kind = INSERT;
    uint32_t index = 0;
    std::string value; MarkMap marks;   // insert
    bool add = true; int markType = 0; uint32_t startIndex = 0, endIndex = 0; bool hasAttrs = false; Value attrs;  // mark
    Op op;                              // makeList: {...op, path:["text"]}
};

// ---------------------------------------------------------------------------------------------
// JSON <-> structures
// ---------------------------------------------------------------------------------------------
static Value markMapJson(const MarkMap& m) {
    Value o = Value::object();
    // canonical key order = ALL_MARKS (quirk Q8: JS key order is apply-order dependent; equality ignores it)
    for (int t : {STRONG, EM}) if (m.single[t] >= 0) o.set(kMarkNames[t], pjson::parse(g_strings.name(m.single[t])));
    if (m.hasComment) {
        Value arr = Value::array();
        for (auto& c : m.comments) arr.push(pjson::parse(g_strings.name(c.second)));
        o.set("comment", arr);
    }
    if (m.single[LINK] >= 0) o.set("link", pjson::parse(g_strings.name(m.single[LINK])));
    return o;
}
static Value boundaryJson(const Boundary& b) {
    Value o = Value::object();
    static const char* names[4] = {"before", "after", "startOfText", "endOfText"};
    o.set("type", Value::string(names[b.type]));
    if (b.type == B_BEFORE || b.type == B_AFTER) o.set("elemId", Value::string(opIdStr(b.elemId)));
    return o;
}
static Boundary boundaryFrom(const Value& v) {
    Boundary b;
    const Value* t = v.get("type");
    if (!t || !t->isStr()) error("bad boundary");
    if (t->str == "before") b.type = B_BEFORE; else if (t->str == "after") b.type = B_AFTER;
    else if (t->str == "startOfText") b.type = B_START; else if (t->str == "endOfText") b.type = B_END;
    else error("bad boundary type " + t->str);
    if (b.type == B_BEFORE || b.type == B_AFTER) {
        const Value* e = v.get("elemId");
        if (!e || !e->isStr()) error("boundary without elemId");
        b.elemId = parseOpId(e->str);
    }
    return b;
}
static void setMarkAttrs(MarkOp& m, const Value* attrs) {
    if (attrs && attrs->isObj()) {
        m.hasAttrs = true; m.attrs = *attrs;
        m.attrsId = g_strings.intern(pjson::canon(*attrs));
        if (m.markType == COMMENT) {
            const Value* id = attrs->get("id");
            m.commentId = g_strings.intern(id && id->isStr() ? id->str : pjson::canon(id ? *id : Value()));
        }
    } else if (m.markType == COMMENT) {
        error("comment mark without attrs");
    }
}
static Value opJson(const Op& op) {
    Value o = Value::object();
    o.set("opId", Value::string(opIdStr(op.opId)));
    o.set("action", Value::string(kActionNames[op.action]));
    o.set("obj", Value::string(opIdStr(op.obj)));
    if (op.action == A_ADDMARK || op.action == A_REMOVEMARK) {
        o.set("start", boundaryJson(op.mark->start));
        o.set("end", boundaryJson(op.mark->end));
        o.set("markType", Value::string(kMarkNames[op.mark->markType]));
        if (op.mark->hasAttrs) o.set("attrs", op.mark->attrs);
    } else {
        if (op.hasElemId) o.set("elemId", Value::string(opIdStr(op.elemId)));
        if (op.insert) o.set("insert", Value::boolean(true));
        if (op.hasKey) o.set("key", Value::string(op.key));
        if (op.action == A_SET) o.set("value", op.value);
    }
    return o;
}
// Accepts the reference's JSON form of an Operation, including the lossy saved traces where Symbol
// valued fields were dropped (quirk Q6): missing `obj` => ROOT, insert without `elemId` => HEAD.
static Op opFrom(const Value& v) {
    Op op;
    const Value* id = v.get("opId"); if (!id || !id->isStr()) error("op without opId");
    op.opId = parseOpId(id->str);
    const Value* a = v.get("action"); if (!a || !a->isStr()) error("op without action");
    int act = -1; for (int i = 0; i < 6; i++) if (a->str == kActionNames[i]) act = i;
    if (act < 0) error("unknown action " + a->str);
    op.action = (Action)act;
    const Value* obj = v.get("obj");
    op.obj = (obj && obj->isStr() && obj->str != "_root") ? parseOpId(obj->str) : ROOT;
    if (op.action == A_ADDMARK || op.action == A_REMOVEMARK) {
        auto m = std::make_shared<MarkOp>();
        m->opId = op.opId; m->obj = op.obj; m->add = op.action == A_ADDMARK;
        const Value* mt = v.get("markType"); if (!mt || !mt->isStr()) error("mark op without markType");
        m->markType = markTypeOf(mt->str);
        const Value* s = v.get("start"); const Value* e = v.get("end");
        if (!s || !e) error("mark op without start/end");
        m->start = boundaryFrom(*s); m->end = boundaryFrom(*e);
        setMarkAttrs(*m, v.get("attrs"));
        op.mark = m;
        return op;
    }
    const Value* ins = v.get("insert"); op.insert = ins && ins->kind == Value::Bool && ins->b;
    const Value* el = v.get("elemId");
    if (el && el->isStr()) { op.hasElemId = true; op.elemId = el->str == "_head" ? HEAD : parseOpId(el->str); }
    else if (op.insert || (op.action == A_DEL && !v.get("key"))) { op.hasElemId = true; op.elemId = HEAD; }
    const Value* k = v.get("key"); if (k && k->isStr()) { op.hasKey = true; op.key = k->str; }
    const Value* val = v.get("value"); if (val) op.value = *val;
    return op;
}
static Value patchJson(const Patch& p) {
    Value path = Value::array(); path.push(Value::string("text"));
    Value o = Value::object();
    switch (p.kind) {
        case Patch::INSERT: {  // src/micromerge.ts:661-671
            o.set("path", path); o.set("action", Value::string("insert")); o.set("index", Value::number(p.index));
            Value vals = Value::array(); vals.push(Value::string(p.value)); o.set("values", vals);
            o.set("marks", markMapJson(p.marks));
            break;
        }
        case Patch::DELETE:    // src/micromerge.ts:696-703
            o.set("path", path); o.set("action", Value::string("delete")); o.set("index", Value::number(p.index));
            o.set("count", Value::number(1));
            break;
        case Patch::MARK:      // src/peritext.ts:251-281
            o.set("action", Value::string(p.add ? "addMark" : "removeMark"));
            o.set("markType", Value::string(kMarkNames[p.markType]));
            o.set("path", path); o.set("startIndex", Value::number(p.startIndex));
            if (p.hasAttrs) o.set("attrs", p.attrs);
            o.set("endIndex", Value::number(p.endIndex));
            break;
        case Patch::MAKELIST:  // src/micromerge.ts:592
            o = opJson(p.op); o.set("path", path);
            break;
    }
    return o;
}

struct Span { std::string text; MarkMap marks; };  // FormatSpanWithText (src/peritext.ts:35-38)

// ---------------------------------------------------------------------------------------------
// peritext.ts functions
// ---------------------------------------------------------------------------------------------
static OpSetRef withOp(const OpSet& cur, const MarkOp* op) {  // new Set([...currentOps, op])
    auto s = std::make_shared<OpSet>(cur);
    if (std::find(s->begin(), s->end(), op) == s->end()) s->push_back(op);
    return s;
}
static OpSetRef withoutOp(const OpSet& cur, const MarkOp* op) {  // new Set([...].filter(o => o !== op))
    auto s = std::make_shared<OpSet>();
    for (const MarkOp* o : cur) if (o != op) s->push_back(o);
    return s;
}
static const OpSet kEmptySet;

// applyAddRemoveMark (src/peritext.ts:154-223) + calculateOpsForPosition (:225-249)
static void applyAddRemoveMark(const MarkOp* op, Object& list, std::vector<Patch>* patches) {
    std::vector<Elem>& metadata = list.meta;
    uint32_t visibleIndex = 0;
    OpSetRef currentOpsRef;                 // new Set() (:176)
    const OpSet* currentOps = &kEmptySet;
    enum { BEFORE, DURING, AFTER } opState = BEFORE;
    bool havePartial = false; uint32_t partialStart = 0;
    const uint32_t objLength = (uint32_t)list.text.size();   // :179

    auto finishPartial = [&](uint32_t endIndex) {            // finishPartialPatch :269-281
        bool notZero = endIndex > partialStart;
        bool affectsVisible = partialStart < objLength;
        if (notZero && affectsVisible && patches) {
            Patch p; p.kind = Patch::MARK; p.add = op->add; p.markType = op->markType;
            p.startIndex = partialStart; p.endIndex = std::min(endIndex, objLength);
            if (op->add && (op->markType == LINK || op->markType == COMMENT)) { p.hasAttrs = op->hasAttrs; p.attrs = op->attrs; }  // :262-264
            patches->push_back(p);
        }
        havePartial = false;
    };

    const size_t n = metadata.size();
    for (size_t i = 0; i < n && opState != AFTER; i++) {
        for (int side = 0; side < 2; side++) {               // positions: [i,"markOpsBefore"], [i,"markOpsAfter"] (:168-171)
            Elem& elMeta = metadata[i];
            OpSetRef& slot = side == 0 ? elMeta.markOpsBefore : elMeta.markOpsAfter;
            if (slot) { currentOpsRef = slot; currentOps = currentOpsRef.get(); }   // :183 (an empty Set is truthy)
            // calculateOpsForPosition :225-249
            OpSetRef changedOps;
            const int opSide = side == 0 ? B_BEFORE : B_AFTER;
            if (op->start.type == opSide && op->start.elemId == elMeta.elemId) {
                opState = DURING; changedOps = withOp(*currentOps, op);              // :236-238
            } else if (op->end.type == opSide && op->end.elemId == elMeta.elemId) {
                opState = AFTER; changedOps = withoutOp(*currentOps, op);            // :239-241
            } else if (opState == DURING && slot) {
                changedOps = withOp(*currentOps, op);                                // :242-244
            }
            if (changedOps) slot = changedOps;                                       // :186
            if (side == 1 && !elMeta.deleted) visibleIndex += 1;                     // :192-196
            if (changedOps) {
                if (havePartial) finishPartial(visibleIndex);                        // :201-205
                if (opState == DURING && patches && !(opsToMarks(*currentOps) == opsToMarks(*changedOps))) {  // :208
                    havePartial = true; partialStart = visibleIndex;
                }
            }
            if (opState == AFTER) break;                                             // :213
        }
    }
    if (havePartial) finishPartial(visibleIndex);                                    // :217-220
}

// findClosestMarkOpsToLeft (src/peritext.ts:405-436), side fixed to "before" as at its only call site (:329)
static OpSet findClosestMarkOpsToLeft(const std::vector<Elem>& metadata, size_t index) {
    for (size_t i = index; i-- > 0;) {
        if (metadata[i].markOpsAfter) return *metadata[i].markOpsAfter;
        if (metadata[i].markOpsBefore) return *metadata[i].markOpsBefore;
    }
    return OpSet();
}
// getActiveMarksAtIndex (src/peritext.ts:328-330)
static MarkMap getActiveMarksAtIndex(const std::vector<Elem>& metadata, size_t index) {
    return opsToMarks(findClosestMarkOpsToLeft(metadata, index));
}

// addCharactersToSpans (src/peritext.ts:438-455)
static void addCharactersToSpans(std::vector<std::string>& characters, const MarkMap& marks, std::vector<Span>& spans) {
    if (characters.empty()) return;
    std::string joined; for (auto& c : characters) joined += c;
    if (!spans.empty() && spans.back().marks == marks) spans.back().text += joined;
    else spans.push_back({joined, marks});
}
// getTextWithFormatting (src/peritext.ts:337-395)
static std::vector<Span> getTextWithFormatting(const Object& list) {
    std::vector<Span> spans;
    std::vector<std::string> characters;
    MarkMap marks;
    size_t visible = 0;
    const auto& metadata = list.meta;
    for (size_t index = 0; index < metadata.size(); index++) {
        const Elem& elMeta = metadata[index];
        bool haveNew = false; MarkMap newMarks;
        if (elMeta.markOpsBefore) { newMarks = opsToMarks(*elMeta.markOpsBefore); haveNew = true; }               // :372-373
        else if (index > 0 && metadata[index - 1].markOpsAfter) { newMarks = opsToMarks(*metadata[index - 1].markOpsAfter); haveNew = true; }  // :374-375
        if (haveNew) { addCharactersToSpans(characters, marks, spans); characters.clear(); marks = newMarks; }   // :378-383
        if (!elMeta.deleted) { characters.push_back(list.text[visible]); visible += 1; }                         // :385-389
    }
    addCharactersToSpans(characters, marks, spans);                                                               // :392
    return spans;
}

// getListElementId (src/micromerge.ts:762-805)
static OpId getListElementId(const std::vector<Elem>& meta, int64_t index, bool lookAfterTombstones) {
    int64_t visible = -1;
    for (size_t metaIndex = 0; metaIndex < meta.size(); metaIndex++) {
        const Elem& element = meta[metaIndex];
        if (!element.deleted) {
            visible++;
            if (visible == index) {
                if (lookAfterTombstones) {
                    size_t elemIndex = metaIndex, peekIndex = metaIndex + 1;
                    size_t latest = 0;  // `if (latestIndexAfterTombstone)` truthiness; 0 is never assigned
                    while (peekIndex < meta.size() && meta[peekIndex].deleted) {
                        if (meta[peekIndex].markOpsAfter) latest = peekIndex;
                        peekIndex++;
                    }
                    if (latest) elemIndex = latest;
                    return meta[elemIndex].elemId;
                }
                return element.elemId;
            }
        }
    }
    rangeError("List index out of bounds: " + std::to_string(index));
}

// ---------------------------------------------------------------------------------------------
// Micromerge (src/micromerge.ts:262-756)
// ---------------------------------------------------------------------------------------------
struct Doc {
    std::string actorId;
    uint32_t seq = 0, maxOp = 0;
    std::vector<std::pair<std::string, uint32_t>> clock;   // insertion ordered Record<string, number>
    std::map<uint64_t, Object> objects;                    // objects + metadata keyed by ObjectId

    explicit Doc(const std::string& a) : actorId(a) { objects[ROOT.key()] = Object(); }

    uint32_t clockGet(const std::string& a) const { for (auto& kv : clock) if (kv.first == a) return kv.second; return 0; }
    void clockSet(const std::string& a, uint32_t v) { for (auto& kv : clock) if (kv.first == a) { kv.second = v; return; } clock.push_back({a, v}); }
    Object* objGet(const OpId& id) { auto it = objects.find(id.key()); return it == objects.end() ? nullptr : &it->second; }

    // getObjectIdForPath (:446-463)
    OpId getObjectIdForPath(const std::vector<std::string>& path) {
        OpId objectId = ROOT;
        for (auto& pathElem : path) {
            Object* meta = objGet(objectId);
            if (!meta) rangeError("No object at path");
            if (meta->isList) rangeError("Object " + pathElem + " in path is a list");
            OpId* child = Object::find(meta->children, pathElem);
            if (!child) error("Child not found: " + pathElem + " in " + opIdStr(objectId));
            objectId = *child;
        }
        return objectId;
    }

    // findListElement (:731-755)
    void findListElement(const OpId& objectId, const OpId& elemId, size_t& index, uint32_t& visible) {
        index = 0; visible = 0;
        Object* o = objGet(objectId);
        if (!o) error("Object ID not found: " + opIdStr(objectId));
        if (!o->isList) error("Expected array metadata for findListElement");
        auto& meta = o->meta;
        while (index < meta.size() && meta[index].elemId != elemId) {   // :747-750
            if (!meta[index].deleted) visible++;
            index++;
        }
        if (index == meta.size()) rangeError("List element not found: " + opIdStr(elemId));
    }

    // applyListInsert (:614-672)
    void applyListInsert(const Op& op, std::vector<Patch>* patches) {
        Object* list = objGet(op.obj);
        if (!list || !list->isList) error("Not a list: " + opIdStr(op.obj));
        auto& metadata = list->meta;
        int64_t index; uint32_t visible = 0;
        if (op.elemId == HEAD) { index = -1; visible = 0; }
        else { size_t i; findListElement(op.obj, op.elemId, i, visible); index = (int64_t)i; }
        if (index >= 0 && !metadata[index].deleted) visible++;                          // :623-625
        index++;
        while ((size_t)index < metadata.size() && compareOpIds(op.opId, metadata[index].elemId) < 0) {  // :630-635
            if (!metadata[index].deleted) visible++;
            index++;
        }
        Elem e; e.elemId = op.opId; e.valueId = op.opId; e.deleted = false;
        metadata.insert(metadata.begin() + index, e);                                    // :638
        if (!op.value.isStr()) error("Expected value inserted into text to be a string"); // :654
        list->text.insert(list->text.begin() + visible, op.value.str);                   // :657
        if (patches) {
            Patch p; p.kind = Patch::INSERT; p.index = visible; p.value = op.value.str;
            p.marks = getActiveMarksAtIndex(metadata, (size_t)index);                    // :659
            patches->push_back(p);
        }
    }

    // applyListUpdate (:677-724)
    void applyListUpdate(const Op& op, std::vector<Patch>* patches) {
        size_t index; uint32_t visible;
        findListElement(op.obj, op.elemId, index, visible);
        Object* list = objGet(op.obj);
        Elem& meta = list->meta[index];
        if (op.action == A_DEL) {
            if (!meta.deleted) {
                meta.deleted = true;                                                     // :694
                list->text.erase(list->text.begin() + visible);                          // :695
                if (patches) { Patch p; p.kind = Patch::DELETE; p.index = visible; patches->push_back(p); }
            }
        } else if (compareOpIds(meta.valueId, op.opId) < 0) {
            error("Not implemented yet");                                                // :706
        }
    }

    // applyOp (:534-608)
    void applyOp(const Op& op, std::vector<Patch>* patches) {
        Object* metadata = objGet(op.obj);
        if (!metadata) rangeError("Object does not exist: " + opIdStr(op.obj));          // :538-540
        if (op.action == A_MAKEMAP) { Object o; o.isList = false; objects[op.opId.key()] = o; }   // :541-543
        else if (op.action == A_MAKELIST) { Object o; o.isList = true; objects[op.opId.key()] = o; }  // :544-547
        metadata = objGet(op.obj);
        if (metadata->isList) {
            if (op.action == A_SET) {
                if (!op.hasElemId) error("Must specify elemId when calling set on an array");
                applyListInsert(op, patches);
            } else if (op.action == A_DEL) {
                if (!op.hasElemId) error("Must specify elemId when calling del on an array");
                applyListUpdate(op, patches);
            } else if (op.action == A_ADDMARK || op.action == A_REMOVEMARK) {
                applyAddRemoveMark(op.mark.get(), *metadata, patches);                   // :565
                heldMarks.push_back(op.mark);   // keep the op object alive: sets hold it by identity
            } else {
                error("Unimplemented");                                                  // :567
            }
        } else {
            if (op.action == A_ADDMARK || op.action == A_REMOVEMARK) error("Can't call addMark or removeMark on a map");
            if (!op.hasKey) error("Must specify key when calling set or del on a map");
            OpId* keyMeta = Object::find(metadata->keyMeta, op.key);
            if (!keyMeta || compareOpIds(*keyMeta, op.opId) == -1) {                     // :585
                Object::put(metadata->keyMeta, op.key, op.opId);
                if (op.action == A_DEL) {                                                // :588 (CHILDREN entry is NOT removed)
                    auto& f = metadata->fields;
                    for (size_t i = 0; i < f.size(); i++) if (f[i].first == op.key) { f.erase(f.begin() + i); break; }
                } else if (op.action == A_MAKELIST) {                                    // :589-592
                    MapEntry e; e.isChild = true; e.child = op.opId;
                    Object::put(metadata->fields, op.key, e);
                    Object::put(metadata->children, op.key, op.opId);
                    if (patches) { Patch p; p.kind = Patch::MAKELIST; p.op = op; patches->push_back(p); }
                } else if (op.action == A_MAKEMAP) {                                     // :593-596 (no patch: reference BUG note)
                    MapEntry e; e.isChild = true; e.child = op.opId;
                    Object::put(metadata->fields, op.key, e);
                    Object::put(metadata->children, op.key, op.opId);
                } else if (op.action == A_SET) {                                         // :597-598
                    MapEntry e; e.isChild = false; e.prim = op.value;
                    Object::put(metadata->fields, op.key, e);
                }
            }
        }
    }
    std::vector<std::shared_ptr<MarkOp>> heldMarks;

    // applyChange (:499-514)
    struct Change { std::string actor; uint32_t seq = 0; std::vector<std::pair<std::string, uint32_t>> deps; uint32_t startOp = 0; std::vector<Op> ops; };
    void applyChange(const Change& change, std::vector<Patch>* patches) {
        uint32_t lastSeq = clockGet(change.actor);
        if (change.seq != lastSeq + 1)
            rangeError("Expected sequence number " + std::to_string(lastSeq + 1) + ", got " + std::to_string(change.seq));
        for (auto& d : change.deps) {
            uint32_t have = clockGet(d.first);
            if (!have || have < d.second) rangeError("Missing dependency: change " + std::to_string(d.second) + " by actor " + d.first);
        }
        clockSet(change.actor, change.seq);
        maxOp = std::max(maxOp, change.startOp + (uint32_t)change.ops.size() - 1);
        for (auto& op : change.ops) applyOp(op, patches);
    }

    // makeNewOp (:483-493)
    OpId makeNewOp(Change& change, Op op, std::vector<Patch>* patches) {
        maxOp += 1;
        op.opId.ctr = maxOp; op.opId.actor = g_actors.intern(actorId);
        if (op.mark) op.mark->opId = op.opId;
        applyOp(op, patches);
        change.ops.push_back(op);
        return op.opId;
    }

    // changeMark (src/peritext.ts:458-501)
    std::shared_ptr<MarkOp> changeMark(const Value& inputOp, const OpId& objId, Object& list) {
        auto m = std::make_shared<MarkOp>();
        m->obj = objId;
        m->add = inputOp.get("action")->str == "addMark";
        m->markType = markTypeOf(inputOp.get("markType")->str);
        int64_t startIndex = (int64_t)inputOp.get("startIndex")->num, endIndex = (int64_t)inputOp.get("endIndex")->num;
        const bool endGrows = kInclusive[m->markType];                                   // :467 (startGrows = false :466)
        m->start.type = B_BEFORE; m->start.elemId = getListElementId(list.meta, startIndex, false);   // :488
        if (endGrows && endIndex >= (int64_t)list.text.size()) { m->end.type = B_END; }  // :491-492
        else if (endGrows) { m->end.type = B_BEFORE; m->end.elemId = getListElementId(list.meta, endIndex, false); }   // :494
        else { m->end.type = B_AFTER; m->end.elemId = getListElementId(list.meta, endIndex - 1, false); }             // :496
        setMarkAttrs(*m, inputOp.get("attrs"));
        return m;
    }

    // change (:308-441)
    Change change(const Value& inputOps, std::vector<Patch>* patches) {
        Change change;
        change.deps = clock;                                                             // :314
        seq += 1; clockSet(actorId, seq);                                                // :318-319
        change.actor = actorId; change.seq = seq; change.startOp = maxOp + 1;
        for (size_t k = 0; k < inputOps.size(); k++) {
            const Value& inputOp = inputOps.at(k);
            std::vector<std::string> path;
            if (const Value* p = inputOp.get("path")) for (size_t i = 0; i < p->size(); i++) path.push_back(p->at(i).str);
            OpId objId = getObjectIdForPath(path);
            Object* obj = objGet(objId);
            if (!obj) error("Object doesn't exist: " + opIdStr(objId));
            const std::string action = inputOp.get("action") ? inputOp.get("action")->str : "";
            if (obj->isList) {
                if (action == "insert") {
                    int64_t index = (int64_t)inputOp.get("index")->num;
                    OpId elemId = index == 0 ? HEAD : getListElementId(obj->meta, index - 1, true);   // :347-350
                    const Value* values = inputOp.get("values");
                    for (size_t i = 0; values && i < values->size(); i++) {
                        Op op; op.action = A_SET; op.obj = objId; op.hasElemId = true; op.elemId = elemId; op.insert = true; op.value = values->at(i);
                        elemId = makeNewOp(change, op, patches);                         // :352-359
                    }
                } else if (action == "delete") {
                    int64_t index = (int64_t)inputOp.get("index")->num, count = (int64_t)inputOp.get("count")->num;
                    for (int64_t i = 0; i < count; i++) {
                        obj = objGet(objId);
                        Op op; op.action = A_DEL; op.obj = objId; op.hasElemId = true; op.elemId = getListElementId(obj->meta, index, false);  // :385
                        makeNewOp(change, op, patches);
                    }
                } else if (action == "addMark" || action == "removeMark") {
                    Op op; op.action = action == "addMark" ? A_ADDMARK : A_REMOVEMARK; op.obj = objId;
                    op.mark = changeMark(inputOp, objId, *obj);                          // :394
                    makeNewOp(change, op, patches);
                } else if (action == "del") error("Use the remove action");
                else error("Unimplemented");
            } else {
                if (action == "makeList" || action == "makeMap" || action == "del") {    // :406-418
                    Op op; op.action = action == "makeList" ? A_MAKELIST : action == "makeMap" ? A_MAKEMAP : A_DEL;
                    op.obj = objId; op.hasKey = true; op.key = inputOp.get("key")->str;
                    makeNewOp(change, op, patches);
                } else if (action == "set") {
                    Op op; op.action = A_SET; op.obj = objId; op.hasKey = true; op.key = inputOp.get("key")->str;
                    if (const Value* v = inputOp.get("value")) op.value = *v;
                    makeNewOp(change, op, patches);
                } else error("Not a list: " + (path.empty() ? std::string("") : path[0]));   // :433
            }
        }
        return change;
    }

    Value objectJson(const OpId& id) {
        Object* o = objGet(id);
        if (!o) return Value();
        if (o->isList) { Value a = Value::array(); for (auto& s : o->text) a.push(Value::string(s)); return a; }
        Value m = Value::object();
        for (auto& f : o->fields) m.set(f.first, f.second.isChild ? objectJson(f.second.child) : f.second.prim);
        return m;
    }
};

static Value changeJson(const Doc::Change& c) {
    Value o = Value::object();
    o.set("actor", Value::string(c.actor)); o.set("seq", Value::number(c.seq));
    Value deps = Value::object(); for (auto& d : c.deps) deps.set(d.first, Value::number(d.second));
    o.set("deps", deps); o.set("startOp", Value::number(c.startOp));
    Value ops = Value::array(); for (auto& op : c.ops) ops.push(opJson(op));
    o.set("ops", ops);
    return o;
}
static Doc::Change changeFrom(const Value& v) {
    Doc::Change c;
    c.actor = v.get("actor")->str; c.seq = (uint32_t)v.get("seq")->num; c.startOp = (uint32_t)v.get("startOp")->num;
    if (const Value* d = v.get("deps")) if (d->isObj()) for (auto& kv : *d->obj) c.deps.push_back({kv.first, (uint32_t)kv.second.num});
    const Value* ops = v.get("ops");
    for (size_t i = 0; ops && i < ops->size(); i++) c.ops.push_back(opFrom(ops->at(i)));
    return c;
}
static Value spansJson(const std::vector<Span>& spans) {
    Value a = Value::array();
    for (auto& s : spans) { Value o = Value::object(); o.set("marks", markMapJson(s.marks)); o.set("text", Value::string(s.text)); a.push(o); }
    return a;
}
static Value patchesJson(const std::vector<Patch>& ps) { Value a = Value::array(); for (auto& p : ps) a.push(patchJson(p)); return a; }

}  // namespace po

// =================================================================================================
// C interface (ctypes).  Strings returned are malloc'd; free with po_free.  Failures return a string
// starting with "!<ErrorKind>:<message>" (mirrors the thrown JS error).
// =================================================================================================
using namespace po;

static char* dupstr(const std::string& s) { char* p = (char*)malloc(s.size() + 1); memcpy(p, s.c_str(), s.size() + 1); return p; }
template <class F> static char* guarded(F f) {
    try { return dupstr(f()); }
    catch (const JsError& e) { return dupstr("!" + e.kind + ":" + e.what()); }
    catch (const std::exception& e) { return dupstr(std::string("!Error:") + e.what()); }
}

extern "C" {

void po_free(char* p) { free(p); }
void* po_doc_new(const char* actorId) { return new Doc(actorId); }
void po_doc_free(void* d) { delete (Doc*)d; }

char* po_doc_change(void* d, const char* inputOpsJson) {
    return guarded([&] {
        std::vector<Patch> patches;
        Value in = pjson::parse(inputOpsJson);
        Doc::Change c = ((Doc*)d)->change(in, &patches);
        Value o = Value::object(); o.set("change", changeJson(c)); o.set("patches", patchesJson(patches));
        return pjson::dump(o);
    });
}
char* po_doc_apply_change(void* d, const char* changeJsonStr) {
    return guarded([&] {
        std::vector<Patch> patches;
        ((Doc*)d)->applyChange(changeFrom(pjson::parse(changeJsonStr)), &patches);
        return pjson::dump(patchesJson(patches));
    });
}
char* po_doc_spans(void* d) {   // getTextWithFormatting(["text"]) (src/micromerge.ts:516-529)
    return guarded([&] {
        Doc* doc = (Doc*)d;
        OpId id = doc->getObjectIdForPath({"text"});
        Object* o = doc->objGet(id);
        if (!o || !o->isList) error("Expected a list at object ID " + opIdStr(id));
        return pjson::dump(spansJson(getTextWithFormatting(*o)));
    });
}
char* po_doc_root(void* d) { return guarded([&] { return pjson::dump(((Doc*)d)->objectJson(ROOT)); }); }
char* po_doc_clock(void* d) {
    return guarded([&] { Value o = Value::object(); for (auto& kv : ((Doc*)d)->clock) o.set(kv.first, Value::number(kv.second)); return pjson::dump(o); });
}
char* po_doc_get_cursor(void* d, int64_t index) {   // getCursor (src/micromerge.ts:465-473)
    return guarded([&] {
        Doc* doc = (Doc*)d; OpId id = doc->getObjectIdForPath({"text"});
        Object* o = doc->objGet(id);
        if (!o || !o->isList) error("Expected array metadata for findListElement");
        Value c = Value::object(); c.set("objectId", Value::string(opIdStr(id)));
        c.set("elemId", Value::string(opIdStr(getListElementId(o->meta, index, false))));
        return pjson::dump(c);
    });
}
char* po_doc_resolve_cursor(void* d, const char* cursorJson) {   // resolveCursor (src/micromerge.ts:475-477)
    return guarded([&] {
        Value c = pjson::parse(cursorJson);
        size_t index; uint32_t visible;
        ((Doc*)d)->findListElement(parseOpId(c.get("objectId")->str), parseOpId(c.get("elemId")->str), index, visible);
        return std::to_string(visible);
    });
}
// Dump of the element sequence (elemIds, deleted flags) for closed-form cross-checks in tests.
char* po_doc_elements(void* d) {
    return guarded([&] {
        Doc* doc = (Doc*)d; OpId id = doc->getObjectIdForPath({"text"});
        Object* o = doc->objGet(id);
        Value a = Value::array();
        for (auto& e : o->meta) { Value x = Value::object(); x.set("elemId", Value::string(opIdStr(e.elemId))); x.set("deleted", Value::boolean(e.deleted));
            x.set("before", Value::boolean((bool)e.markOpsBefore)); x.set("after", Value::boolean((bool)e.markOpsAfter)); a.push(x); }
        return pjson::dump(a);
    });
}
int po_compare_op_ids(const char* a, const char* b) {
    try { return compareOpIds(parseOpId(a), parseOpId(b)); } catch (...) { return -2; }
}

// -------------------------------------------------------------------------------------------------
// Packed replay: apply packed logs (include/peritext_b200.h) with the SAME sequential reference
// algorithm and emit results in the engine's binary result format, for bulk parity checks and for
// the CPU baseline timing.  One log = applyOp over its records in arrival order, then
// getTextWithFormatting.  Output arrays are caller-allocated with the engine's capacities.
// -------------------------------------------------------------------------------------------------
static std::once_flag g_packedNames;
static std::vector<int> g_rankActor;     // actor rank -> interned actor id whose name sorts by rank
static std::vector<int> g_attrStr;       // attr id -> interned string sorting by id
static std::mutex g_packedMu;
static void ensurePackedNames(uint32_t nActors, uint32_t nAttrs) {
    std::lock_guard<std::mutex> g(g_packedMu);
    char buf[32];
    while (g_rankActor.size() < nActors) { snprintf(buf, sizeof buf, "r%08u", (unsigned)g_rankActor.size()); g_rankActor.push_back(g_actors.intern(buf)); }
    while (g_attrStr.size() < nAttrs) { snprintf(buf, sizeof buf, "a%010u", (unsigned)g_attrStr.size()); g_attrStr.push_back(g_strings.intern(buf)); }
    activeTrueId();
}

struct PackedOut {
    pt_log_result* results; const uint64_t* text_off; const uint64_t* span_off;
    uint32_t* text; pt_span* spans; uint32_t* comment_pool; uint64_t comment_cap; std::atomic<uint64_t>* comment_used;
};

static void replayOne(const pt_packed_ops* in, uint32_t li, const PackedOut& out, bool flatten) {
    const pt_log_desc& L = in->logs[li];
    pt_log_result& R = out.results[li];
    memset(&R, 0, sizeof R);
    Doc doc("oracle");
    const OpId listId{1, g_rankActor[0]};
    { Object o; o.isList = true; doc.objects[listId.key()] = o; }
    Object::put(doc.objGet(ROOT)->children, std::string("text"), listId);
    Object& list = *doc.objGet(listId);
    std::vector<uint32_t> tokens;  // token per element, in op order; looked up by value string index
    const pt_insdel_rec* ids = in->insdel + L.insdel_off;
    const pt_mark_rec* mks = in->marks + L.mark_off;
    uint32_t mi = 0;
    auto applyMark = [&](const pt_mark_rec& r) {
        auto m = std::make_shared<MarkOp>();
        m->opId = OpId{r.ctr, g_rankActor[r.actor]}; m->obj = listId;
        m->add = (r.kind & 1) == 0; m->markType = (r.kind >> 1) & 3;
        m->start.type = r.bounds & 3; m->end.type = (r.bounds >> 2) & 3;
        if (m->start.type <= B_AFTER) m->start.elemId = OpId{r.start_ctr, g_rankActor[r.start_actor]};
        if (m->end.type <= B_AFTER) m->end.elemId = OpId{r.end_ctr, g_rankActor[r.end_actor]};
        if (r.attr != PT_ATTR_NONE) { m->hasAttrs = true; m->attrsId = g_attrStr[r.attr]; m->commentId = g_attrStr[r.attr]; }
        Op op; op.opId = m->opId; op.obj = listId; op.action = m->add ? A_ADDMARK : A_REMOVEMARK; op.mark = m;
        doc.applyOp(op, nullptr);
    };
    try {
        for (uint32_t k = 0; k < L.n_insdel; k++) {
            while (mi < L.n_mark && mks[mi].arrival <= k) applyMark(mks[mi++]);
            const pt_insdel_rec& r = ids[k];
            Op op; op.opId = OpId{r.ctr, g_rankActor[r.actor]}; op.obj = listId; op.hasElemId = true;
            op.elemId = (r.ref_ctr == 0) ? HEAD : OpId{r.ref_ctr, g_rankActor[r.ref_actor]};
            uint32_t kind = PT_PAYLOAD_KIND(r.payload);
            if (kind == PT_KIND_INSERT) {
                op.action = A_SET; op.insert = true;
                // the value string carries the 30-bit token verbatim (4 bytes), so text[] round-trips tokens
                uint32_t tok = PT_PAYLOAD_TOKEN(r.payload);
                op.value = Value::string(std::string((const char*)&tok, 4));
            } else if (kind == PT_KIND_DELETE) {
                if (op.elemId == HEAD) rangeError("List element not found: _head");
                op.action = A_DEL;
            } else { R.status = PT_LOG_BAD_KIND; return; }
            doc.applyOp(op, nullptr);
        }
        while (mi < L.n_mark) applyMark(mks[mi++]);
    } catch (const JsError& e) {
        R.status = PT_LOG_ELEM_NOT_FOUND;   // the only reachable throw for well-typed packed input (src/micromerge.ts:752)
        return;
    }
    R.n_elems = (uint32_t)list.meta.size();
    R.n_visible = (uint32_t)list.text.size();
    if (!flatten) return;
    std::vector<Span> spans = getTextWithFormatting(list);
    R.n_spans = (uint32_t)spans.size();
    uint32_t* text = out.text + out.text_off[li];
    pt_span* sp = out.spans + out.span_off[li];
    uint64_t d0 = 0, d1 = 0;
    auto addTerm = [&](uint64_t t) { d0 += t; d1 ^= pt_term_hi(t); };
    for (uint32_t i = 0; i < R.n_visible; i++) { uint32_t tok; memcpy(&tok, list.text[i].data(), 4); text[i] = tok; addTerm(pt_term_text(i, tok)); }
    uint32_t start = 0;
    for (uint32_t j = 0; j < R.n_spans; j++) {
        const MarkMap& m = spans[j].marks;
        pt_span s; s.start = start;
        uint32_t flags = 0;
        if (m.single[STRONG] >= 0) flags |= PT_SPAN_STRONG;
        if (m.single[EM] >= 0) flags |= PT_SPAN_EM;
        s.link_attr = PT_ATTR_NONE;
        if (m.single[LINK] >= 0) { flags |= PT_SPAN_LINK; uint32_t a = 0; sscanf(g_strings.name(m.single[LINK]).c_str() + 1, "%u", &a); s.link_attr = a; }
        if (m.hasComment) flags |= PT_SPAN_COMMENT;
        uint32_t nc = (uint32_t)m.comments.size();
        flags |= nc << 8;
        s.flags = flags; s.comment_off = 0;
        if (nc) {
            uint64_t off = out.comment_used->fetch_add(nc);
            if (off + nc > out.comment_cap) { R.status = PT_LOG_OVERFLOW; return; }
            s.comment_off = (uint32_t)off;
            for (uint32_t k = 0; k < nc; k++) { uint32_t a = 0; sscanf(g_strings.name(m.comments[k].first).c_str() + 1, "%u", &a); out.comment_pool[off + k] = a; addTerm(pt_term_comment(j, k, a)); }
        }
        addTerm(pt_term_span(j, s.start, s.flags, s.link_attr));
        sp[j] = s;
        start += (uint32_t)(spans[j].text.size() / 4);
    }
    addTerm(pt_term_counts(R.n_visible, R.n_spans));
    R.digest[0] = d0; R.digest[1] = d1;
}

// Replays logs [first, first+count) on `threads` host threads (one log per thread at a time).
// flatten=0 skips getTextWithFormatting (apply-only timing).  Returns 0.
int po_replay_packed(const pt_packed_ops* in, uint32_t first, uint32_t count, int threads, int flatten,
                     pt_log_result* results, const uint64_t* text_off, const uint64_t* span_off,
                     uint32_t* text, pt_span* spans, uint32_t* comment_pool, uint64_t comment_cap, uint64_t* comment_used) {
    uint32_t maxActors = 1, maxAttr = 0;
    for (uint32_t i = first; i < first + count; i++) {
        maxActors = std::max(maxActors, in->logs[i].n_actors);
        const pt_mark_rec* mk = in->marks + in->logs[i].mark_off;
        for (uint32_t k = 0; k < in->logs[i].n_mark; k++) if (mk[k].attr != PT_ATTR_NONE) maxAttr = std::max(maxAttr, mk[k].attr + 1);
    }
    ensurePackedNames(maxActors + 1, maxAttr);
    std::atomic<uint64_t> used{comment_used ? *comment_used : 0};
    PackedOut out{results, text_off, span_off, text, spans, comment_pool, comment_cap, &used};
    std::atomic<uint32_t> next{first};
    auto worker = [&] { for (;;) { uint32_t i = next.fetch_add(1); if (i >= first + count) break; replayOne(in, i, out, flatten != 0); } };
    if (threads <= 1) worker();
    else { std::vector<std::thread> ts; for (int t = 0; t < threads; t++) ts.emplace_back(worker); for (auto& t : ts) t.join(); }
    if (comment_used) *comment_used = used.load();
    return 0;
}

const char* po_version(void) { return "peritext-oracle 1 (restates inkandswitch/peritext@89c162d3)"; }

}  // extern "C"
