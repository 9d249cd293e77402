// workload.cpp — seeded synthetic trace generator for the BASELINE.json configs, host only.
//
// Produces packed op logs (include/peritext_b200.h) directly: for every document, R replicas edit concurrently and
// exchange changes; log (doc, r) is what replica r applied, in ITS arrival order.  Ops are generated the way the
// reference's Micromerge.change() does it (reference src/micromerge.ts:346-396): visible index -> elemId, including
// `lookAfterTombstones` (src/micromerge.ts:775-797) for inserts and changeMark's boundary choice
// (src/peritext.ts:458-501), Lamport counters maxOp+1 (src/micromerge.ts:487), so every log is causally valid.
// The generator keeps its own light replica state (element order + tombstones); it is NOT the oracle and is not
// used to check anything.
//
// Shapes:
//   kind 2  "insdel": 70% inserts in typing runs (geometric mean 8), 30% deletes (runs, mean 4); epochs of 32 input ops
//                     per actor with a full sync between epochs (ops inside an epoch are mutually concurrent)
//   kind 3  "marks" : kind 2 with 10% of the op records replaced by add/removeMark (35% strong, 25% em, 25% link,
//                     15% comment; 70% add; length geometric mean 64)
//   kind 4  "fuzz"  : reference test/fuzz.ts shape: one random op by one random replica, then a pairwise two-way sync of
//                     a random pair; op type uniform over insert/remove/addMark/removeMark; final full sync
//   kind 5  "long"  : long-form typing (runs mean 32) then dense overlapping marks (length mean 256)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/peritext_b200.h"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return n ? (uint32_t)(uni() * n) : 0; }
    uint32_t geom(double mean) { double p = 1.0 / mean; double u = uni(); if (u <= 0) u = 1e-300; return 1 + (uint32_t)std::floor(std::log(u) / std::log(1.0 - p)); }
};

struct Elem { uint32_t ctr; uint16_t actor; uint8_t deleted; uint8_t hasAfter; };
inline bool idLess(uint32_t c1, uint16_t a1, uint32_t c2, uint16_t a2) { return c1 < c2 || (c1 == c2 && a1 < a2); }

struct OpRec {
    bool isMark;
    pt_insdel_rec id;
    pt_mark_rec mk;
};
struct Change { uint16_t actor; uint32_t seq; uint32_t startOp; std::vector<OpRec> ops; };

struct Replica {
    std::vector<Elem> seq;
    uint32_t nvis = 0, maxOp = 0;
    std::vector<uint32_t> clock;           // per actor: changes applied
    std::vector<uint32_t> seenComments;
    std::vector<OpRec> log;                // arrival order

    size_t findId(uint32_t ctr, uint16_t actor, size_t hint) const {
        if (hint < seq.size() && seq[hint].ctr == ctr && seq[hint].actor == actor) return hint;
        for (size_t i = 0; i < seq.size(); i++) if (seq[i].ctr == ctr && seq[i].actor == actor) return i;
        return (size_t)-1;
    }
    size_t posOfVisible(uint32_t k) const {   // position of the k-th visible element
        uint32_t v = 0;
        for (size_t i = 0; i < seq.size(); i++) if (!seq[i].deleted) { if (v == k) return i; v++; }
        return (size_t)-1;
    }
    // applyOp for list ops (reference src/micromerge.ts:614-724, src/peritext.ts:154): order + tombstones + defined-after flags
    size_t lastPos = (size_t)-1;
    void apply(const OpRec& op) {
        log.push_back(op);
        if (op.isMark) {
            uint32_t eb = (op.mk.bounds >> 2) & 3u;
            if (eb == PT_BOUND_AFTER) { size_t p = findId(op.mk.end_ctr, op.mk.end_actor, (size_t)-1); if (p != (size_t)-1) seq[p].hasAfter = 1; }
            maxOp = std::max(maxOp, op.mk.ctr);
            return;
        }
        const pt_insdel_rec& r = op.id;
        maxOp = std::max(maxOp, r.ctr);
        if (PT_PAYLOAD_KIND(r.payload) == PT_KIND_INSERT) {
            size_t pos;
            if (r.ref_ctr == 0) pos = 0;
            else { size_t p = findId(r.ref_ctr, r.ref_actor, lastPos); pos = p + 1; }
            while (pos < seq.size() && idLess(r.ctr, r.actor, seq[pos].ctr, seq[pos].actor)) pos++;   // :630-635
            seq.insert(seq.begin() + pos, Elem{r.ctr, r.actor, 0, 0});
            nvis++; lastPos = pos;
        } else {
            size_t p = findId(r.ref_ctr, r.ref_actor, lastPos == (size_t)-1 ? lastPos : lastPos + 1);
            if (p != (size_t)-1 && !seq[p].deleted) { seq[p].deleted = 1; nvis--; }
            lastPos = p;
        }
    }
};

struct DocGen {
    uint32_t R; Rng rng; uint32_t kind; uint32_t docId;
    std::vector<Replica> reps;
    std::vector<std::vector<Change>> queues;   // per actor
    uint32_t uniqueOps = 0;
    uint32_t nextComment = 0;

    DocGen(uint32_t R_, uint64_t seed, uint32_t kind_, uint32_t docId_) : R(R_), rng(seed), kind(kind_), docId(docId_), reps(R_), queues(R_) {
        for (auto& r : reps) r.clock.assign(R, 0);
    }

    Change begin(uint32_t a) { Change c; c.actor = (uint16_t)a; c.seq = reps[a].clock[a] + 1; c.startOp = reps[a].maxOp + 1; return c; }
    void commit(uint32_t a, Change& c) {
        if (c.ops.empty()) return;
        reps[a].clock[a] = c.seq; uniqueOps += (uint32_t)c.ops.size();
        queues[a].push_back(std::move(c));
    }
    void localOp(uint32_t a, Change& c, OpRec op) { reps[a].apply(op); c.ops.push_back(op); }

    // insert `len` characters at visible index idx (reference src/micromerge.ts:346-361)
    void genInsert(uint32_t a, Change& c, uint32_t idx, uint32_t len, bool hex) {
        Replica& rp = reps[a];
        uint32_t rc = 0; uint16_t ra = 0;
        if (idx > 0) {
            size_t p = rp.posOfVisible(idx - 1);
            size_t e = p, peek = p + 1, latest = 0;
            while (peek < rp.seq.size() && rp.seq[peek].deleted) { if (rp.seq[peek].hasAfter) latest = peek; peek++; }   // :788-796
            if (latest) e = latest;
            rc = rp.seq[e].ctr; ra = rp.seq[e].actor; rp.lastPos = e;
        }
        for (uint32_t k = 0; k < len; k++) {
            OpRec op{}; op.isMark = false;
            op.id.ctr = rp.maxOp + 1; op.id.actor = (uint16_t)a; op.id.ref_ctr = rc; op.id.ref_actor = ra;
            uint32_t ch = hex ? (uint32_t)"0123456789abcdef"[rng.below(16)] : (uint32_t)"abcdefghijklmnopqrstuvwxyz "[rng.below(27)];
            op.id.payload = (PT_KIND_INSERT << 30) | ch;
            localOp(a, c, op);
            rc = op.id.ctr; ra = op.id.actor;
        }
    }
    // delete `count` characters at visible index idx (reference src/micromerge.ts:362-392)
    void genDelete(uint32_t a, Change& c, uint32_t idx, uint32_t count) {
        Replica& rp = reps[a];
        size_t p = rp.posOfVisible(idx);
        for (uint32_t k = 0; k < count && p != (size_t)-1 && p < rp.seq.size(); k++) {
            while (p < rp.seq.size() && rp.seq[p].deleted) p++;
            if (p >= rp.seq.size()) break;
            OpRec op{}; op.isMark = false;
            op.id.ctr = rp.maxOp + 1; op.id.actor = (uint16_t)a; op.id.ref_ctr = rp.seq[p].ctr; op.id.ref_actor = rp.seq[p].actor;
            op.id.payload = (PT_KIND_DELETE << 30);
            rp.lastPos = p == 0 ? (size_t)-1 : p - 1;
            localOp(a, c, op);
        }
    }
    // add/removeMark over visible [start, end) (reference src/peritext.ts:458-501)
    void genMark(uint32_t a, Change& c, bool add, uint32_t type, uint32_t start, uint32_t end) {
        Replica& rp = reps[a];
        OpRec op{}; op.isMark = true;
        pt_mark_rec& m = op.mk;
        m.ctr = rp.maxOp + 1; m.actor = (uint16_t)a; m.kind = (uint8_t)((add ? 0 : 1) | (type << 1));
        size_t ps = rp.posOfVisible(start);
        m.start_ctr = rp.seq[ps].ctr; m.start_actor = rp.seq[ps].actor;
        uint32_t sb = PT_BOUND_BEFORE, eb;
        const bool inclusive = type == PT_MARK_STRONG || type == PT_MARK_EM;
        if (inclusive && end >= rp.nvis) { eb = PT_BOUND_END_OF_TEXT; }
        else if (inclusive) { size_t pe = rp.posOfVisible(end); eb = PT_BOUND_BEFORE; m.end_ctr = rp.seq[pe].ctr; m.end_actor = rp.seq[pe].actor; }
        else { size_t pe = rp.posOfVisible(end - 1); eb = PT_BOUND_AFTER; m.end_ctr = rp.seq[pe].ctr; m.end_actor = rp.seq[pe].actor; }
        m.bounds = (uint8_t)(sb | (eb << 2));
        m.attr = PT_ATTR_NONE;
        if (type == PT_MARK_LINK && add) m.attr = rng.below(26);
        if (type == PT_MARK_COMMENT) {
            if (add) { m.attr = docId * 4096u + (nextComment++ & 4095u); rp.seenComments.push_back(m.attr); }
            else {
                if (rp.seenComments.empty()) return;
                m.attr = rp.seenComments[rng.below((uint32_t)rp.seenComments.size())];
            }
        }
        localOp(a, c, op);
    }

    // deliver every change `dst` is missing from `src`'s knowledge, in a causal order (ascending startOp, actor)
    void sync(uint32_t src, uint32_t dst) {
        std::vector<const Change*> missing;
        for (uint32_t a = 0; a < R; a++)
            for (uint32_t s = reps[dst].clock[a]; s < reps[src].clock[a]; s++) missing.push_back(&queues[a][s]);
        std::sort(missing.begin(), missing.end(), [](const Change* x, const Change* y) { return x->startOp < y->startOp || (x->startOp == y->startOp && x->actor < y->actor); });
        for (const Change* ch : missing) {
            reps[dst].lastPos = (size_t)-1;
            for (const OpRec& op : ch->ops) {
                reps[dst].apply(op);
                if (op.isMark && ((op.mk.kind >> 1) & 3u) == PT_MARK_COMMENT && (op.mk.kind & 1u) == 0) reps[dst].seenComments.push_back(op.mk.attr);
            }
            reps[dst].clock[ch->actor] = ch->seq;
        }
    }
    void fullSync() { for (int pass = 0; pass < 2; pass++) for (uint32_t a = 0; a < R; a++) for (uint32_t b = 0; b < R; b++) if (a != b) sync(a, b); }

    void initial(const char* text) {
        Change c = begin(0);
        Replica& rp = reps[0];
        // makeList is op 1@actor0 (not packed: it targets ROOT); the characters follow (test/generateDocs.ts:26-34)
        rp.maxOp = 1;
        c.startOp = 1;
        uint32_t rc = 0; uint16_t ra = 0;
        for (const char* p = text; *p; p++) {
            OpRec op{}; op.isMark = false;
            op.id.ctr = rp.maxOp + 1; op.id.actor = 0; op.id.ref_ctr = rc; op.id.ref_actor = ra; op.id.payload = (PT_KIND_INSERT << 30) | (uint32_t)(unsigned char)*p;
            localOp(0, c, op); rc = op.id.ctr; ra = 0;
        }
        commit(0, c);
        for (uint32_t b = 1; b < R; b++) sync(0, b);
    }

    void inputOpInsDel(uint32_t a, Change& c, double runIns, double runDel) {
        Replica& rp = reps[a];
        if (rp.nvis == 0 || rng.uni() < 0.538) genInsert(a, c, rng.below(rp.nvis + 1), rng.geom(runIns), false);
        else { uint32_t st = rng.below(rp.nvis); genDelete(a, c, st, std::min(rng.geom(runDel), rp.nvis - st)); }
    }
    void inputOpMark(uint32_t a, Change& c, double meanLen) {
        Replica& rp = reps[a];
        if (rp.nvis == 0) { genInsert(a, c, 0, 1, false); return; }
        double u = rng.uni();
        uint32_t type = u < 0.35 ? PT_MARK_STRONG : u < 0.60 ? PT_MARK_EM : u < 0.85 ? PT_MARK_LINK : PT_MARK_COMMENT;
        bool add = rng.uni() < 0.7;
        uint32_t st = rng.below(rp.nvis), len = std::min(rng.geom(meanLen), rp.nvis - st);
        genMark(a, c, add, type, st, st + std::max(1u, len));
    }

    void runEpochs(uint32_t targetOps, double markInputProb, double runIns, double runDel, double markLen) {
        initial("ABCDE");
        while (uniqueOps < targetOps) {
            for (uint32_t a = 0; a < R && uniqueOps < targetOps; a++) {
                for (int k = 0; k < 32 && uniqueOps < targetOps; k++) {
                    Change c = begin(a);
                    if (markInputProb > 0 && rng.uni() < markInputProb) inputOpMark(a, c, markLen); else inputOpInsDel(a, c, runIns, runDel);
                    // do not overshoot the target by a whole run
                    commit(a, c);
                }
            }
            fullSync();
        }
        fullSync();
    }

    void runFuzz(uint32_t targetOps) {   // reference test/fuzz.ts:167-199
        initial("ABCDE");
        while (uniqueOps < targetOps) {
            uint32_t a = rng.below(R);
            Replica& rp = reps[a];
            Change c = begin(a);
            uint32_t t = rng.below(4);
            uint32_t len = rp.nvis;
            if (len == 0) t = 0;
            if (t == 0) { uint32_t idx = len ? rng.below(len) : 0; uint32_t nch = len ? 2 * rng.below(2) : 2; genInsert(a, c, idx, nch, true); }   // fuzz.ts:109-113
            else if (t == 1) { uint32_t idx = rng.below(len) + 1; uint32_t cnt = (uint32_t)std::ceil(rng.uni() * (double)(len - idx)); if (idx < len && cnt) genDelete(a, c, idx, cnt); }  // :128-129
            else { uint32_t st = rng.below(len); uint32_t en = st + rng.below(len - st) + 1; genMark(a, c, t == 2, rng.below(4), st, en); }   // :34-36
            commit(a, c);
            uint32_t l = rng.below(R), r = rng.below(R);
            while (R > 1 && r == l) r = rng.below(R);
            if (R > 1) { sync(l, r); sync(r, l); }
        }
        fullSync();
    }

    void runLong(uint32_t targetChars, uint32_t nMarks) {
        initial("ABCDE");
        // long-form typing: mostly appends near a moving cursor, occasional jumps; 10% of the records are deletes
        uint32_t a = 0;
        uint32_t cursor[8] = {5, 5, 5, 5, 5, 5, 5, 5};
        while (reps[0].nvis < targetChars) {
            for (a = 0; a < R; a++) {
                for (int k = 0; k < 32; k++) {
                    Replica& rp = reps[a];
                    Change c = begin(a);
                    if (rng.uni() < 0.05) cursor[a] = rng.below(rp.nvis + 1);
                    if (cursor[a] > rp.nvis) cursor[a] = rp.nvis;
                    if (rp.nvis > 8 && rng.uni() < 0.27) { uint32_t cnt = std::min(rng.geom(4), cursor[a]); if (cnt) { genDelete(a, c, cursor[a] - cnt, cnt); cursor[a] -= cnt; } }
                    else { uint32_t len = rng.geom(32); genInsert(a, c, cursor[a], len, false); cursor[a] += len; }
                    commit(a, c);
                }
            }
            fullSync();
        }
        uint32_t made = 0;
        while (made < nMarks) {
            for (a = 0; a < R && made < nMarks; a++)
                for (int k = 0; k < 32 && made < nMarks; k++) { Change c = begin(a); inputOpMark(a, c, 256.0); made += (uint32_t)c.ops.size(); commit(a, c); }
            fullSync();
        }
        fullSync();
    }
};

struct DocOut { std::vector<pt_insdel_rec> insdel; std::vector<pt_mark_rec> marks; std::vector<pt_log_desc> desc; };

}  // namespace

extern "C" {

typedef struct ptw_config {
    uint32_t kind;         /* 2 insdel, 3 marks, 4 fuzz, 5 long */
    uint32_t n_docs;
    uint32_t doc_first;    /* ids of the generated docs: [doc_first, doc_first + n_docs)  (sharding across ranks) */
    uint32_t ops_per_doc;  /* unique internal op records per document (kind 5: visible characters) */
    uint32_t replicas;
    uint32_t n_marks;      /* kind 5: mark ops per doc */
    uint64_t seed;
    uint32_t threads;
    uint32_t reserved;
} ptw_config;

typedef struct ptw_batch {
    uint32_t n_logs;
    pt_log_desc* desc;
    pt_insdel_rec* insdel; uint64_t n_insdel;
    pt_mark_rec* marks; uint64_t n_marks;
    uint64_t unique_ops;   /* sum over docs of unique op records (one replica's worth) */
} ptw_batch;

int ptw_generate(const ptw_config* cfg, ptw_batch** out) {
    if (!cfg || !out || cfg->replicas == 0 || cfg->replicas > 8) return 1;
    const uint32_t nd = cfg->n_docs, R = cfg->replicas;
    std::vector<DocOut> outs(nd);
    std::vector<uint32_t> uniq(nd, 0);
    std::atomic<uint32_t> next{0};
    auto worker = [&] {
        for (;;) {
            uint32_t d = next.fetch_add(1);
            if (d >= nd) break;
            uint32_t docId = cfg->doc_first + d;
            DocGen g(R, cfg->seed ^ (0x9E3779B97F4A7C15ull * (uint64_t)(docId + 1)), cfg->kind, docId);
            switch (cfg->kind) {
                case 2: g.runEpochs(cfg->ops_per_doc, 0.0, 8.0, 4.0, 64.0); break;
                case 3: g.runEpochs(cfg->ops_per_doc, 0.406, 8.0, 4.0, 64.0); break;
                case 4: g.runFuzz(cfg->ops_per_doc); break;
                case 5: g.runLong(cfg->ops_per_doc, cfg->n_marks); break;
                default: return;
            }
            uniq[d] = g.uniqueOps;
            DocOut& o = outs[d];
            for (uint32_t r = 0; r < R; r++) {
                pt_log_desc L{}; L.insdel_off = o.insdel.size(); L.mark_off = o.marks.size();
                uint32_t maxc = 1, nid = 0;
                for (const OpRec& op : g.reps[r].log) {
                    if (op.isMark) { pt_mark_rec m = op.mk; m.arrival = nid; m.reserved = 0; o.marks.push_back(m); maxc = std::max(maxc, m.ctr); }
                    else { o.insdel.push_back(op.id); nid++; maxc = std::max(maxc, op.id.ctr); }
                }
                L.n_insdel = (uint32_t)(o.insdel.size() - L.insdel_off); L.n_mark = (uint32_t)(o.marks.size() - L.mark_off);
                L.n_actors = R; L.max_ctr = maxc;
                o.desc.push_back(L);
            }
        }
    };
    uint32_t nt = cfg->threads ? cfg->threads : std::max(1u, std::thread::hardware_concurrency());
    nt = std::min(nt, std::max(1u, nd));
    std::vector<std::thread> ts;
    for (uint32_t t = 0; t < nt; t++) ts.emplace_back(worker);
    for (auto& t : ts) t.join();

    ptw_batch* b = (ptw_batch*)calloc(1, sizeof(ptw_batch));
    uint64_t ni = 0, nm = 0, uo = 0;
    for (uint32_t d = 0; d < nd; d++) { ni += outs[d].insdel.size(); nm += outs[d].marks.size(); uo += uniq[d]; }
    b->n_logs = nd * R; b->n_insdel = ni; b->n_marks = nm; b->unique_ops = uo;
    b->desc = (pt_log_desc*)malloc(std::max<size_t>(1, b->n_logs) * sizeof(pt_log_desc));
    b->insdel = (pt_insdel_rec*)malloc(std::max<uint64_t>(1, ni) * sizeof(pt_insdel_rec));
    b->marks = (pt_mark_rec*)malloc(std::max<uint64_t>(1, nm) * sizeof(pt_mark_rec));
    uint64_t io = 0, mo = 0; uint32_t li = 0;
    for (uint32_t d = 0; d < nd; d++) {
        DocOut& o = outs[d];
        for (auto L : o.desc) { L.insdel_off += io; L.mark_off += mo; b->desc[li++] = L; }
        if (!o.insdel.empty()) memcpy(b->insdel + io, o.insdel.data(), o.insdel.size() * sizeof(pt_insdel_rec));
        if (!o.marks.empty()) memcpy(b->marks + mo, o.marks.data(), o.marks.size() * sizeof(pt_mark_rec));
        io += o.insdel.size(); mo += o.marks.size();
        DocOut().insdel.swap(o.insdel); DocOut().marks.swap(o.marks);
    }
    *out = b;
    return 0;
}

void ptw_free(ptw_batch* b) {
    if (!b) return;
    free(b->desc); free(b->insdel); free(b->marks); free(b);
}

}  // extern "C"
