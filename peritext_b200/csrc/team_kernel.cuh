// team_kernel.cuh — op-log apply + flatten for MEDIUM logs WITHOUT mark ops: a team of TEAM warps (one CTA) per log (sm_100a).
//
// BASELINE.json configs[1] ("1K docs x 10K ops, insert/delete only"): logs of ~10K records whose working set is the id table
// plus per-word bitmaps and per-run arrays.  This kernel is the warp-per-log kernel's lean pipeline (warp_kernel.cuh: 16-bit
// state, one uint4 of bit state per 32 records, run heads ranked by a key-space bitmap and threaded in ascending key order with
// match_any, Euler tour + splitter ranking of the visible weights only) cut for a CTA of 8 warps: ~50 KB of shared memory per
// log instead of the CTA-per-log kernel's ~110 KB, so FOUR logs are in flight per SM instead of two and a barrier stall of one
// team is covered by the other three.  Same closed form, same results (reference src/micromerge.ts:534-724, src/peritext.ts:337-455
// for a document without marks: one span {}).  Logs with mark ops stay on merge_kernel.cuh.
#pragma once
#include "warp_kernel.cuh"

namespace ptk {

template <int TEAM>
struct TeamCtx {
    uint32_t wa[TEAM];          // per-warp totals of the packed scans
    uint32_t lastK[3][TEAM];    // key of every warp's last record of a trip if it is an insert (chain-link detection); three trips:
                                // a warp may already publish trip t+1 while another still reads trips t and t-1
    uint32_t carry, total;
    uint32_t status, n_ins, occ;
    unsigned long long dig0, dig1;
    uint32_t work_next;
};

// exclusive scan of a packed pair (two 16-bit counters in one word) over the CTA's threads; returns the exclusive prefix, total in ctx.total
template <int TEAM>
__device__ __forceinline__ uint32_t team_scan(uint32_t v, TeamCtx<TEAM>& c) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t inc = warp_incl_scan(v, lane);
    if (lane == 31) c.wa[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t w = lane < TEAM ? c.wa[lane] : 0u;
        const uint32_t s = warp_incl_scan(w, lane);
        if (lane < TEAM) c.wa[lane] = s - w;
        if (lane == 31) c.total = s;
    }
    __syncthreads();
    const uint32_t ex = c.wa[warp] + inc - v;
    __syncthreads();
    return ex;
}

// returns 0: done, 1: defer (does not fit / not eligible)
template <int TEAM>
__device__ int team_merge_one_log(const BatchParams& P, const uint32_t li, TeamCtx<TEAM>& c) {
    constexpr uint32_t NT = TEAM * 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const pt_log_desc L = P.desc[li];
    const uint32_t n = L.n_insdel, R = L.n_actors ? L.n_actors : 1u, C = L.max_ctr;
    const unsigned long long KS64 = (unsigned long long)C * R;
    if (L.n_mark != 0 || KS64 >= 0xFFFFull || n >= 0xFFFFu) return 1;
    const uint32_t KS = (uint32_t)KS64;
    const pt_insdel_rec* __restrict__ ins = P.insdel + L.insdel_off;
    uint32_t* text_out = P.text + P.text_off[li];
    pt_log_result* res = P.results + li;
    auto keyOf = [&](uint32_t ctr, uint32_t actor) -> uint32_t { return (ctr - 1u) * R + actor; };
    auto badId = [&](uint32_t ctr, uint32_t actor) -> bool { return ctr - 1u >= C || actor >= R; };
    uint32_t st = 0;
    auto fail = [&](uint32_t code) { st = max(st, code); };

    WArena A; A.base = 0; A.used = 0; A.cap = P.smem_arena_bytes;
    const uint32_t NWr = (n + 31) / 32 + 1;
    uint4* WI = A.alloc<uint4>(NWr);            // per word: insert bits, chain-link -> head bits, visible bits, heads before | visible before << 16
    uint16_t* T = A.alloc<uint16_t>(KS);        // opId key -> insert record
    const uint32_t markC = A.used;
    uint2* OD = A.alloc<uint2>(NWr);            // other-child bits, tombstone bits (dead after C)
    if (!A.fits()) return 1;
    if (tid == 0) { c.status = 0; c.n_ins = 0; c.occ = 0; c.dig0 = 0; c.dig1 = 0; c.carry = 0; }
    {
        const uint4 f = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), z = make_uint4(0, 0, 0, 0);
        uint4* tv = reinterpret_cast<uint4*>(T);
        for (uint32_t v = tid; v < (KS * 2u + 15u) / 16u; v += NT) tv[v] = f;
        uint4* ov = reinterpret_cast<uint4*>(OD);
        for (uint32_t v = tid; v < (NWr * 8u + 15u) / 16u; v += NT) ov[v] = z;
        if (tid == 0) WI[NWr - 1] = z;
    }
    __syncthreads();

    // ---- A+B: TEAM*32 records per trip; loads three trips ahead in registers, an L2 prefetch stream eight trips ahead ----------
    {
        const uint32_t nm1 = n ? n - 1u : 0u;
        uint4 ra = make_uint4(0, 0, 0, 0), rb = ra, rc = ra;
        if (n) { ra = ld_rec(ins + min(tid, nm1)); rb = ld_rec(ins + min(NT + tid, nm1)); rc = ld_rec(ins + min(2u * NT + tid, nm1)); }
        const char* insb = reinterpret_cast<const char*>(ins);
        const uint32_t insBytes = n * 16u;
        for (uint32_t o = 3u * NT * 16u + tid * 128u; o < min(insBytes, 11u * NT * 16u); o += NT * 128u) prefetch_l2(insb + o);
        const char* pfp = insb + 11u * NT * 16u + tid * 128u;      // threads 0 .. NT/8-1: the lines of one trip
        uint32_t pfo = 11u * NT * 16u + tid * 128u + (tid < NT / 8u ? 0u : 0x40000000u);
        if (tid < TEAM) { c.lastK[0][tid] = 0xFFFFFFFFu; c.lastK[1][tid] = 0xFFFFFFFFu; c.lastK[2][tid] = 0xFFFFFFFFu; }
        __syncthreads();
        uint32_t par = 1, prv = 0;                                  // trip t publishes into lastK[par], the previous trip's keys are in lastK[prv]
#pragma unroll 1
        for (uint32_t base = 0; base < n; base += NT) {
            const uint32_t i = base + tid;
            if (pfo < insBytes) prefetch_l2(pfp);
            pfp += NT * 16u; pfo += NT * 16u;
            const uint4 rd = ld_rec(ins + min(i + 3u * NT, nm1));
            const uint4 r = ra;
            const uint32_t ctr = r.x, ref_ctr = r.y, actor = r.z & 0xFFFFu, ref_actor = r.z >> 16, kind = r.w >> 30;
            bool isIns = false, valid = false;
            uint32_t key = 0;
            if (i < n) {
                if (kind > 1u) fail(PT_LOG_BAD_KIND);
                else if (badId(ctr, actor)) fail(PT_LOG_BAD_OPID);
                else { valid = true; key = keyOf(ctr, actor); if (kind == PT_KIND_INSERT) { isIns = true; T[key] = (uint16_t)i; } }
            }
            const uint32_t myK = isIns ? key : 0xFFFFFFFFu;
            uint32_t prevK = __shfl_up_sync(kFull, myK, 1);
            if (lane == 31) c.lastK[par][warp] = myK;
            __syncthreads();                                       // the trip's ids are in T; every warp's last key is published
            if (lane == 0) prevK = warp ? c.lastK[par][warp - 1] : c.lastK[prv][TEAM - 1];    // the left neighbour sits in another warp
            const bool refOk = ref_ctr != 0 && !badId(ref_ctr, ref_actor);
            const uint32_t rkey = keyOf(ref_ctr, ref_actor);
            bool cand = isIns && refOk && rkey == prevK;           // typing-chain link: the reference element is record i-1
            if (cand && rkey >= key) { fail(PT_LOG_CYCLE); cand = false; }
            const uint32_t insW = __ballot_sync(kFull, isIns), candW = __ballot_sync(kFull, cand);
            if (lane == 0 && i < n) *reinterpret_cast<uint2*>(&WI[i >> 5]) = make_uint2(insW, candW);
            if (valid && !cand) {
                if (ref_ctr == 0) { if (!isIns) fail(PT_LOG_ELEM_NOT_FOUND); }
                else {
                    const uint32_t j = refOk ? (uint32_t)T[rkey] : kNone16;
                    if (j == kNone16 || j >= i) fail(PT_LOG_ELEM_NOT_FOUND);
                    else if (isIns && rkey >= key) fail(PT_LOG_CYCLE);
                    else atomicOr(reinterpret_cast<uint32_t*>(&OD[j >> 5]) + (isIns ? 0u : 1u), 1u << (j & 31));
                }
            }
            ra = rb; rb = rc; rc = rd; prv = par; par = par == 2u ? 0u : par + 1u;
        }
    }
    if (st) atomicMax(&c.status, st);
    {   // duplicate insert opIds leave one table entry: occupied entries must equal the number of inserts
        const uint4* tv = reinterpret_cast<const uint4*>(T);
        uint32_t occ = 0;
        for (uint32_t v = tid; v < (KS * 2u + 15u) / 16u; v += NT) {
            const uint4 q = tv[v];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; k++) occ += ((w[k] & 0xFFFFu) != 0xFFFFu) + ((w[k] >> 16) != 0xFFFFu);
        }
        occ = __reduce_add_sync(kFull, occ);
        if (lane == 0 && occ) atomicAdd(&c.occ, occ);
    }
    __syncthreads();
    auto bail = [&]() { if (tid == 0) { pt_log_result r{}; r.status = c.status; *res = r; } __syncthreads(); };
    if (c.status) { bail(); return 0; }

    // ---- C: run heads, visible bits, prefix popcounts ------------------------------------------------------------------------
    uint32_t M, nvis;
    {
        uint32_t carryH = 0, carryV = 0, nins = 0;
        for (uint32_t wb = 0; wb < NWr; wb += NT) {
            const uint32_t w = wb + tid;
            uint32_t insW = 0, candW = 0, otherW = 0, delW = 0, prevBit = 0;
            if (w < NWr) {
                const uint2 ic = *reinterpret_cast<const uint2*>(&WI[w]); const uint2 od = OD[w];
                insW = ic.x; candW = ic.y; otherW = od.x; delW = od.y;
                if (w) prevBit = OD[w - 1].x >> 31;
            }
            const uint32_t head = insW & (~candW | ((otherW << 1) | prevBit));
            const uint32_t vis = insW & ~delW;
            const uint32_t pc = __popc(head) | (__popc(vis) << 16);
            const uint32_t ex = team_scan<TEAM>(pc, c);
            const uint32_t tot = c.total;
            if (w < NWr) WI[w] = make_uint4(insW, head, vis, (carryH + (ex & 0xFFFFu)) | ((carryV + (ex >> 16)) << 16));
            carryH += tot & 0xFFFFu; carryV += tot >> 16;
            nins += __popc(insW);
        }
        nins = __reduce_add_sync(kFull, nins);
        if (lane == 0 && nins) atomicAdd(&c.n_ins, nins);
        M = carryH; nvis = carryV;
    }
    A.used = markC;
    __syncthreads();
    const uint32_t N = c.n_ins;
    if (c.occ != N) { __syncthreads(); if (tid == 0) c.status = PT_LOG_BAD_OPID; __syncthreads(); bail(); return 0; }
    auto runOf = [&](uint32_t i) -> uint32_t { const uint4 q = WI[i >> 5]; return (q.w & 0xFFFFu) + __popc(q.y & (0xFFFFFFFFu >> (31 - (i & 31)))) - 1u; };
    auto visBefore = [&](uint32_t i) -> uint32_t { const uint4 q = WI[i >> 5]; return (q.w >> 16) + __popc(q.z & ((1u << (i & 31)) - 1u)); };

    // ---- D: run tree; E: Euler tour + splitter ranking of the visible weights -------------------------------------------------
    // tour nodes as in warp_kernel.cuh: one 32-bit word per node (successor | visible weight; later owner splitter | weight prefix);
    // END (terminator id) is a multiple of 8 like every splitter node; splitter ids: k < SPEND: node 8k; SPEND: terminator; SPEND + 1: head
    const uint32_t E = 2 * (M + 1), END = (E + 7u) & ~7u;
    if (END + 1 >= 0xFFFFu) return 1;
    const uint32_t SPEND = END >> 3, nSp = SPEND + 2, KW = (KS + 31) / 32;
    uint32_t* Node = A.alloc<uint32_t>(E);
    uint16_t* N16 = reinterpret_cast<uint16_t*>(Node);             // N16[2x] = successor of x, N16[2x + 1] = weight of x
    uint16_t* HV = A.alloc<uint16_t>(M + 1);
    uint16_t* Prun = A.alloc<uint16_t>(M + 1);
    uint16_t* VisBase = Prun;                                      // written at the very end, when the parent links are dead
    uint16_t* RKey = A.alloc<uint16_t>(M + 2);
    uint16_t* Last = RKey;
    // ByG[pos] (run with the pos-th smallest head key) lives in the weight halves of the exit nodes until the tour is threaded
    auto ByG = [&](uint32_t pos) -> uint16_t& { return N16[2 * ((M + 1) + pos) + 1]; };
    const uint32_t uBytes = max(max((uint32_t)(((KW + 1) * 4 + 15) & ~15u) + (uint32_t)(((KW + 1) * 2 + 15) & ~15u), 2u * (uint32_t)((nSp * 4 + 15) & ~15u)),
                                (uint32_t)(((M + 32) * 2 + 15) & ~15u));
    char* U = A.alloc<char>(uBytes);     // successively: key bitmap + prefix | per-warp run lists of the threading | splitter summaries
    if (!A.fits()) return 1;
    uint32_t* KBits = reinterpret_cast<uint32_t*>(U);
    uint16_t* KPre = reinterpret_cast<uint16_t*>(U + (((KW + 1) * 4 + 15) & ~15u));
    uint32_t* Sub = reinterpret_cast<uint32_t*>(U);
    uint32_t* Sub2 = reinterpret_cast<uint32_t*>(U + ((nSp * 4 + 15) & ~15u));
    for (uint32_t w = tid; w < NWr; w += NT) {
        const uint4 q = WI[w];
        uint32_t hb = q.y, rid = q.w & 0xFFFFu;
        while (hb) { const uint32_t b = __ffs(hb) - 1; hb &= hb - 1; HV[rid++] = (uint16_t)(w * 32 + b); }
    }
    for (uint32_t w = tid; w < KW + 1; w += NT) KBits[w] = 0;
    __syncthreads();
    for (uint32_t r = tid; r < M; r += NT) {
        const uint32_t i = HV[r], w = i >> 5, b = i & 31;
        const uint4 q0 = WI[w];
        uint32_t stop = (q0.y | ~q0.x) & ~(0xFFFFFFFFu >> (31 - b));
        uint32_t ww = w;
        while (!stop) { ww++; const uint2 q1 = *reinterpret_cast<const uint2*>(&WI[ww]); stop = q1.y | ~q1.x; }
        const uint32_t end = ww * 32 + (__ffs(stop) - 1);
        const uint4 rec = ld_rec(ins + i);
        const uint32_t key = keyOf(rec.x, rec.z & 0xFFFFu);
        const uint32_t p = rec.y == 0 ? n : (uint32_t)T[keyOf(rec.y, rec.z >> 16)];
        const uint32_t q = p == n ? M : runOf(p);
        const uint32_t hv = visBefore(i);
        N16[2 * r + 1] = (uint16_t)(visBefore(end) - hv);
        Prun[r] = (uint16_t)q; RKey[r] = (uint16_t)key;
        atomicOr(&KBits[key >> 5], 1u << (key & 31));
        HV[r] = (uint16_t)hv;                                       // (each thread rewrites only its own entries)
    }
    __syncthreads();
    {
        uint32_t carry = 0;
        for (uint32_t wb = 0; wb < KW; wb += NT) {
            const uint32_t w = wb + tid;
            const uint32_t cnt = w < KW ? __popc(KBits[w]) : 0u;
            const uint32_t ex = team_scan<TEAM>(cnt, c);
            if (w < KW) KPre[w] = (uint16_t)(carry + ex);
            carry += c.total;
        }
    }
    __syncthreads();
    for (uint32_t r = tid; r < M; r += NT) {
        const uint32_t key = RKey[r];
        ByG((uint32_t)KPre[key >> 5] + __popc(KBits[key >> 5] & ((1u << (key & 31)) - 1u))) = (uint16_t)r;
    }
    __syncthreads();
    for (uint32_t x = tid; x < M + 2; x += NT) Last[x] = (uint16_t)kNone16;      // RKey, KBits, KPre are dead from here
    __syncthreads();
    {
        // thread the runs in ASCENDING key order: the previously threaded child of the same parent is the next sibling in
        // descending-opId order (src/micromerge.ts:628-635), the last one threaded is the first child.  The chain only links runs
        // of ONE parent, so the parents are dealt over the warps (parent mod TEAM): every warp first extracts its runs from the
        // key-ordered list (two streaming passes, stable), then threads its own short list; no two warps touch the same parent.
        uint16_t* Lst = reinterpret_cast<uint16_t*>(U);            // (the key bitmap is dead; the splitter summaries come later)
        uint32_t mine = 0;
        for (uint32_t cb = 0; cb < M; cb += 32) {
            const uint32_t pos = cb + lane;
            const bool ok = pos < M && ((uint32_t)Prun[ByG(pos)] % TEAM) == warp;
            mine += __popc(__ballot_sync(kFull, ok));
        }
        if (lane == 0) c.wa[warp] = mine;
        __syncthreads();
        uint32_t off = 0;
        for (uint32_t w2 = 0; w2 < warp; w2++) off += c.wa[w2];
        __syncthreads();                                           // (the bitmap region may be overwritten from here)
        uint32_t o = off;
        for (uint32_t cb = 0; cb < M; cb += 32) {
            const uint32_t pos = cb + lane;
            const uint32_t r = pos < M ? (uint32_t)ByG(pos) : 0u;
            const bool ok = pos < M && ((uint32_t)Prun[r] % TEAM) == warp;
            const uint32_t bal = __ballot_sync(kFull, ok);
            if (ok) Lst[o + __popc(bal & lt)] = (uint16_t)r;
            o += __popc(bal);
        }
        __syncwarp();
#pragma unroll 1
        for (uint32_t cb = 0; cb < mine; cb += 32) {
            const uint32_t pos = cb + lane;
            const bool valid = pos < mine;
            const uint32_t r = valid ? (uint32_t)Lst[off + pos] : 0u;
            const uint32_t q = valid ? (uint32_t)Prun[r] : (0x10000u + lane);
            // siblings inside this chunk of 32: MATCH.ANY over the parents (measured: cheaper than any probe that would avoid it)
            uint32_t mask = 1u << lane;
            const uint32_t pm = __ballot_sync(kFull, valid);
            if (valid) mask = __match_any_sync(pm, q);
            const uint32_t lower = mask & lt;
            const uint32_t src = lower ? (31u - __clz(lower)) : lane;
            const uint32_t rs = __shfl_sync(kFull, r, src);
            uint32_t ns = kNone16;
            if (valid) ns = lower ? rs : (uint32_t)Last[q];
            __syncwarp();
            if (valid) {
                N16[2 * ((M + 1) + r)] = (uint16_t)(ns != kNone16 ? ns : (M + 1) + q);   // exit(r): next sibling, else exit(parent); (other warps may still read ByG)
                if (((mask >> lane) >> 1) == 0) Last[q] = (uint16_t)r;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (uint32_t r = tid; r <= M; r += NT) { const uint32_t f = Last[r]; N16[2 * r] = (uint16_t)(f != kNone16 ? f : (M + 1) + r); if (r < M) N16[2 * ((M + 1) + r) + 1] = 0; }   // exits weigh 0 (ByG is dead)
    if (tid == 0) { N16[2 * M + 1] = 0; Node[(M + 1) + M] = END; }
    __syncthreads();
    {
        const uint32_t headNode = M;                               // (no node's successor is the tour's first node)
        for (uint32_t k = tid; k < nSp; k += NT) {
            uint32_t cur = k < SPEND ? 8 * k : headNode, acc = 0, nx = END;
            const bool valid = k < SPEND ? cur < E : (k > SPEND && (headNode & 7u) != 0);
            if (valid) {
                for (;;) {
                    const uint32_t a = Node[cur];
                    nx = a & 0xFFFFu;
                    Node[cur] = k | (acc << 16);                   // owner | weight prefix before this node
                    acc += a >> 16;
                    if ((nx & 7u) == 0) break;
                    cur = nx;
                }
            }
            Sub[k] = (acc << 16) | (nx >> 3);                      // invalid / terminator entries: weight 0, successor SPEND
        }
        __syncthreads();
        uint32_t *cur = Sub, *nxt2 = Sub2;
        for (uint32_t span = 1; span < nSp + 1; span <<= 1) {
            for (uint32_t x = tid; x < nSp; x += NT) {
                const uint32_t a = cur[x], b = cur[a & 0xFFFFu];
                nxt2[x] = ((a & 0xFFFF0000u) + (b & 0xFFFF0000u)) | (b & 0xFFFFu);
            }
            __syncthreads();
            uint32_t* t = cur; cur = nxt2; nxt2 = t;
        }
        for (uint32_t r = tid; r < M; r += NT) {
            const uint32_t a = Node[r];
            const uint32_t suf = (cur[a & 0xFFFFu] >> 16) - (a >> 16);
            VisBase[r] = (uint16_t)((nvis - suf) - (uint32_t)HV[r]);
        }
    }
    __syncthreads();

    // ---- F: text out (micromerge.ts:747-750) + digest: one warp per record word ------------------------------------------------
    {
        unsigned long long d0 = 0, d1 = 0;
        // the value tokens are 4-byte reads scattered over the record array: two words per warp in flight
        for (uint32_t w = warp; w + 1 < NWr; w += 2 * TEAM) {
            const uint32_t w2 = w + TEAM;
            const uint4 qa = WI[w], qb = w2 + 1 < NWr ? WI[w2] : make_uint4(0, 0, 0, 0);
            const bool va = (qa.z >> lane) & 1u, vb = (qb.z >> lane) & 1u;
            uint32_t ta = 0, tb = 0;
            if (va) ta = __ldg(&ins[w * 32 + lane].payload);
            if (vb) tb = __ldg(&ins[w2 * 32 + lane].payload);
            if (va) {
                const uint32_t run = (qa.w & 0xFFFFu) + __popc(qa.y & (0xFFFFFFFFu >> (31 - lane))) - 1u;
                const uint32_t vr = ((uint32_t)VisBase[run] + (qa.w >> 16) + __popc(qa.z & lt)) & 0xFFFFu;
                const uint32_t tok = PT_PAYLOAD_TOKEN(ta);
                text_out[vr] = tok;
                digest_add(d0, d1, pt_term_text(vr, tok));
            }
            if (vb) {
                const uint32_t run = (qb.w & 0xFFFFu) + __popc(qb.y & (0xFFFFFFFFu >> (31 - lane))) - 1u;
                const uint32_t vr = ((uint32_t)VisBase[run] + (qb.w >> 16) + __popc(qb.z & lt)) & 0xFFFFu;
                const uint32_t tok = PT_PAYLOAD_TOKEN(tb);
                text_out[vr] = tok;
                digest_add(d0, d1, pt_term_text(vr, tok));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(kFull, d0, o); d1 ^= __shfl_xor_sync(kFull, d1, o); }
        if (lane == 0 && (d0 | d1)) { atomicAdd(&c.dig0, d0); atomicXor(&c.dig1, d1); }
    }
    __syncthreads();
    if (tid == 0) {
        // no marks: one span {} covering all visible text (peritext.ts:392), none if the text is empty
        unsigned long long d0 = c.dig0, d1 = c.dig1;
        const uint32_t nspans = nvis ? 1u : 0u;
        if (nvis) {
            pt_span s; s.start = 0; s.flags = 0; s.link_attr = PT_ATTR_NONE; s.comment_off = 0;
            (P.spans + P.span_off[li])[0] = s;
            digest_add(d0, d1, pt_term_span(0, 0, 0, PT_ATTR_NONE));
        }
        pt_log_result r;
        r.status = PT_LOG_OK; r.n_elems = N; r.n_visible = nvis; r.n_spans = nspans;
        const uint64_t t = pt_term_counts(nvis, nspans);
        r.digest[0] = d0 + t; r.digest[1] = d1 ^ pt_term_hi(t);
        *res = r;
    }
    __syncthreads();
    return 0;
}

template <int TEAM>
__global__ void __launch_bounds__(TEAM * 32, 2048 / (TEAM * 32) > 4 ? 4 : 2048 / (TEAM * 32)) merge_logs_team_kernel(const BatchParams P) {
    __shared__ TeamCtx<TEAM> ctx;
    const uint32_t n_work = P.n_work;
    uint32_t done = 0, deferred = 0;
    if (threadIdx.x == 0) ctx.work_next = atomicAdd(P.work_counter, 1u);
    __syncthreads();
    for (;;) {
        const uint32_t w = ctx.work_next;
        __syncthreads();
        if (w >= n_work) break;
        if (threadIdx.x == 0) ctx.work_next = atomicAdd(P.work_counter, 1u);
        const uint32_t li = P.order[w];
        if (P.admit && P.admit[li]) { __syncthreads(); continue; }
        const int rc = team_merge_one_log<TEAM>(P, li, ctx);
        __syncthreads();
        if (rc) { if (threadIdx.x == 0) P.retry_list[atomicAdd(P.retry_count, 1u)] = li; deferred++; } else done++;
    }
    if (threadIdx.x == 0) {
        if (done) atomicAdd(&P.stats[0], (unsigned long long)done);
        if (deferred) atomicAdd(&P.stats[2], (unsigned long long)deferred);
    }
}

}  // namespace ptk
