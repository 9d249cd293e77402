// warp_kernel.cuh — op-log apply + flatten for SHORT logs: ONE WARP materialises one log (sm_100a).
//
// Same closed form as merge_kernel.cuh (reference src/micromerge.ts:534-724, src/peritext.ts:154-455),
// re-cut for logs of up to a few thousand records (BASELINE.json configs[3]: 100K docs x 1K ops x 3 replicas):
//   * no block barriers: every scan is a shuffle scan, every reduction a redux.sync, phases are separated by __syncwarp;
//     several warps of a CTA work on different logs, each in its own slice of dynamic shared memory
//   * 16-bit everything (record indices, opId keys K = (ctr-1)*R + actor, Euler nodes = next:16 | visible weight:16)
//   * sibling order without sorting groups: run heads are ranked by K with a key-space bitmap (unique keys: counting sort),
//     then threaded in ASCENDING K with __match_any_sync — the previously threaded sibling of the same parent is the next
//     sibling in the reference's descending order (src/micromerge.ts:628-635), the last one threaded is the first child
//   * marks are resolved in VISIBLE space: a mark op covers visible element v iff vis(start slot) <= v < vis(end slot); ops
//     that cover no visible element (most of them in fuzz-shaped logs, where nearly everything is a tombstone) are dropped
//     right after their two boundary lookups; spans come from the few survivors by stabbing the elementary segments
//   * comment ops fold in ARRIVAL order (src/peritext.ts:314-322 has no opId comparison; quirk Q4)
// A log that does not fit the warp's slice, or whose surviving mark set is too large for the stabbing loops, is DEFERRED on
// the device to the block kernel's bins (merge_kernel.cuh) — results are identical either way.
#pragma once
#include "merge_kernel.cuh"

namespace ptk {

constexpr uint32_t kFull = 0xffffffffu;
constexpr uint32_t kNone16 = 0xFFFFu;
constexpr uint32_t kWarpGrab = 4;            // logs taken from the work queue per atomic
constexpr uint32_t kMaxSegSurvivorWork = 1536;   // ceil(S/32) * nS above this: defer to the block kernel's segment trees
constexpr uint32_t kMaxCommentSurvivors = 48;    // the comment loops are quadratic in the surviving comment ops

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(kFull, v, o); if (lane >= (uint32_t)o) v += y; }
    return v;
}

// per-warp bump allocator over the warp's slice of dynamic shared memory (offsets from the __shared__ symbol: LDS/STS)
struct WArena {
    uint32_t base, used, cap;
    // bump; the caller checks fits() once per group of allocations, BEFORE touching any of them
    template <class T> __device__ __forceinline__ T* alloc(uint32_t count) {
        const uint32_t off = used;
        used += (uint32_t)((count * sizeof(T) + 15u) & ~15u);
        return reinterpret_cast<T*>(ptk_smem + base + off);
    }
    __device__ __forceinline__ bool fits() const { return used <= cap; }
};

// Phase alignment of the warps of one CTA (optional).  The kernel's code is far larger than the instruction caches
// (L0 ~6 KB per scheduler, L1.5 32 KB per SM), and warps that drift apart each stream their own part of it from L2; named
// barriers at a few phase boundaries keep the warps of a CTA in the same code region.  A warp that leaves a log early
// (error status, deferral, no work) ARRIVES at the remaining barriers without waiting, so nobody waits for it.
constexpr uint32_t kFirstPhaseBar = 2, kLastPhaseBar = 5;      // barrier 1 = start of a round (work loop)
__device__ __noinline__ void phase_arrive_rest(uint32_t next, uint32_t nthreads, uint32_t skip) {
    for (; next <= kLastPhaseBar; next++)
        if (!((skip >> (next - kFirstPhaseBar)) & 1u)) asm volatile("barrier.arrive %0, %1;" ::"r"(next), "r"(nthreads) : "memory");
}
struct PhaseSync {
    uint32_t on, nthreads, next, skip;      // skip: bit k set = phase barrier kFirstPhaseBar + k is not used (tuning)
    __device__ __forceinline__ void pass() {
        if (on && !((skip >> (next - kFirstPhaseBar)) & 1u)) asm volatile("barrier.sync %0, %1;" ::"r"(next), "r"(nthreads) : "memory");
        next++;
    }
    __device__ __forceinline__ void leave() {
        if (on && next <= kLastPhaseBar) phase_arrive_rest(next, nthreads, skip);
        next = kLastPhaseBar + 1;
    }
};

template <class T>
__device__ __forceinline__ void wfill(T* p, uint32_t count, T v, uint32_t lane) {   // allocations are padded to 16 B
    const uint32_t nvec = (uint32_t)((count * sizeof(T) + 15u) >> 4);
    uint32_t w;
    if (sizeof(T) == 1) w = 0x01010101u * (uint32_t)(uint8_t)v;
    else if (sizeof(T) == 2) w = 0x00010001u * (uint32_t)(uint16_t)v;
    else w = (uint32_t)v;
    const uint4 q = make_uint4(w, w, w, w);
    uint4* d = reinterpret_cast<uint4*>(p);
#pragma unroll 1
    for (uint32_t i = lane; i < nvec; i += 32) d[i] = q;
}

// id-table forms of the warp kernel (one kernel instantiation each; the host sorts the logs into the launches)
constexpr int kIdDirect = 0, kIdCompact = 1, kIdPacked3 = 2;

// returns 0: done (result header written), 1: defer to the block kernel.  IDM selects the id-table form (below).
template <int IDM>
__device__ __forceinline__ int warp_merge_one_log(const BatchParams& P, const uint32_t li, const uint32_t slice_base, const uint32_t slice_bytes, const uint32_t li_next, PhaseSync ps) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt = (1u << lane) - 1u;

    // descriptor: all lanes read the same 32 bytes (one broadcast transaction)
    const uint4 dsc0 = __ldg(reinterpret_cast<const uint4*>(P.desc + li)), dsc1 = __ldg(reinterpret_cast<const uint4*>(P.desc + li) + 1);
    const unsigned long long insdel_off = (unsigned long long)dsc0.x | ((unsigned long long)dsc0.y << 32);
    const unsigned long long mark_off = (unsigned long long)dsc0.z | ((unsigned long long)dsc0.w << 32);
    const uint32_t n = dsc1.x, m = dsc1.y, R = dsc1.z ? dsc1.z : 1u, C = dsc1.w;
    const unsigned long long KS64 = (unsigned long long)C * R;
    if (KS64 >= 0xFFFFull || n >= 0xFFFFu || m >= 0xFFFFu) { ps.leave(); return 1; }     // 16-bit keys / indices only
    const uint32_t KS = (uint32_t)KS64;
    const pt_insdel_rec* __restrict__ ins = P.insdel + insdel_off;
    const pt_mark_rec* __restrict__ mk = P.marks + mark_off;
    uint32_t* text_out = P.text + P.text_off[li];
    pt_span* span_out = P.spans + P.span_off[li];
    pt_log_result* res = P.results + li;

    if ((P.warp_flags & 1u) && m) {       // this log's mark records are needed late: pull them into L2 now
        const char* p0 = reinterpret_cast<const char*>(mk);
        const uint32_t lines = (m * (uint32_t)sizeof(pt_mark_rec) + 127u) >> 7;
        for (uint32_t l = lane; l < lines; l += 32) prefetch_l2(p0 + ((size_t)l << 7));
    }

    WArena A; A.base = slice_base; A.used = 0; A.cap = slice_bytes;
    uint32_t st = 0;                                           // lane-local failures, bit (1 << code); the checkpoints report the highest code
    auto keyOf = [&](uint32_t ctr, uint32_t actor) -> uint32_t { return (ctr - 1u) * R + actor; };
    auto badId = [&](uint32_t ctr, uint32_t actor) -> bool { return ctr - 1u >= C || actor >= R; };
    auto fail = [&](uint32_t code) { st |= 1u << code; };
    auto bail = [&](uint32_t code) { if (lane == 0) { pt_log_result r{}; r.status = code; *res = r; } };

    // ---- id table: opId -> insert record index ------------------------------------------------------------------------------
    // direct : T[K(ctr, actor)], 2 bytes per key of the key space C*R (what merge_kernel.cuh does).
    // compact: with >= 3 actors most of that space is empty (c4: 250 inserts in 2200 keys), and shared memory per warp is what
    //          bounds the number of resident warps.  T[ctr-1] = actor:5 | index:11 of ONE insert with that counter; the few
    //          inserts that share a counter with an earlier one (concurrent edits) go to a small open-addressing overflow
    //          table OV (key:16 | index:16, linear probing).  More than kOvMax of those: the log is deferred.
    // packed3: exactly 3 actors and <= 1022 records (c4's shape: three concurrent replicas): one 32-bit word per COUNTER holds
    //          the three actors' record indices + 1 in 10-bit fields.  An insert is ONE atomicOr (the old value tells a
    //          duplicate opId), a lookup one LDS + shift + mask; no overflow table, nothing to probe.
    constexpr bool compact = IDM == kIdCompact, packed = IDM == kIdPacked3;
    if (compact && !(R >= 3u && R <= 30u && n <= 2046u)) { ps.leave(); return 1; }     // (the host only sends such logs to this launch)
    if (packed && !(R == 3u && n <= 1022u)) { ps.leave(); return 1; }
    constexpr uint32_t kOvSlots = 128, kOvMax = 96, kOvEmpty = 0xFFFFFFFFu;
    const uint32_t NWr = (n + 31) / 32 + 1;                    // bit words over record indices (+1 zero pad word)
    // layout: the arrays at FIXED offsets first (their addresses are one add away from the slice base)
    uint32_t* OV = A.alloc<uint32_t>(compact ? kOvSlots : 0u);
    // per 32-record word: x = insert bits, y = chain-link bits (after C: run-head bits), z = visible bits (after C),
    // w = run heads before the word | visible elements before the word << 16 (after C)
    uint4* WI = A.alloc<uint4>(NWr);
    uint16_t* T = A.alloc<uint16_t>(packed ? 2u * C : compact ? C : KS);
    uint32_t* T32 = reinterpret_cast<uint32_t*>(T);              // packed3 view
    if (!A.fits()) { ps.leave(); return 1; }
    if (packed) wfill<uint32_t>(T32, C, 0u, lane);
    else wfill<uint16_t>(T, compact ? C : KS, (uint16_t)kNone16, lane);
    if (compact) wfill<uint32_t>(OV, kOvSlots, kOvEmpty, lane);
    // during A+B the z / w fields of WI collect the "element has a child that is not its log successor" and tombstone bits (atomicOr)
    wfill<uint32_t>(reinterpret_cast<uint32_t*>(WI), 4 * NWr, 0u, lane);
    __syncwarp();
    auto ovHash = [&](uint32_t key) -> uint32_t { return ((key * 40503u) >> 7) & (kOvSlots - 1u); };
    // index of the insert record with opId (ctr, actor), kNone16 if there is none; the id must be in range (!badId)
    auto lookup = [&](uint32_t ctr, uint32_t actor) -> uint32_t {
        if (packed) { const uint32_t f = (T32[ctr - 1u] >> (10u * actor)) & 1023u; return f ? f - 1u : kNone16; }
        if (!compact) return T[keyOf(ctr, actor)];
        const uint32_t e = T[ctr - 1u];
        if (e == kNone16) return kNone16;                      // no insert with this counter at all
        if ((e >> 11) == actor) return e & 0x7FFu;
        const uint32_t key = keyOf(ctr, actor);
        for (uint32_t h = ovHash(key);; h = (h + 1u) & (kOvSlots - 1u)) {
            const uint32_t v = OV[h];
            if (v == kOvEmpty) return kNone16;
            if ((v >> 16) == key) return v & 0xFFFFu;
        }
    };

    // ---- A+B: one pass over the ins/del records, 32 per trip, two trips in flight ------------------------------------------
    // A: id table, insert bits, chain-link bits (reference element == the insert at record i-1: compare with the left
    //    neighbour's key, no lookup).  B: parents of non-chain inserts ("has another child" bits) and deletes (tombstones, OR).
    //    A referenced element must have arrived EARLIER in the log (src/micromerge.ts:752 throws otherwise).
    uint32_t nOv = 0;
    {
        uint32_t carryK = 0xFFFFFFFFu;                         // key of the last record of the previous trip if it is an insert
        // records past the end are loaded from a clamped index and ignored (every use is guarded by i < n)
        const uint32_t nm1 = n ? n - 1u : 0u;
        uint4 ra = make_uint4(0, 0, 0, 0);
        if (n) ra = ld_rec(ins + min(lane, nm1));
        // an L2 prefetch stream runs >= 10 trips (512 B each) ahead of the register loads: DRAM latency under load is longer than
        // two trips.  Every 8th trip all 32 lanes fetch the 32 lines of 8 later trips (one uniform branch per trip otherwise).
        const char* insb = reinterpret_cast<const char*>(ins);
        const uint32_t insBytes = n * 16u;
        if (lane * 128u + 1024u < insBytes) prefetch_l2(insb + 1024u + lane * 128u);          // trips 2 .. 9
#pragma unroll 2
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t i = base + lane;
            if ((base & 255u) == 0u) {                                                       // trips t+10 .. t+17
                const uint32_t po = (base + 320u) * 16u + lane * 128u;
                if (po < insBytes) prefetch_l2(insb + po);
            }
            const uint4 rc = ld_rec(ins + min(i + 32u, nm1));           // one trip ahead (the lines are in L2 by now); unrolled by 2: no moves
            const uint4 r = ra;
            const uint32_t ctr = r.x, ref_ctr = r.y, actor = r.z & 0xFFFFu, ref_actor = r.z >> 16, kind = r.w >> 30;
            // straight-line form: predicates instead of nested branches
            const bool inb = i < n;
            const bool kbad = kind > 1u, ibad = badId(ctr, actor);
            const bool valid = inb && !kbad && !ibad;
            if (inb && kbad) fail(PT_LOG_BAD_KIND);
            if (inb && ibad) fail(PT_LOG_BAD_OPID);
            const uint32_t key = keyOf(ctr, actor);
            const bool isIns = valid && kind == PT_KIND_INSERT;
            bool toOv = false, wrote = false;
            uint32_t mine = 0;
            if (isIns) {
                if (packed) {
                    const uint32_t sh = 10u * actor;
                    if ((atomicOr(&T32[ctr - 1u], (i + 1u) << sh) >> sh) & 1023u) fail(PT_LOG_BAD_OPID);   // two inserts with one opId
                } else if (!compact) {
                    if (T[key] != kNone16) fail(PT_LOG_BAD_OPID);      // two inserts with one opId (earlier trip)
                    T[key] = (uint16_t)i;
                } else {
                    mine = (actor << 11) | i;
                    const uint32_t e = T[ctr - 1u];
                    if (e == kNone16) { T[ctr - 1u] = (uint16_t)mine; wrote = true; }
                    else if ((e >> 11) == actor) fail(PT_LOG_BAD_OPID);
                    else toOv = true;
                }
            }
            const uint32_t myK = isIns ? key : 0xFFFFFFFFu;
            uint32_t prevK = __shfl_up_sync(kFull, myK, 1);
            if (lane == 0) prevK = carryK;
            carryK = __shfl_sync(kFull, myK, 31);
            const bool refOk = ref_ctr != 0 && !badId(ref_ctr, ref_actor);
            const uint32_t rkey = keyOf(ref_ctr, ref_actor);
            bool cand = isIns && refOk && rkey == prevK;           // typing-chain link: the reference element is record i-1
            if (cand && rkey >= key) { fail(PT_LOG_CYCLE); cand = false; }
            const uint32_t insW = __ballot_sync(kFull, isIns), candW = __ballot_sync(kFull, cand);
            if (lane == 0) *reinterpret_cast<uint2*>(&WI[base >> 5]) = make_uint2(insW, candW);
            __syncwarp();                                          // the trip's ids are in T
            if (IDM == kIdDirect) {
                if (isIns && T[key] != (uint16_t)i) fail(PT_LOG_BAD_OPID);   // two inserts with one opId (same trip)
            } else if (compact) {
                if (wrote) {                                       // same counter twice in one trip: one lane owns the slot
                    const uint32_t e2 = T[ctr - 1u];
                    if (e2 != mine) { if ((e2 >> 11) == actor) fail(PT_LOG_BAD_OPID); else toOv = true; }
                }
                const uint32_t ovW = __ballot_sync(kFull, toOv);
                if (ovW) {
                    nOv += __popc(ovW);
                    if (toOv && nOv <= kOvMax) {
                        const uint32_t val = (key << 16) | i;
                        for (uint32_t h = ovHash(key);; h = (h + 1u) & (kOvSlots - 1u)) {
                            const uint32_t old = atomicCAS(&OV[h], kOvEmpty, val);
                            if (old == kOvEmpty) break;
                            if ((old >> 16) == key) { fail(PT_LOG_BAD_OPID); break; }
                        }
                    }
                    __syncwarp();
                }
            }
            {   // B: the reference element of a non-chain record (ref_ctr == 0: an insert at the head of the list)
                const bool need = valid && !cand, hasRef = ref_ctr != 0;
                uint32_t j = kNone16;
                if (need && refOk) j = lookup(ref_ctr, ref_actor);
                const bool found = j != kNone16 && j < i;              // must have arrived earlier
                const bool cyc = isIns && rkey >= key;
                if (need) {
                    if (hasRef ? !found : !isIns) fail(PT_LOG_ELEM_NOT_FOUND);
                    else if (hasRef && cyc) fail(PT_LOG_CYCLE);
                }
                if (need && found && !cyc) atomicOr(reinterpret_cast<uint32_t*>(&WI[j >> 5]) + (isIns ? 2u : 3u), 1u << (j & 31));   // deletes: OR, idempotent (micromerge.ts:689)
            }
            ra = rc;
        }
    }
    __syncwarp();
    ps.pass();                                                     // (2) end of the record pass
    if (compact && nOv > kOvMax) { ps.leave(); return 1; }                                    // too many concurrent-counter inserts for the compact table
    st = __reduce_or_sync(kFull, st);
    if (st) { bail(31u - __clz(st)); ps.leave(); return 0; }

    // ---- C: runs, bit-parallel: head = insert & (!chain-link | predecessor has another child); visible = insert & !deleted
    uint32_t M, nvis, N = 0;
    {
        uint32_t carryH = 0, carryV = 0, otherCarry = 0;
        for (uint32_t wb = 0; wb < NWr; wb += 32) {
            const uint32_t w = wb + lane;
            uint32_t insW = 0, candW = 0, otherW = 0, delW = 0;
            if (w < NWr) { const uint4 q = WI[w]; insW = q.x; candW = q.y; otherW = q.z; delW = q.w; }
            const uint32_t up = __shfl_up_sync(kFull, otherW, 1);
            const uint32_t prevBit = lane ? (up >> 31) : otherCarry;
            otherCarry = __shfl_sync(kFull, otherW, 31) >> 31;
            const uint32_t head = insW & (~candW | ((otherW << 1) | prevBit));
            const uint32_t vis = insW & ~delW;
            const uint32_t pc = __popc(head) | (__popc(vis) << 16);
            const uint32_t inc = warp_incl_scan(pc, lane), ex = inc - pc, tot = __shfl_sync(kFull, inc, 31);
            if (w < NWr) WI[w] = make_uint4(insW, head, vis, (carryH + (ex & 0xFFFFu)) | ((carryV + (ex >> 16)) << 16));
            carryH += tot & 0xFFFFu; carryV += tot >> 16;
            N += __reduce_add_sync(kFull, (uint32_t)__popc(insW));
        }
        M = carryH; nvis = carryV;
    }
    __syncwarp();

    auto runOf = [&](uint32_t i) -> uint32_t {
        const uint4 q = WI[i >> 5];
        return (q.w & 0xFFFFu) + __popc(q.y & (0xFFFFFFFFu >> (31 - (i & 31)))) - 1u;
    };
    auto visBefore = [&](uint32_t i) -> uint32_t {
        const uint4 q = WI[i >> 5];
        return (q.w >> 16) + __popc(q.z & ((1u << (i & 31)) - 1u));
    };

    if ((P.warp_flags & 2u) && li_next != 0xFFFFFFFFu && lane == 0) prefetch_l2(P.desc + li_next);   // read in phase F
    if (m) {       // the first 6 trips of mark records (needed in phase G) start their way to L2 now
        const uint32_t pfb = min(m * 32u, 6u * 1024u);
        for (uint32_t o = lane * 128u; o < pfb; o += 32u * 128u) prefetch_l2(reinterpret_cast<const char*>(mk) + o);
    }
    // ---- D: run tree; E: Euler tour + splitter list ranking of the VISIBLE weights ------------------------------------------
    const uint32_t E = 2 * (M + 1), END = (E + 7u) & ~7u;          // END: terminator id, a multiple of 8 like every splitter node
    if (END + 1 >= 0xFFFFu) { ps.leave(); return 1; }
    uint16_t* VisBase = A.alloc<uint16_t>(M + 1);                  // vis(i) = VisBase[run(i)] + visBefore(i)   (mod 2^16)
    const uint32_t markD = A.used;
    {
        // Euler tour nodes: enter(r) = r, exit(r) = (M+1) + r, r in 0..M (M = HEAD).  One 32-bit word per node: low half =
        // successor (later: owner splitter), high half = visible weight of the node (later: weight prefix inside the owner's
        // sublist); exits weigh 0.  Splitter ids: k < SPEND: node 8k; SPEND: the terminator; SPEND + 1: the tour's first node.
        const uint32_t SPEND = END >> 3, nSp = SPEND + 2, KW = (KS + 31) / 32;
        uint32_t* Node = A.alloc<uint32_t>(E);
        uint16_t* N16 = reinterpret_cast<uint16_t*>(Node);         // N16[2x] = successor of x, N16[2x + 1] = weight of x
        uint16_t* HV = A.alloc<uint16_t>(M + 1);                   // run head record index, then visBefore(head)
        uint16_t* Prun = A.alloc<uint16_t>(M + 1);
        uint16_t* RKey = A.alloc<uint16_t>(M + 2);                 // key of the run head; dead after the ranking, then:
        uint16_t* Last = RKey;                                     // last threaded child of run q (q = M: HEAD)
        // ByG[pos] (run with the pos-th smallest head key) lives in the weight halves of the exit nodes until the tour is threaded
        auto ByG = [&](uint32_t pos) -> uint16_t& { return N16[2 * ((M + 1) + pos) + 1]; };
        // key bitmap + prefix (ranking of the head keys); dead after the ranking, then the splitter summaries live there
        const uint32_t uBytes = max((uint32_t)(((KW + 1) * 4 + 15) & ~15u) + (uint32_t)(((KW + 1) * 2 + 15) & ~15u), 2u * (uint32_t)((nSp * 4 + 15) & ~15u));
        char* U = A.alloc<char>(uBytes);
        if (!A.fits()) { ps.leave(); return 1; }
        uint32_t* KBits = reinterpret_cast<uint32_t*>(U);
        uint16_t* KPre = reinterpret_cast<uint16_t*>(U + (((KW + 1) * 4 + 15) & ~15u));
        uint32_t* Sub = reinterpret_cast<uint32_t*>(U);
        uint32_t* Sub2 = reinterpret_cast<uint32_t*>(U + ((nSp * 4 + 15) & ~15u));
#pragma unroll 1
        for (uint32_t wb = 0; wb < NWr; wb += 32) {                // compact the run heads (one bit word per lane)
            const uint32_t w = wb + lane;
            if (w < NWr) {
                const uint4 q = WI[w];
                uint32_t hb = q.y, rid = q.w & 0xFFFFu;
                while (hb) { const uint32_t b = __ffs(hb) - 1; hb &= hb - 1; HV[rid++] = (uint16_t)(w * 32 + b); }
            }
        }
        wfill<uint32_t>(KBits, KW + 1, 0u, lane);
        __syncwarp();
#pragma unroll 1
        for (uint32_t rb = 0; rb < M; rb += 32) {                  // one lane per run: extent, parent run, visible weight, key bit
            const uint32_t r = rb + lane;
            if (r < M) {
                const uint32_t i = HV[r], w = i >> 5, b = i & 31;
                const uint4 q0 = WI[w];
                uint32_t stop = (q0.y | ~q0.x) & ~(0xFFFFFFFFu >> (31 - b));
                uint32_t ww = w;
                while (!stop) { ww++; const uint2 q1 = *reinterpret_cast<const uint2*>(&WI[ww]); stop = q1.y | ~q1.x; }   // pad word: insert bits == 0 -> stops
                const uint32_t end = ww * 32 + (__ffs(stop) - 1);
                const uint4 rec = ld_rec(ins + i);
                const uint32_t key = keyOf(rec.x, rec.z & 0xFFFFu);
                const uint32_t p = rec.y == 0 ? n : lookup(rec.y, rec.z >> 16);
                const uint32_t q = p == n ? M : runOf(p);
                const uint32_t hv = visBefore(i);
                N16[2 * r + 1] = (uint16_t)(visBefore(end) - hv);
                HV[r] = (uint16_t)hv;
                Prun[r] = (uint16_t)q; RKey[r] = (uint16_t)key;
                atomicOr(&KBits[key >> 5], 1u << (key & 31));
            }
        }
        __syncwarp();
        {
            uint32_t carry = 0;
#pragma unroll 1
            for (uint32_t wb = 0; wb < KW; wb += 32) {
                const uint32_t w = wb + lane;
                const uint32_t cnt = w < KW ? __popc(KBits[w]) : 0u;
                const uint32_t inc = warp_incl_scan(cnt, lane);
                if (w < KW) KPre[w] = (uint16_t)(carry + inc - cnt);
                carry += __shfl_sync(kFull, inc, 31);
            }
        }
        __syncwarp();
#pragma unroll 1
        for (uint32_t rb = 0; rb < M; rb += 32) {                  // rank of the run head's key among all run heads (unique keys)
            const uint32_t r = rb + lane;
            if (r < M) {
                const uint32_t key = RKey[r];
                ByG((uint32_t)KPre[key >> 5] + __popc(KBits[key >> 5] & ((1u << (key & 31)) - 1u))) = (uint16_t)r;
            }
        }
        __syncwarp();
        ps.pass();                                                 // (3) run heads ranked
        wfill<uint16_t>(Last, M + 2, (uint16_t)kNone16, lane);     // RKey, KBits, KPre are dead from here
        __syncwarp();
        // thread the runs in ASCENDING key order: among the children of one parent, the previously threaded one is the
        // NEXT sibling in descending-opId order (src/micromerge.ts:628-635), the last one threaded is the FIRST child
#pragma unroll 1
        for (uint32_t cb = 0; cb < M; cb += 32) {
            const uint32_t pos = cb + lane;
            const bool valid = pos < M;
            const uint32_t r = valid ? (uint32_t)ByG(pos) : 0u;
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
                N16[2 * ((M + 1) + r)] = (uint16_t)(ns != kNone16 ? ns : (M + 1) + q);   // exit(r): next sibling, else exit(parent)
                if (((mask >> lane) >> 1) == 0) Last[q] = (uint16_t)r;              // highest lane of its group
            }
            __syncwarp();
        }
#pragma unroll 1
        for (uint32_t rb = 0; rb <= M; rb += 32) {                 // enter(r): first child, else exit(r)
            const uint32_t r = rb + lane;
            if (r <= M) { const uint32_t f = Last[r]; N16[2 * r] = (uint16_t)(f != kNone16 ? f : (M + 1) + r); if (r < M) N16[2 * ((M + 1) + r) + 1] = 0; }   // exits weigh 0 (ByG is dead)
        }
        if (lane == 0) { N16[2 * M + 1] = 0; Node[(M + 1) + M] = END; }
        __syncwarp();
        // splitter list ranking: every 8th node id (and the tour's first node) walks its sublist once; only the splitter
        // summaries are ranked by pointer jumping; suffix(x) = suffix(owner sublist) - prefix(x)
        // (no node's successor is the tour's first node, and END is a multiple of 8: a sublist ends where the successor id is)
        const uint32_t headNode = M;
#pragma unroll 1
        for (uint32_t kb = 0; kb < nSp; kb += 32) {
            const uint32_t k = kb + lane;
            if (k < nSp) {
                uint32_t cur = k < SPEND ? 8 * k : headNode, acc = 0, nx = END;
                const bool valid = k < SPEND ? cur < E : (k > SPEND && (headNode & 7u) != 0);
                if (valid) {
                    for (;;) {
                        const uint32_t a = Node[cur];
                        nx = a & 0xFFFFu;
                        Node[cur] = k | (acc << 16);               // owner | weight prefix before this node
                        acc += a >> 16;
                        if ((nx & 7u) == 0) break;
                        cur = nx;
                    }
                }
                Sub[k] = (acc << 16) | (nx >> 3);                  // invalid / terminator entries: weight 0, successor SPEND
            }
        }
        __syncwarp();
        {
            uint32_t *cur = Sub, *nxt2 = Sub2;
            for (uint32_t span = 1; span < nSp + 1; span <<= 1) {
#pragma unroll 1
                for (uint32_t x = lane; x < nSp; x += 32) {
                    const uint32_t a = cur[x], b = cur[a & 0xFFFFu];
                    nxt2[x] = ((a & 0xFFFF0000u) + (b & 0xFFFF0000u)) | (b & 0xFFFFu);
                }
                __syncwarp();
                uint32_t* t = cur; cur = nxt2; nxt2 = t;
            }
            Sub = cur;
        }
#pragma unroll 1
        for (uint32_t rb = 0; rb < M; rb += 32) {
            const uint32_t r = rb + lane;
            if (r < M) {
                const uint32_t a = Node[r];
                const uint32_t suf = (Sub[a & 0xFFFFu] >> 16) - (a >> 16);          // visible elements from run r to the end
                VisBase[r] = (uint16_t)((nvis - suf) - (uint32_t)HV[r]);
            }
        }
        __syncwarp();
    }
    A.used = markD;                                                // release the run-tree temporaries
    ps.pass();                                                     // (4) sequence ranked

    // ---- F: per-element visible rank table + text out (visible index = prefix count of non-deleted elements,
    // micromerge.ts:747-750).  EV[i] = visible elements before insert record i in the sequence | visible << 15: every mark
    // boundary below is then one table read instead of run / prefix arithmetic.
    uint16_t* EV = A.alloc<uint16_t>(n + 1);
    if (!A.fits() || nvis >= 0x8000u) { ps.leave(); return 1; }
    unsigned long long d0 = 0, d1 = 0;
#pragma unroll 1
    for (uint32_t w = 0; w + 1 < NWr; w++) {
        const uint4 q = WI[w];                                     // uniform: one broadcast LDS.128
        const uint32_t ib = q.x;
        if (!ib) continue;
        const uint32_t hbits = q.y, vbits = q.z, hp = q.w & 0xFFFFu, vp = q.w >> 16;
        if ((ib >> lane) & 1u) {
            const uint32_t i = w * 32 + lane;
            const uint32_t run = hp + __popc(hbits & (0xFFFFFFFFu >> (31 - lane))) - 1u;
            const uint32_t vr = ((uint32_t)VisBase[run] + vp + __popc(vbits & lt)) & 0xFFFFu;
            const uint32_t isv = (vbits >> lane) & 1u;
            EV[i] = (uint16_t)(vr | (isv << 15));
            if (isv) {
                const uint32_t tok = PT_PAYLOAD_TOKEN(__ldg(&ins[i].payload));
                text_out[vr] = tok;
                digest_add(d0, d1, pt_term_text(vr, tok));
            }
        }
    }
    __syncwarp();

    uint32_t nspans = 0;
    if ((P.warp_flags & 2u) && li_next != 0xFFFFFFFFu) {        // the next log's ins/del records -> L2 while this one does its marks
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(P.desc + li_next)), q1 = __ldg(reinterpret_cast<const uint4*>(P.desc + li_next) + 1);
        const char* p0 = reinterpret_cast<const char*>(P.insdel + ((unsigned long long)q0.x | ((unsigned long long)q0.y << 32)));
        const uint32_t lines = (q1.x * (uint32_t)sizeof(pt_insdel_rec) + 127u) >> 7;
        for (uint32_t l = lane; l < lines; l += 32) prefetch_l2(p0 + ((size_t)l << 7));
    }

    uint32_t nS = 0, nC = 0;
    uint4* Sv = nullptr;                                           // survivors: {va | vb << 16, priority:16 | kind << 16, attr, -}
    uint16_t* CIdx = nullptr;                                      // surviving comment ops (indices into Sv), arrival order
    if (m) {
        // ---- G: mark ops -> visible intervals [va, vb); only ops that cover a visible element survive ---------------------
        const uint32_t KW = (KS + 31) / 32;
        uint32_t* KBits = A.alloc<uint32_t>(KW + 1);               // duplicate mark opIds
        CIdx = A.alloc<uint16_t>(kMaxCommentSurvivors + 32);
        if (!A.fits()) { ps.leave(); return 1; }
        const uint32_t room = A.cap - A.used, svStart = A.used;
        uint32_t capS = room > 256u ? (room - 256u) / 16u : 0u;
        if (capS > m) capS = m;
        Sv = A.alloc<uint4>(capS + 1);
        if (!A.fits()) { ps.leave(); return 1; }
        wfill<uint32_t>(KBits, KW + 1, 0u, lane);
        __syncwarp();
        const uint4* mq = reinterpret_cast<const uint4*>(mk);
        const uint32_t mm1 = m - 1u;                               // m > 0 here; past-the-end lanes load a clamped record, unused
        uint4 a0 = __ldg(mq + 2 * min(lane, mm1)), a1 = __ldg(mq + 2 * min(lane, mm1) + 1);
        const uint32_t mkBytes = m * 32u;
#pragma unroll 2
        for (uint32_t kb = 0; kb < m; kb += 32) {
            const uint32_t k = kb + lane;
            if ((kb & 127u) == 0u) {                               // every 4th trip: the 32 lines of trips t+6 .. t+9 (1 KB per trip)
                const uint32_t po = (kb + 192u) * 32u + lane * 128u;
                if (po < mkBytes) prefetch_l2(reinterpret_cast<const char*>(mk) + po);
            }
            const uint32_t kn = min(k + 32u, mm1);
            const uint4 b0 = __ldg(mq + 2 * kn), b1 = __ldg(mq + 2 * kn + 1);   // one trip ahead (L2 hits); unrolled by 2: no moves
            // a0 = {ctr, actor|kind<<16|bounds<<24, start_ctr, end_ctr}; a1 = {start_actor|end_actor<<16, attr, arrival, reserved}
            const uint32_t ctr = a0.x, actor = a0.y & 0xFFFFu, kind = (a0.y >> 16) & 0xFFu, bounds = a0.y >> 24;
            const uint32_t start_ctr = a0.z, end_ctr = a0.w, start_actor = a1.x & 0xFFFFu, end_actor = a1.x >> 16, attr = a1.y, arrival = a1.z;
            const uint32_t type = (kind >> 1) & 3u;
            // straight-line form (predicated loads instead of nested branches).  A boundary element must exist AND have arrived
            // before the mark op: the reference's walk never matches anything else (peritext.ts:236-241) — a missing start is a
            // no-op, a missing end never ends
            const bool inb = k < m;
            const bool idok = inb && !badId(ctr, actor);
            const uint32_t key = keyOf(ctr, actor);
            bool dup = false;
            if (idok) {
                const uint32_t bit = 1u << (key & 31);
                dup = (atomicOr(&KBits[key >> 5], bit) & bit) != 0 || lookup(ctr, actor) != kNone16;      // duplicate opId
            }
            if (inb && (!idok || dup)) fail(PT_LOG_BAD_OPID);
            const uint32_t sb = bounds & 3u, eb = (bounds >> 2) & 3u;
            const bool sOk = idok && sb <= PT_BOUND_AFTER && !badId(start_ctr, start_actor);
            uint32_t js = kNone16;
            if (sOk) js = lookup(start_ctr, start_actor);
            const bool sHit = js != kNone16 && js < arrival;
            uint32_t es = 0;
            if (sHit) es = EV[js];
            const uint32_t va = (es & 0x7FFFu) + (sb & (es >> 15));
            const bool eOk = sHit && eb <= PT_BOUND_AFTER && !badId(end_ctr, end_actor);
            uint32_t je = kNone16;
            if (eOk) je = lookup(end_ctr, end_actor);
            // same slot: the start branch wins and the op never ends (quirk Q2)
            const bool eHit = je != kNone16 && je < arrival && !(je == js && eb == sb);
            uint32_t vb = nvis;
            if (eHit) { const uint32_t ee = EV[je]; vb = (ee & 0x7FFFu) + (eb & (ee >> 15)); }
            const bool surv = sHit && va < vb;
            const bool isC = surv && type == PT_MARK_COMMENT;
            const uint32_t bal = __ballot_sync(kFull, surv), balC = __ballot_sync(kFull, isC);
            if (surv) {
                const uint32_t idx = nS + __popc(bal & lt);
                if (idx < capS) {
                    // priority: LWW types compare opIds (peritext.ts:304-313) = keys; comments fold in arrival order (Q4)
                    Sv[idx] = make_uint4(va | (vb << 16), (type == PT_MARK_COMMENT ? k : key) | (kind << 16), attr, 0u);
                    if (isC) { const uint32_t ci = nC + __popc(balC & lt); if (ci <= kMaxCommentSurvivors) CIdx[ci] = (uint16_t)idx; }
                }
            }
            nS += __popc(bal); nC += __popc(balC);
            a0 = b0; a1 = b1;
        }
        __syncwarp();
        st = __reduce_or_sync(kFull, st);
        if (st) { bail(31u - __clz(st)); ps.leave(); return 0; }
        if (nS > capS) { ps.leave(); return 1; }
        A.used = svStart + ((nS * 16u + 15u) & ~15u);               // keep only the survivors
    }

    ps.pass();                                                     // (5) marks resolved
    unsigned long long pool_base = 0;
    if (nvis == 0) nspans = 0;
    else if (nS == 0) {
        // no mark op touches a visible element: one span {} (peritext.ts:392)
        if (lane == 0) {
            pt_span s; s.start = 0; s.flags = 0; s.link_attr = PT_ATTR_NONE; s.comment_off = 0;
            span_out[0] = s;
            digest_add(d0, d1, pt_term_span(0, 0, 0, PT_ATTR_NONE));
        }
        nspans = 1;
    } else {
        // ---- I: elementary segments of the visible text -> marks per segment -> spans ---------------------------------------
        if (nC > kMaxCommentSurvivors) { ps.leave(); return 1; }
        if (nvis <= 32u && nC <= 32u) {
            // ---- I (short text): at most 32 visible characters: one lane per visible POSITION instead of elementary segments — no
            // boundary bitmap, no segment arrays.  Marks / link / comment-id set per position; a span starts where any of them
            // differs from the position before (same result as the segment form below: nothing changes inside a segment).
            const uint32_t x = lane;
            uint32_t w0 = 0, w1 = 0, w2 = 0, flags = 0, link = PT_ATTR_NONE;
#pragma unroll 1
            for (uint32_t j = 0; j < nS; j++) {
                const uint4 sv = Sv[j];                              // one broadcast LDS.128
                const uint32_t ab = sv.x, pk = sv.y;
                const bool cover = x >= (ab & 0xFFFFu) && x < (ab >> 16);
                const uint32_t t = (pk >> 17) & 3u, val = (((pk & 0xFFFFu) << 16) | j) + 1u;
                if (cover) {
                    if (t == PT_MARK_STRONG) w0 = max(w0, val);
                    else if (t == PT_MARK_EM) w1 = max(w1, val);
                    else if (t == PT_MARK_LINK) w2 = max(w2, val);
                    else flags |= PT_SPAN_COMMENT;                   // `comment` key present iff any comment op covers (quirk Q3)
                }
            }
            // LWW winners (peritext.ts:304-313): the max-opId covering op of the type; present iff it is an addMark
            if (w0 && !((Sv[(w0 - 1u) & 0xFFFFu].y >> 16) & 1u)) flags |= PT_SPAN_STRONG;
            if (w1 && !((Sv[(w1 - 1u) & 0xFFFFu].y >> 16) & 1u)) flags |= PT_SPAN_EM;
            if (w2) { const uint32_t j2 = (w2 - 1u) & 0xFFFFu; if (!((Sv[j2].y >> 16) & 1u)) { flags |= PT_SPAN_LINK; link = Sv[j2].z; } }
            // comment ids present at x: ids whose last-arrived covering op is an add (peritext.ts:314-322); the set is a bit mask
            // over the surviving comment ops, an id being named by the FIRST surviving op that carries it
            uint32_t cmask = 0;
            if (nC) {
                uint32_t canon = lane;
                if (lane < nC) {
                    const uint32_t id = Sv[CIdx[lane]].z;
                    for (uint32_t c0 = 0; c0 < lane; c0++) if (Sv[CIdx[c0]].z == id) { canon = c0; break; }
                }
#pragma unroll 1
                for (uint32_t cj = 0; cj < nC; cj++) {
                    const uint4 s1 = Sv[CIdx[cj]];                   // uniform
                    const uint32_t rep = __shfl_sync(kFull, canon, cj);
                    const bool cov = x >= (s1.x & 0xFFFFu) && x < (s1.x >> 16);
                    if (!__any_sync(kFull, cov)) continue;
                    bool later = false;
                    for (uint32_t c2 = cj + 1; c2 < nC; c2++) {
                        const uint4 s2 = Sv[CIdx[c2]];
                        if (s2.z != s1.z) continue;                  // uniform
                        if (x >= (s2.x & 0xFFFFu) && x < (s2.x >> 16)) later = true;
                    }
                    if (cov && !later && !((s1.y >> 16) & 1u)) cmask |= 1u << rep;
                }
            }
            const uint32_t pf = __shfl_up_sync(kFull, flags, 1), pl = __shfl_up_sync(kFull, link, 1), pm = __shfl_up_sync(kFull, cmask, 1);
            const bool head = x < nvis && (x == 0 || ((flags ^ pf) & 0xFu) != 0 || link != pl || cmask != pm);
            const uint32_t hb = __ballot_sync(kFull, head);
            const uint32_t cnt = head ? (uint32_t)__popc(cmask) : 0u;
            const uint32_t inc = warp_incl_scan(cnt, lane), off = inc - cnt, totalC = __shfl_sync(kFull, inc, 31);
            nspans = __popc(hb);
            if (totalC) {
                uint32_t pst = 0;
                if (lane == 0) pool_base = pool_reserve(P, totalC, pst);
                pst = __shfl_sync(kFull, pst, 0);
                if (pst) { bail(pst); ps.leave(); return 0; }
                pool_base = __shfl_sync(kFull, pool_base, 0);
            }
            if (head) {
                const uint32_t jo = __popc(hb & lt);
                uint32_t* pool = P.comment_pool + pool_base;
                uint32_t filled = 0;
                for (uint32_t mm = cmask; mm; mm &= mm - 1u) {       // ascending id order (sortBy, peritext.ts:318); the lists are short
                    const uint32_t id = Sv[CIdx[__ffs(mm) - 1]].z;
                    uint32_t y = filled;
                    while (y > 0 && pool[off + y - 1] > id) { pool[off + y] = pool[off + y - 1]; y--; }
                    pool[off + y] = id; filled++;
                }
                pt_span sp; sp.start = x; sp.flags = (flags & 0xFu) | (cnt << 8); sp.link_attr = link;
                sp.comment_off = cnt ? (uint32_t)(pool_base + off) : 0u;
                span_out[jo] = sp;
                for (uint32_t y = 0; y < cnt; y++) digest_add(d0, d1, pt_term_comment(jo, y, pool[off + y]));
                digest_add(d0, d1, pt_term_span(jo, sp.start, sp.flags, sp.link_attr));
            }
        } else {
        const uint32_t BW = nvis / 32 + 1;
        uint32_t* Bnd = A.alloc<uint32_t>(BW + 1);
        uint16_t* BPre = A.alloc<uint16_t>(BW + 1);
        if (!A.fits()) { ps.leave(); return 1; }
        wfill<uint32_t>(Bnd, BW + 1, 0u, lane);
        __syncwarp();
#pragma unroll 1
        for (uint32_t s = lane; s < nS; s += 32) {
            const uint32_t ab = Sv[s].x, va = ab & 0xFFFFu, vb = ab >> 16;
            atomicOr(&Bnd[va >> 5], 1u << (va & 31));
            if (vb < nvis) atomicOr(&Bnd[vb >> 5], 1u << (vb & 31));
        }
        __syncwarp();
        uint32_t nB = 0;
#pragma unroll 1
        for (uint32_t wb = 0; wb < BW; wb += 32) {
            const uint32_t w = wb + lane;
            const uint32_t cnt = w < BW ? __popc(Bnd[w]) : 0u;
            const uint32_t inc = warp_incl_scan(cnt, lane);
            if (w < BW) BPre[w] = (uint16_t)(nB + inc - cnt);
            nB += __shfl_sync(kFull, inc, 31);
        }
        const uint32_t S = nB + 1;                                   // segment s >= 1 starts at the s-th boundary; segment 0 = [0, first)
        if (((S + 31) / 32) * nS > kMaxSegSurvivorWork) { ps.leave(); return 1; }
        uint16_t* SegStart = A.alloc<uint16_t>(S + 1);
        uint32_t* SegFlags = A.alloc<uint32_t>(S + 1);               // bits3:0 span flags, bit4 comment set differs from x-1, bit5 head
        uint32_t* SegLink = A.alloc<uint32_t>(S + 1);
        uint16_t* SegCnt = A.alloc<uint16_t>(S + 1);                 // comment ids of the span starting here
        uint16_t* SegOut = A.alloc<uint16_t>(S + 1);                 // span index
        uint32_t* SegCOff = A.alloc<uint32_t>(S + 1);                // offset of its comment list in the log's pool reservation
        if (!A.fits()) { ps.leave(); return 1; }
        __syncwarp();
#pragma unroll 1
        for (uint32_t wb = 0; wb < BW; wb += 32) {
            const uint32_t w = wb + lane;
            if (w < BW) {
                uint32_t bb = Bnd[w], id = (uint32_t)BPre[w] + 1u;
                while (bb) { const uint32_t b = __ffs(bb) - 1; bb &= bb - 1; SegStart[id++] = (uint16_t)(w * 32 + b); }
            }
        }
        if (lane == 0) SegStart[0] = 0;
        const bool seg0_empty = (Bnd[0] & 1u) != 0;                  // position 0 is itself a boundary
        __syncwarp();
        // pass 1: marks of every segment = stabbing query over the survivors (uniform loop, broadcast reads)
#pragma unroll 1
        for (uint32_t sb = 0; sb < S; sb += 32) {
            const uint32_t s = sb + lane;
            const uint32_t x = s < S ? (uint32_t)SegStart[s] : 0u;
            uint32_t w0 = 0, w1 = 0, w2 = 0, flags = 0, link = PT_ATTR_NONE;
#pragma unroll 1
            for (uint32_t j = 0; j < nS; j++) {
                const uint4 sv = Sv[j];                              // one broadcast LDS.128
                const uint32_t ab = sv.x, pk = sv.y;
                const bool cover = x >= (ab & 0xFFFFu) && x < (ab >> 16);
                const uint32_t t = (pk >> 17) & 3u, val = (((pk & 0xFFFFu) << 16) | j) + 1u;
                if (cover) {
                    if (t == PT_MARK_STRONG) w0 = max(w0, val);
                    else if (t == PT_MARK_EM) w1 = max(w1, val);
                    else if (t == PT_MARK_LINK) w2 = max(w2, val);
                    else flags |= PT_SPAN_COMMENT;                   // `comment` key present iff any comment op covers (quirk Q3)
                }
            }
            // LWW winners (peritext.ts:304-313): the max-opId covering op of the type; present iff it is an addMark
            if (w0 && !((Sv[(w0 - 1u) & 0xFFFFu].y >> 16) & 1u)) flags |= PT_SPAN_STRONG;
            if (w1 && !((Sv[(w1 - 1u) & 0xFFFFu].y >> 16) & 1u)) flags |= PT_SPAN_EM;
            if (w2) { const uint32_t j2 = (w2 - 1u) & 0xFFFFu; if (!((Sv[j2].y >> 16) & 1u)) { flags |= PT_SPAN_LINK; link = Sv[j2].z; } }
            // comment ids differ between x-1 and x?  only ids with a boundary exactly at x can change; presence of an id =
            // "its last-arrived covering op is an add" (peritext.ts:314-322)
            if (nC) {
                const bool live = s >= 1 && s < S && x > 0;
                bool cd = false;
                for (uint32_t cj = 0; cj < nC; cj++) {
                    const uint32_t j = CIdx[cj], ab = Sv[j].x;
                    const bool touch = live && ((ab & 0xFFFFu) == x || (ab >> 16) == x);
                    if (!__any_sync(kFull, touch)) continue;
                    const uint32_t id = Sv[j].z;
                    bool pPrev = false, pCur = false;
                    for (uint32_t c2 = 0; c2 < nC; c2++) {
                        const uint32_t j2 = CIdx[c2];
                        const uint4 s2 = Sv[j2];
                        if (s2.z != id) continue;                    // uniform
                        const uint32_t a2 = s2.x & 0xFFFFu, b2 = s2.x >> 16;
                        const bool add2 = !((s2.y >> 16) & 1u);
                        if (x - 1u >= a2 && x - 1u < b2) pPrev = add2;
                        if (x >= a2 && x < b2) pCur = add2;
                    }
                    if (touch && pPrev != pCur) cd = true;
                }
                if (cd) flags |= 16u;
            }
            if (s < S) { SegFlags[s] = flags; SegLink[s] = link; }
        }
        __syncwarp();
        // pass 2: span heads, span indices, comment counts
        uint32_t totalC = 0;
#pragma unroll 1
        for (uint32_t sb = 0; sb < S; sb += 32) {
            const uint32_t s = sb + lane;
            bool head = false;
            uint32_t cnt = 0;
            const uint32_t x = s < S ? (uint32_t)SegStart[s] : 0u;
            if (s < S) {
                const uint32_t f = SegFlags[s];
                if (s == 0) head = !seg0_empty;
                else if (x == 0) head = true;
                else head = ((f ^ SegFlags[s - 1]) & 0xFu) != 0 || SegLink[s] != SegLink[s - 1] || (f & 16u);
            }
            if (nC) {
                for (uint32_t cj = 0; cj < nC; cj++) {               // members: ids whose last-arrived covering op is an add
                    const uint32_t j = CIdx[cj], ab = Sv[j].x, id = Sv[j].z;
                    const bool cov = head && x >= (ab & 0xFFFFu) && x < (ab >> 16);
                    if (!__any_sync(kFull, cov)) continue;
                    bool later = false;
                    for (uint32_t c2 = cj + 1; c2 < nC; c2++) {
                        const uint32_t j2 = CIdx[c2];
                        if (Sv[j2].z != id) continue;
                        const uint32_t ab2 = Sv[j2].x;
                        if (x >= (ab2 & 0xFFFFu) && x < (ab2 >> 16)) later = true;
                    }
                    if (cov && !later && !((Sv[j].y >> 16) & 1u)) cnt++;
                }
            }
            const uint32_t hb = __ballot_sync(kFull, head);
            const uint32_t inc = warp_incl_scan(cnt, lane);
            __syncwarp();                                            // every lane has read its left neighbour's flags
            if (s < S) {
                SegFlags[s] = (SegFlags[s] & 0x1Fu) | (head ? 32u : 0u);
                SegCnt[s] = (uint16_t)cnt;
                SegOut[s] = (uint16_t)(nspans + __popc(hb & lt));
                SegCOff[s] = totalC + inc - cnt;
            }
            nspans += __popc(hb);
            totalC += __shfl_sync(kFull, inc, 31);
            __syncwarp();
        }
        if (totalC) {
            uint32_t pst = 0;
            if (lane == 0) pool_base = pool_reserve(P, totalC, pst);
            pst = __shfl_sync(kFull, pst, 0);
            if (pst) { bail(pst); ps.leave(); return 0; }
            pool_base = __shfl_sync(kFull, pool_base, 0);
        }
        __syncwarp();
        // pass 3: span records, comment lists (ascending id, sortBy peritext.ts:318), digest
        uint32_t* pool = P.comment_pool + pool_base;
#pragma unroll 1
        for (uint32_t sb = 0; sb < S; sb += 32) {
            const uint32_t s = sb + lane;
            const bool head = s < S && (SegFlags[s] & 32u);
            const uint32_t x = s < S ? (uint32_t)SegStart[s] : 0u;
            const uint32_t cnt = head ? (uint32_t)SegCnt[s] : 0u, off = head ? SegCOff[s] : 0u;
            if (nC && __any_sync(kFull, cnt != 0)) {
                uint32_t filled = 0;
                for (uint32_t cj = 0; cj < nC; cj++) {
                    const uint32_t j = CIdx[cj], ab = Sv[j].x, id = Sv[j].z;
                    const bool cov = cnt != 0 && x >= (ab & 0xFFFFu) && x < (ab >> 16);
                    if (!__any_sync(kFull, cov)) continue;
                    bool later = false;
                    for (uint32_t c2 = cj + 1; c2 < nC; c2++) {
                        const uint32_t j2 = CIdx[c2];
                        if (Sv[j2].z != id) continue;
                        const uint32_t ab2 = Sv[j2].x;
                        if (x >= (ab2 & 0xFFFFu) && x < (ab2 >> 16)) later = true;
                    }
                    if (cov && !later && !((Sv[j].y >> 16) & 1u)) {  // insert in ascending id order (lists are short)
                        uint32_t y = filled;
                        while (y > 0 && pool[off + y - 1] > id) { pool[off + y] = pool[off + y - 1]; y--; }
                        pool[off + y] = id; filled++;
                    }
                }
            }
            if (head) {
                const uint32_t jo = SegOut[s];
                pt_span sp; sp.start = x; sp.flags = (SegFlags[s] & 0xFu) | (cnt << 8); sp.link_attr = SegLink[s];
                sp.comment_off = cnt ? (uint32_t)(pool_base + off) : 0u;
                span_out[jo] = sp;
                for (uint32_t y = 0; y < cnt; y++) digest_add(d0, d1, pt_term_comment(jo, y, pool[off + y]));
                digest_add(d0, d1, pt_term_span(jo, sp.start, sp.flags, sp.link_attr));
            }
        }
        }   // segment form
    }

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(kFull, d0, o); d1 ^= __shfl_xor_sync(kFull, d1, o); }
    if (lane == 0) {
        pt_log_result r;
        r.status = PT_LOG_OK; r.n_elems = N; r.n_visible = nvis; r.n_spans = nspans;
        const uint64_t t = pt_term_counts(nvis, nspans);
        r.digest[0] = d0 + t; r.digest[1] = d1 ^ pt_term_hi(t);
        *res = r;
    }
    ps.leave();
    return 0;
}

// Persistent warps: every warp pulls logs (largest first) from the bin's work queue.  Two modes:
//   free  : each warp takes kWarpGrab logs per atomic and runs on its own;
//   phased: (warp_flags bit 2, the default) the CTA takes one log per warp per ROUND: its warps start a round together (consecutive
//           logs of the size-sorted queue are nearly the same size, so they also finish together) and share the instruction caches;
//           the named barriers at the phase boundaries INSIDE a log are optional (bits 8-11 skip them; default: only the last one).
template <int WARPS, int IDM>
__global__ void __launch_bounds__(WARPS * 32, (32 / WARPS) > 0 ? (32 / WARPS) : 1) merge_logs_warp_kernel(const BatchParams P) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t n_work = P.n_work;
    const uint32_t slice = P.smem_arena_bytes;
    uint32_t base = warp * slice;
    asm volatile("" : "+r"(base));         // opaque: keep it in a register instead of re-deriving it from threadIdx at every shared-memory access
    uint32_t done = 0, deferred = 0;
    const bool phased = (P.warp_flags & 4u) != 0;
    __shared__ uint32_t s_base[2], s_nxt[2];
    PhaseSync ps; ps.on = phased ? 1u : 0u; ps.nthreads = WARPS * 32; ps.next = kFirstPhaseBar; ps.skip = (P.warp_flags >> 8) & 0xFu;
    uint32_t nextb = 0, nextb2 = 0, par = 0;   // phased: the CTA's next two rounds (held by thread 0)
    uint32_t w = 0, wend = 0, wn = 0;      // free: this warp's current grab [w, wend) and the next one
    if (phased) { if (threadIdx.x == 0) { nextb = atomicAdd(P.work_counter, (uint32_t)WARPS); nextb2 = atomicAdd(P.work_counter, (uint32_t)WARPS); } }
    else {
        if (lane == 0) { w = atomicAdd(P.work_counter, kWarpGrab); wn = atomicAdd(P.work_counter, kWarpGrab); }
        w = __shfl_sync(kFull, w, 0); wn = __shfl_sync(kFull, wn, 0);
        wend = min(w + kWarpGrab, n_work);
    }
    for (;;) {
        uint32_t x, xn = 0xFFFFFFFFu;
        if (phased) {
            // the round after this one is known too (fetched a round ago): its logs' records can be on their way to L2
            if (threadIdx.x == 0) { s_base[par] = nextb; s_nxt[par] = nextb2; nextb = nextb2; nextb2 = atomicAdd(P.work_counter, (uint32_t)WARPS); }
            asm volatile("barrier.sync 1, %0;" ::"r"((uint32_t)(WARPS * 32)) : "memory");
            const uint32_t b0 = s_base[par], bn = s_nxt[par];
            par ^= 1u;
            if (b0 >= n_work) break;
            x = b0 + warp;
            xn = bn + warp;
            ps.next = kFirstPhaseBar;
        } else {
            if (w >= wend) {
                w = wn;
                if (w >= n_work) break;
                if (lane == 0) wn = atomicAdd(P.work_counter, kWarpGrab);
                wn = __shfl_sync(kFull, wn, 0);
                wend = min(w + kWarpGrab, n_work);
            }
            x = w++;
            xn = w < wend ? w : wn;
        }
        if (x < n_work) {
            const uint32_t li = P.order[x];
            if (P.admit && P.admit[li]) { ps.leave(); continue; }      // rejected by the admission pre-pass
            const uint32_t li_next = ((P.warp_flags & 2u) && xn < n_work) ? P.order[xn] : 0xFFFFFFFFu;
            const int rc = warp_merge_one_log<IDM>(P, li, base, slice, li_next, ps);   // the host put the log in the right launch
            __syncwarp();
            if (rc) { if (lane == 0) P.retry_list[atomicAdd(P.retry_count, 1u)] = li; deferred++; } else done++;
        } else ps.leave();
    }
    if (lane == 0) {
        if (done) atomicAdd(&P.stats[0], (unsigned long long)done);
        if (deferred) atomicAdd(&P.stats[2], (unsigned long long)deferred);
    }
}

}  // namespace ptk
