// merge_kernel.cuh — the op-log apply + flatten kernel (sm_100a).
//
// One CTA materialises one LOG (one replica's op log of one document) end to end:
//   packed records in HBM  ->  element sequence (RGA order)  ->  visible text + formatted spans + digest in HBM.
// It computes the ORDER-INDEPENDENT CLOSED FORM of what the reference does sequentially in
//   Micromerge.applyOp / applyListInsert / applyListUpdate     (reference src/micromerge.ts:534-724)
//   applyAddRemoveMark                                          (reference src/peritext.ts:154-249)
//   getTextWithFormatting / opsToMarks / addCharactersToSpans   (reference src/peritext.ts:294-455)
// (proven equal to the sequential oracle by tests/test_closed_form.py on the CPU model
//  tests/kernel_model.py, whose phase names A..I this file follows).
//
// No floating point, no tensor cores: integer/index work bounded by HBM traffic and shared-memory latency.
// Working arrays live in a per-CTA ARENA: dynamic shared memory first, a per-CTA global slab (L2 resident) as
// spill for logs that do not fit.  Index arrays are u16 when the log is small enough (halves the footprint).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/peritext_b200.h"
#include "../../include/pt_digest.h"

namespace ptk {

struct BatchParams {
    const pt_log_desc* __restrict__ desc;
    const pt_insdel_rec* __restrict__ insdel;
    const pt_mark_rec* __restrict__ marks;
    const uint32_t* __restrict__ order;   // log indices of this launch (bin), largest first
    uint32_t n_work;
    uint32_t* work_counter;               // persistent-CTA work queue head
    pt_log_result* results;
    const uint64_t* __restrict__ text_off;
    const uint64_t* __restrict__ span_off;
    uint32_t* text;
    pt_span* spans;
    uint32_t* comment_pool;
    unsigned long long* comment_used;
    unsigned long long comment_cap;
    char* slab;                           // spill: slab_bytes per CTA
    unsigned long long slab_bytes;
    uint32_t smem_arena_bytes;            // dynamic shared memory given to the arena
};

// ---------------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_rec(const pt_insdel_rec* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t lanemask_le() { uint32_t m; asm("mov.u32 %0, %%lanemask_le;" : "=r"(m)); return m; }

template <int BLOCK>
struct BlockCtx {
    // static shared scratch shared by all phases
    uint32_t warp_sums[32];
    uint32_t scan_total;
    uint32_t status;
    uint32_t work;
    uint32_t n_ins;        // number of insert records (elements)
    uint32_t M;            // number of runs
    uint32_t nvis;
    uint32_t misc[8];
    unsigned long long dig0, dig1;
    unsigned long long pool_base;
};

// exclusive block scan of one value per thread; returns exclusive prefix, total via ctx (valid after return)
template <int BLOCK>
__device__ __forceinline__ uint32_t block_scan_excl(uint32_t v, BlockCtx<BLOCK>& c, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= (uint32_t)o) x += y; }
    if (lane == 31) c.warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (BLOCK / 32) ? c.warp_sums[lane] : 0;
        uint32_t s = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (uint32_t)o) s += y; }
        if (lane < (BLOCK / 32)) c.warp_sums[lane] = s - w;
        if (lane == 31) c.scan_total = s;
    }
    __syncthreads();
    uint32_t res = c.warp_sums[warp] + x - v;
    total = c.scan_total;
    __syncthreads();   // warp_sums reusable
    return res;
}

struct Arena {
    char* sm; uint32_t sm_cap, sm_used;
    char* gm; unsigned long long gm_cap, gm_used;
    bool overflow;
    template <class T> __device__ __forceinline__ T* alloc(uint32_t count) {
        uint32_t bytes = (uint32_t)((count * sizeof(T) + 15u) & ~15u);
        if (sm_used + bytes <= sm_cap) { T* p = reinterpret_cast<T*>(sm + sm_used); sm_used += bytes; return p; }
        if (gm_used + bytes > gm_cap) { overflow = true; return reinterpret_cast<T*>(gm); }
        T* p = reinterpret_cast<T*>(gm + gm_used); gm_used += bytes; return p;
    }
};

template <class T, int BLOCK>
__device__ __forceinline__ void fill(T* p, uint32_t n, T v) {
    for (uint32_t i = threadIdx.x; i < n; i += BLOCK) p[i] = v;
}

__device__ __forceinline__ void digest_add(unsigned long long& d0, unsigned long long& d1, uint64_t t) {
    d0 += t; d1 += pt_term_hi(t);
}

template <int BLOCK>
__device__ __forceinline__ void digest_flush(BlockCtx<BLOCK>& c, unsigned long long d0, unsigned long long d1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&c.dig0, d0); atomicAdd(&c.dig1, d1); }
}

// =========================================================================================================
// The per-log pipeline.  Idx = uint16_t (logs with < 32000 records) or uint32_t.
// =========================================================================================================
template <class Idx, int BLOCK>
__device__ void merge_one_log(const BatchParams& P, uint32_t li, BlockCtx<BLOCK>& c, char* smem_arena) {
    constexpr Idx NONE = (Idx)~(Idx)0;
    const uint32_t tid = threadIdx.x, lane = tid & 31;

    const pt_log_desc L = P.desc[li];
    const uint32_t n = L.n_insdel, m = L.n_mark, R = L.n_actors ? L.n_actors : 1, C = L.max_ctr;
    const uint32_t KS = C * R;
    const pt_insdel_rec* __restrict__ ins = P.insdel + L.insdel_off;
    const pt_mark_rec* __restrict__ mk = P.marks + L.mark_off;
    uint32_t* text_out = P.text + P.text_off[li];
    pt_span* span_out = P.spans + P.span_off[li];
    pt_log_result* res = P.results + li;

    Arena A;
    A.sm = smem_arena; A.sm_cap = P.smem_arena_bytes; A.sm_used = 0;
    A.gm = P.slab + (unsigned long long)blockIdx.x * P.slab_bytes; A.gm_cap = P.slab_bytes; A.gm_used = 0; A.overflow = false;

    if (tid == 0) { c.status = 0; c.n_ins = 0; c.M = 0; c.nvis = 0; c.dig0 = 0; c.dig1 = 0; c.pool_base = 0; }

    auto keyOf = [&](uint32_t ctr, uint32_t actor) -> uint32_t { return (ctr - 1u) * R + actor; };
    auto badId = [&](uint32_t ctr, uint32_t actor) -> bool { return ctr - 1u >= C || actor >= R; };
    auto fail = [&](uint32_t code) { atomicMax(&c.status, code); };

    // ---- arrays that live through most phases (allocation order = smem priority) ---------------------------
    Idx* T = A.alloc<Idx>(KS);            // A: opId key -> insert record index
    Idx* Par = A.alloc<Idx>(n);           // B: parent record index | n (HEAD) | NONE (not an insert)
    Idx* RunOrPos = A.alloc<Idx>(n);      // C: run id, overwritten by sequence position in F
    Idx* AnyChild = A.alloc<Idx>(n + 1);  // B: some child of each element (slot n = HEAD)
    uint8_t* Multi = A.alloc<uint8_t>(n + 1);   // B2: element has >= 2 children
    uint8_t* Del = A.alloc<uint8_t>(n);         // B: tombstone flag per record index

    fill<Idx, BLOCK>(T, KS, NONE);
    fill<Idx, BLOCK>(AnyChild, n + 1, NONE);
    fill<uint8_t, BLOCK>(Multi, n + 1, (uint8_t)0);
    fill<uint8_t, BLOCK>(Del, n, (uint8_t)0);
    __syncthreads();

    // ---- A: id table ---------------------------------------------------------------------------------------
    {
        uint32_t cnt = 0;
        for (uint32_t i = tid; i < n; i += BLOCK) {
            uint4 r = ld_rec(ins + i);
            uint32_t ctr = r.x, actor = r.z & 0xFFFFu, kind = r.w >> 30;
            if (kind > 1u) { fail(PT_LOG_BAD_KIND); continue; }
            if (badId(ctr, actor)) { fail(PT_LOG_BAD_OPID); continue; }
            if (kind == PT_KIND_INSERT) { T[keyOf(ctr, actor)] = (Idx)i; cnt++; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0 && cnt) atomicAdd(&c.n_ins, cnt);
    }
    __syncthreads();
    if (c.status) { if (tid == 0) { pt_log_result r{}; r.status = c.status; *res = r; } __syncthreads(); return; }
    const uint32_t N = c.n_ins;

    // ---- B: parents, deletes (+ duplicate-id detection) ---------------------------------------------------------
    for (uint32_t i = tid; i < n; i += BLOCK) {
        uint4 r = ld_rec(ins + i);
        uint32_t ctr = r.x, ref_ctr = r.y, actor = r.z & 0xFFFFu, ref_actor = r.z >> 16, kind = r.w >> 30;
        if (kind == PT_KIND_INSERT) {
            uint32_t k = keyOf(ctr, actor);
            if (T[k] != (Idx)i) fail(PT_LOG_BAD_OPID);          // two inserts with one opId
            uint32_t p;
            if (ref_ctr == 0) p = n;
            else {
                Idx j = badId(ref_ctr, ref_actor) ? NONE : T[keyOf(ref_ctr, ref_actor)];
                if (j == NONE) { fail(PT_LOG_ELEM_NOT_FOUND); Par[i] = NONE; continue; }
                if (keyOf(ref_ctr, ref_actor) >= k) { fail(PT_LOG_CYCLE); Par[i] = NONE; continue; }
                p = j;
            }
            Par[i] = (Idx)p;
            AnyChild[p] = (Idx)i;                                 // arbitrary winner among the children
        } else {
            Par[i] = NONE;
            Idx j = (ref_ctr == 0 || badId(ref_ctr, ref_actor)) ? NONE : T[keyOf(ref_ctr, ref_actor)];
            if (j == NONE) { fail(PT_LOG_ELEM_NOT_FOUND); continue; }
            Del[j] = 1;                                           // OR over deletes: idempotent (micromerge.ts:689)
        }
    }
    __syncthreads();
    if (c.status) { if (tid == 0) { pt_log_result r{}; r.status = c.status; *res = r; } __syncthreads(); return; }
    // B2: flag elements with more than one child (children that lost the AnyChild race reveal it)
    for (uint32_t i = tid; i < n; i += BLOCK) {
        Idx p = Par[i];
        if (p != NONE && AnyChild[p] != (Idx)i) Multi[p] = 1;
    }
    __syncthreads();

    // ---- C: runs = log-contiguous only-child chains ------------------------------------------------------------------
    Idx* RunHead = A.alloc<Idx>(N + 1);
    Idx* RunTail = A.alloc<Idx>(N + 1);
    {
        uint32_t carry = 0;
        for (uint32_t base = 0; base < n; base += BLOCK) {
            uint32_t i = base + tid;
            bool isIns = false, head = false, tail = false;
            if (i < n) {
                Idx p = Par[i];
                isIns = p != NONE;
                if (isIns) {
                    bool cont = i > 0 && (uint32_t)p == i - 1 && !Multi[i - 1];
                    head = !cont;
                    bool nextCont = (i + 1 < n) && (uint32_t)Par[i + 1] == i && !Multi[i];
                    tail = !nextCont;
                }
            }
            uint32_t total;
            uint32_t ex = block_scan_excl<BLOCK>(head ? 1u : 0u, c, total);
            uint32_t rid = carry + ex + (head ? 1u : 0u) - 1u;     // run id of element i (inclusive count - 1)
            if (isIns) {
                RunOrPos[i] = (Idx)rid;
                if (head) RunHead[rid] = (Idx)i;
                if (tail) RunTail[rid] = (Idx)i;
            }
            carry += total;
        }
        if (tid == 0) c.M = carry;
    }
    __syncthreads();
    const uint32_t M = c.M;

    // ---- D: run tree; children of every node ordered by DESCENDING opId of the run head ---------------------------------
    Idx* Prun = A.alloc<Idx>(M + 1);
    uint32_t* Key = A.alloc<uint32_t>(M + 1);
    uint32_t* GrpCnt = A.alloc<uint32_t>(M + 2);    // children per node (node M = HEAD); reused as cursor
    Idx* GrpOff = A.alloc<Idx>(M + 2);
    Idx* Unsorted = A.alloc<Idx>(M + 1);
    Idx* Sorted = A.alloc<Idx>(M + 1);
    Idx* SPos = A.alloc<Idx>(M + 1);
    fill<uint32_t, BLOCK>(GrpCnt, M + 2, 0u);
    __syncthreads();
    for (uint32_t r = tid; r < M; r += BLOCK) {
        uint32_t h = RunHead[r];
        uint32_t p = Par[h];
        uint32_t q = (p == n) ? M : (uint32_t)RunOrPos[p];
        Prun[r] = (Idx)q;
        uint4 rec = ld_rec(ins + h);
        Key[r] = keyOf(rec.x, rec.z & 0xFFFFu);
        atomicAdd(&GrpCnt[q], 1u);
    }
    __syncthreads();
    {
        uint32_t carry = 0;
        for (uint32_t base = 0; base < M + 1; base += BLOCK) {
            uint32_t q = base + tid;
            uint32_t v = q < M + 1 ? GrpCnt[q] : 0u, total;
            uint32_t ex = block_scan_excl<BLOCK>(v, c, total);
            if (q < M + 1) GrpOff[q] = (Idx)(carry + ex);
            carry += total;
        }
    }
    __syncthreads();
    uint32_t* GrpCur = A.alloc<uint32_t>(M + 2);
    fill<uint32_t, BLOCK>(GrpCur, M + 2, 0u);
    __syncthreads();
    for (uint32_t r = tid; r < M; r += BLOCK) {
        uint32_t q = Prun[r];
        uint32_t slot = (uint32_t)GrpOff[q] + atomicAdd(&GrpCur[q], 1u);
        Unsorted[slot] = (Idx)r;
    }
    __syncthreads();
    for (uint32_t r = tid; r < M; r += BLOCK) {
        uint32_t q = Prun[r], cnt = GrpCnt[q], off = GrpOff[q];
        uint32_t rank = 0;
        if (cnt > 1) { uint32_t kr = Key[r]; for (uint32_t s = 0; s < cnt; s++) rank += Key[Unsorted[off + s]] > kr ? 1u : 0u; }
        Sorted[off + rank] = (Idx)r;
        SPos[r] = (Idx)(off + rank);
    }
    __syncthreads();

    // ---- E: Euler tour (enter r = r, exit r = (M+1)+r, r in 0..M) + weighted pointer-jumping list ranking ------------------
    const uint32_t E = 2 * (M + 1), END = E;
    Idx* nxtA = A.alloc<Idx>(E + 1);
    Idx* nxtB = A.alloc<Idx>(E + 1);
    Idx* dA = A.alloc<Idx>(E + 1);
    Idx* dB = A.alloc<Idx>(E + 1);
    for (uint32_t r = tid; r <= M; r += BLOCK) {
        uint32_t ent = r, ext = (M + 1) + r;
        uint32_t cnt = GrpCnt[r];
        nxtA[ent] = (Idx)(cnt ? (uint32_t)Sorted[GrpOff[r]] : ext);
        dA[ent] = (Idx)(r < M ? (uint32_t)RunTail[r] - (uint32_t)RunHead[r] + 1u : 0u);
        dA[ext] = 0;
        if (r == M) nxtA[ext] = (Idx)END;
        else {
            uint32_t q = Prun[r], sp = SPos[r];
            bool last = sp + 1 == (uint32_t)GrpOff[q] + GrpCnt[q];
            nxtA[ext] = (Idx)(last ? (M + 1) + q : (uint32_t)Sorted[sp + 1]);
        }
    }
    if (tid == 0) { nxtA[END] = (Idx)END; dA[END] = 0; nxtB[END] = (Idx)END; dB[END] = 0; }
    __syncthreads();
    {
        Idx *nc = nxtA, *nn = nxtB, *dc = dA, *dn = dB;
        for (uint32_t span = 1; span < E + 1; span <<= 1) {
            for (uint32_t x = tid; x < E; x += BLOCK) {
                uint32_t nx = nc[x];
                dn[x] = (Idx)((uint32_t)dc[x] + (uint32_t)dc[nx]);     // dc[END] == 0
                nn[x] = nc[nx];                                          // nc[END] == END
            }
            __syncthreads();
            Idx* t = nc; nc = nn; nn = t; t = dc; dc = dn; dn = t;
        }
        dA = dc;   // dA[r] = number of elements from run r to the end of the sequence
    }

    // ---- F: sequence positions, tombstones in sequence order, visible ranks, text ------------------------------------------
    uint8_t* SeqDel = A.alloc<uint8_t>(N + 1);
    const uint32_t NW = (N + 32) / 32;              // bit words covering positions 0..N
    uint32_t* VisBits = A.alloc<uint32_t>(NW + 1);
    Idx* VisPre = A.alloc<Idx>(NW + 1);
    for (uint32_t i = tid; i < n; i += BLOCK) {
        if (Par[i] == NONE) continue;
        uint32_t r = RunOrPos[i];
        uint32_t pos = N - (uint32_t)dA[r] + (i - (uint32_t)RunHead[r]);
        RunOrPos[i] = (Idx)pos;
        SeqDel[pos] = Del[i];
    }
    __syncthreads();
    {
        uint32_t carry = 0;
        for (uint32_t base = 0; base < NW; base += BLOCK) {
            // each thread builds one 32-position word serially from the byte flags of its 32 positions
            uint32_t w = base + tid, bits = 0;
            if (w < NW) {
                uint32_t x0 = w * 32;
                for (uint32_t b = 0; b < 32; b++) { uint32_t x = x0 + b; if (x < N && !SeqDel[x]) bits |= 1u << b; }
                VisBits[w] = bits;
            }
            uint32_t total, ex = block_scan_excl<BLOCK>(__popc(bits), c, total);
            if (w < NW) VisPre[w] = (Idx)(carry + ex);
            carry += total;
        }
        if (tid == 0) c.nvis = carry;
    }
    __syncthreads();
    const uint32_t nvis = c.nvis;
    auto visRank = [&](uint32_t x) -> uint32_t {   // number of visible elements at positions < x   (x in 0..N)
        uint32_t w = x >> 5, b = x & 31;
        return (uint32_t)VisPre[w] + __popc(VisBits[w] & ((1u << b) - 1u));
    };
    {
        unsigned long long d0 = 0, d1 = 0;
        for (uint32_t i = tid; i < n; i += BLOCK) {
            if (Par[i] == NONE || Del[i]) continue;
            uint32_t tok = PT_PAYLOAD_TOKEN(__ldg(&ins[i].payload));
            uint32_t vr = visRank(RunOrPos[i]);
            text_out[vr] = tok;
            digest_add(d0, d1, pt_term_text(vr, tok));
        }
        digest_flush<BLOCK>(c, d0, d1);
    }

    uint32_t nspans = 0;
    if (m == 0) {
        // no marks: one span {} covering all visible text (peritext.ts:392), none if the text is empty
        if (nvis && tid == 0) {
            pt_span s; s.start = 0; s.flags = 0; s.link_attr = PT_ATTR_NONE; s.comment_off = 0;
            span_out[0] = s;
            unsigned long long d0 = 0, d1 = 0;
            digest_add(d0, d1, pt_term_span(0, 0, 0, PT_ATTR_NONE));
            atomicAdd(&c.dig0, d0); atomicAdd(&c.dig1, d1);
        }
        nspans = nvis ? 1u : 0u;
    } else {
        // ---- G: marks ---------------------------------------------------------------------------------------------------
        // G1: rank mark ops by opId: bitmap over the key space + prefix popcount (a counting sort with unique keys)
        const uint32_t KW = (KS + 31) / 32;
        uint32_t* KBits = A.alloc<uint32_t>(KW + 1);
        Idx* KPre = A.alloc<Idx>(KW + 1);
        Idx* ByRank = A.alloc<Idx>(m + 1);
        Idx* MRank = A.alloc<Idx>(m + 1);
        Idx* IvA = A.alloc<Idx>(m + 1);         // element interval [a,b) per mark op; a == b: covers nothing
        Idx* IvB = A.alloc<Idx>(m + 1);
        uint8_t* Bnd = A.alloc<uint8_t>(N + 2);   // boundary flags over element indices 0..N
        uint32_t* BndBits = A.alloc<uint32_t>(NW + 1);
        Idx* SegPre = A.alloc<Idx>(NW + 1);
        fill<uint32_t, BLOCK>(KBits, KW + 1, 0u);
        fill<uint8_t, BLOCK>(Bnd, N + 2, (uint8_t)0);
        __syncthreads();
        for (uint32_t k = tid; k < m; k += BLOCK) {
            uint32_t ctr = mk[k].ctr, actor = mk[k].actor;
            if (badId(ctr, actor)) { fail(PT_LOG_BAD_OPID); continue; }
            uint32_t key = keyOf(ctr, actor);
            uint32_t old = atomicOr(&KBits[key >> 5], 1u << (key & 31));
            if ((old >> (key & 31)) & 1u) fail(PT_LOG_BAD_OPID);
            if (T[key] != NONE) fail(PT_LOG_BAD_OPID);
        }
        __syncthreads();
        if (c.status) { if (tid == 0) { pt_log_result r{}; r.status = c.status; *res = r; } __syncthreads(); return; }
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < KW; base += BLOCK) {
                uint32_t w = base + tid, total;
                uint32_t ex = block_scan_excl<BLOCK>(w < KW ? __popc(KBits[w]) : 0u, c, total);
                if (w < KW) KPre[w] = (Idx)(carry + ex);
                carry += total;
            }
        }
        __syncthreads();
        // G2: boundary slots -> element intervals
        for (uint32_t k = tid; k < m; k += BLOCK) {
            const pt_mark_rec r = mk[k];
            uint32_t key = keyOf(r.ctr, r.actor);
            uint32_t rank = (uint32_t)KPre[key >> 5] + __popc(KBits[key >> 5] & ((1u << (key & 31)) - 1u));
            MRank[k] = (Idx)rank; ByRank[rank] = (Idx)k;
            uint32_t sb = r.bounds & 3u, eb = (r.bounds >> 2) & 3u;
            // a slot is 2*pos + (after ? 1 : 0); NOSLOT: the walk never matches this boundary (peritext.ts:236-241)
            const uint32_t NOSLOT = 0xFFFFFFFFu;
            uint32_t ps = NOSLOT, pe = NOSLOT;
            if (sb <= PT_BOUND_AFTER && !badId(r.start_ctr, r.start_actor)) {
                Idx j = T[keyOf(r.start_ctr, r.start_actor)];
                if (j != NONE) ps = 2u * (uint32_t)RunOrPos[j] + sb;
            }
            if (eb <= PT_BOUND_AFTER && !badId(r.end_ctr, r.end_actor)) {
                Idx j = T[keyOf(r.end_ctr, r.end_actor)];
                if (j != NONE) pe = 2u * (uint32_t)RunOrPos[j] + eb;
            }
            uint32_t a = 0, b = 0;
            if (ps != NOSLOT) {
                if (pe == ps || pe == NOSLOT) pe = 2u * N;        // same slot: start branch wins, never ends (quirk Q2)
                a = (ps + 1u) >> 1; b = (pe + 1u) >> 1; if (b > N) b = N;
                if (a >= b) { a = 0; b = 0; }
            }
            IvA[k] = (Idx)a; IvB[k] = (Idx)b;
            if (a < b) { Bnd[a] = 1; Bnd[b] = 1; }
        }
        __syncthreads();
        uint32_t S;   // number of segment ids: seg(x) in [0, S)
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < NW; base += BLOCK) {
                uint32_t w = base + tid, bits = 0;
                if (w < NW) {
                    uint32_t x0 = w * 32;
                    for (uint32_t b = 0; b < 32; b++) { uint32_t x = x0 + b; if (x <= N && Bnd[x]) bits |= 1u << b; }
                    BndBits[w] = bits;
                }
                uint32_t total, ex = block_scan_excl<BLOCK>(__popc(bits), c, total);
                if (w < NW) SegPre[w] = (Idx)(carry + ex);
                carry += total;
            }
            S = carry + 1;
        }
        __syncthreads();
        auto segOf = [&](uint32_t x) -> uint32_t {    // popcount(boundary bits[0..x]) inclusive
            uint32_t w = x >> 5, b = x & 31;
            return (uint32_t)SegPre[w] + __popc(BndBits[w] & (0xFFFFFFFFu >> (31 - b)));
        };
        // G3: stabbing max per LWW type on an iterative segment tree over segment ids (range atomicMax, point query)
        uint32_t* Tree = A.alloc<uint32_t>(2 * S + 2);
        uint32_t* SegFlags = A.alloc<uint32_t>(S + 1);
        uint32_t* SegLink = A.alloc<uint32_t>(S + 1);
        int* CDiff = A.alloc<int>(S + 2);
        for (uint32_t s = tid; s < S + 1; s += BLOCK) { SegFlags[s] = 0; SegLink[s] = PT_ATTR_NONE; }
        fill<int, BLOCK>(CDiff, S + 2, 0);
        for (uint32_t t = 0; t < 4; t++) {
            if (t == PT_MARK_COMMENT) continue;
            fill<uint32_t, BLOCK>(Tree, 2 * S + 2, 0u);
            __syncthreads();
            for (uint32_t k = tid; k < m; k += BLOCK) {
                uint32_t a = IvA[k], b = IvB[k];
                if (a >= b || ((uint32_t)(mk[k].kind >> 1) & 3u) != t) continue;
                uint32_t v = (uint32_t)MRank[k] + 1u;
                for (uint32_t l = segOf(a) + S, r = segOf(b) + S; l < r; l >>= 1, r >>= 1) {
                    if (l & 1u) atomicMax(&Tree[l++], v);
                    if (r & 1u) atomicMax(&Tree[--r], v);
                }
            }
            __syncthreads();
            const uint32_t bit = t == PT_MARK_STRONG ? PT_SPAN_STRONG : t == PT_MARK_EM ? PT_SPAN_EM : PT_SPAN_LINK;
            for (uint32_t s = tid; s < S; s += BLOCK) {
                uint32_t w = 0;
                for (uint32_t p = s + S; p >= 1; p >>= 1) w = max(w, Tree[p]);
                if (w) {
                    uint32_t kk = ByRank[w - 1];
                    if ((mk[kk].kind & 1u) == 0) {                         // winner is an addMark (peritext.ts:307-311)
                        SegFlags[s] |= bit;
                        if (t == PT_MARK_LINK) SegLink[s] = mk[kk].attr;
                    }
                }
            }
            __syncthreads();
        }
        // G4: `comment` key present iff at least one comment op (add or remove) covers the segment (quirk Q3)
        uint32_t* CompactC = A.alloc<uint32_t>(m + 1);     // indices of non-empty comment ops
        if (tid == 0) c.misc[0] = 0;
        __syncthreads();
        for (uint32_t k = tid; k < m; k += BLOCK) {
            uint32_t a = IvA[k], b = IvB[k];
            if (a >= b || ((uint32_t)(mk[k].kind >> 1) & 3u) != PT_MARK_COMMENT) continue;
            atomicAdd(&CDiff[segOf(a)], 1); atomicAdd(&CDiff[segOf(b)], -1);
            CompactC[atomicAdd(&c.misc[0], 1u)] = k;
        }
        __syncthreads();
        const uint32_t Mc = c.misc[0];
        {
            int carry = 0;
            for (uint32_t base = 0; base < S; base += BLOCK) {
                uint32_t s = base + tid, total;
                int v = s < S ? CDiff[s] : 0;
                uint32_t ex = block_scan_excl<BLOCK>((uint32_t)v, c, total);   // two's complement sums are fine
                int cover = carry + (int)ex + v;
                if (s < S && cover > 0) SegFlags[s] |= PT_SPAN_COMMENT;
                carry += (int)total;
            }
        }
        __syncthreads();

        // ---- H: comment presence pieces (per comment id, LWW by opId) in visible space; comment-induced span heads ----
        const uint32_t HW = nvis / 32 + 1;
        uint32_t* CHead = A.alloc<uint32_t>(HW + 1);
        uint32_t* PcId = A.alloc<uint32_t>(2 * Mc + 1);
        Idx* PcA = A.alloc<Idx>(2 * Mc + 1);
        Idx* PcB = A.alloc<Idx>(2 * Mc + 1);      // PcA == PcB: dead piece
        fill<uint32_t, BLOCK>(CHead, HW + 1, 0u);
        __syncthreads();
        for (uint32_t e = tid; e < 2 * Mc; e += BLOCK) {
            uint32_t ci = e >> 1, which = e & 1u, k = CompactC[ci];
            uint32_t id = mk[k].attr;
            uint32_t x = which ? (uint32_t)IvB[k] : (uint32_t)IvA[k];
            bool dup = false; uint32_t nextEnd = 0xFFFFFFFFu;
            for (uint32_t cj = 0; cj < Mc; cj++) {
                uint32_t j = CompactC[cj];
                if (mk[j].attr != id) continue;
                uint32_t ja = IvA[j], jb = IvB[j];
                if (ja == x && (cj < ci || (cj == ci && 0u < which))) dup = true;
                if (jb == x && (cj < ci || (cj == ci && 1u < which))) dup = true;
                if (ja > x && ja < nextEnd) nextEnd = ja;
                if (jb > x && jb < nextEnd) nextEnd = jb;
            }
            uint32_t va = 0, vb = 0;
            if (!dup && nextEnd != 0xFFFFFFFFu) {
                uint32_t best = 0; bool bestAdd = false;
                for (uint32_t cj = 0; cj < Mc; cj++) {
                    uint32_t j = CompactC[cj];
                    if (mk[j].attr != id) continue;
                    if ((uint32_t)IvA[j] <= x && nextEnd <= (uint32_t)IvB[j]) {
                        uint32_t rk = (uint32_t)MRank[j] + 1u;
                        if (rk > best) { best = rk; bestAdd = (mk[j].kind & 1u) == 0; }
                    }
                }
                if (best && bestAdd) { va = visRank(x); vb = visRank(nextEnd); if (va >= vb) { va = 0; vb = 0; } }
            }
            PcId[e] = id; PcA[e] = (Idx)va; PcB[e] = (Idx)vb;
        }
        __syncthreads();
        for (uint32_t e = tid; e < 2 * Mc; e += BLOCK) {
            uint32_t va = PcA[e], vb = PcB[e], id = PcId[e];
            if (va >= vb) continue;
            bool startTouch = false, endTouch = false;
            for (uint32_t f = 0; f < 2 * Mc; f++) {
                if (PcId[f] != id) continue;
                uint32_t fa = PcA[f], fb = PcB[f];
                if (fa >= fb) continue;
                if (fb == va) startTouch = true;
                if (fa == vb) endTouch = true;
            }
            if (!startTouch) atomicOr(&CHead[va >> 5], 1u << (va & 31));
            if (!endTouch) atomicOr(&CHead[vb >> 5], 1u << (vb & 31));
        }

        // ---- I: spans ---------------------------------------------------------------------------------------------------
        Idx* VisSeg = A.alloc<Idx>(nvis + 1);
        const uint32_t VW = (nvis + 31) / 32;
        uint32_t* HeadBits = A.alloc<uint32_t>(VW + 1);
        Idx* HeadPre = A.alloc<Idx>(VW + 1);
        for (uint32_t i = tid; i < n; i += BLOCK) {
            if (Par[i] == NONE || Del[i]) continue;
            uint32_t pos = RunOrPos[i];
            VisSeg[visRank(pos)] = (Idx)(segOf(pos) );
        }
        __syncthreads();
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < VW; base += BLOCK) {
                uint32_t w = base + tid, bits = 0;
                if (w < VW) {
                    for (uint32_t b = 0; b < 32; b++) {
                        uint32_t v = w * 32 + b;
                        if (v >= nvis) break;
                        bool h;
                        if (v == 0) h = true;
                        else {
                            uint32_t s1 = VisSeg[v - 1], s2 = VisSeg[v];
                            h = ((CHead[v >> 5] >> (v & 31)) & 1u) ||
                                (s1 != s2 && (SegFlags[s1] != SegFlags[s2] || SegLink[s1] != SegLink[s2]));
                        }
                        if (h) bits |= 1u << b;
                    }
                    HeadBits[w] = bits;
                }
                uint32_t total, ex = block_scan_excl<BLOCK>(__popc(bits), c, total);
                if (w < VW) HeadPre[w] = (Idx)(carry + ex);
                carry += total;
            }
            nspans = carry;
        }
        __syncthreads();
        auto headRank = [&](uint32_t v) -> uint32_t {   // number of span heads at visible positions < v  (v in 0..nvis)
            uint32_t w = v >> 5, b = v & 31;
            if (w >= VW) return nspans;
            return (uint32_t)HeadPre[w] + __popc(HeadBits[w] & ((1u << b) - 1u));
        };
        // comment lists per span: count, reserve pool space, fill, sort
        uint32_t* SpanCC = A.alloc<uint32_t>(nspans + 1);
        uint32_t* SpanCO = A.alloc<uint32_t>(nspans + 1);
        uint32_t* SpanCur = A.alloc<uint32_t>(nspans + 1);
        fill<uint32_t, BLOCK>(SpanCC, nspans + 1, 0u);
        fill<uint32_t, BLOCK>(SpanCur, nspans + 1, 0u);
        __syncthreads();
        for (uint32_t e = tid; e < 2 * Mc; e += BLOCK) {
            uint32_t va = PcA[e], vb = PcB[e];
            if (va >= vb) continue;
            for (uint32_t j = headRank(va), j1 = headRank(vb); j < j1; j++) atomicAdd(&SpanCC[j], 1u);
        }
        __syncthreads();
        uint32_t totalC;
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < nspans; base += BLOCK) {
                uint32_t j = base + tid, total;
                uint32_t ex = block_scan_excl<BLOCK>(j < nspans ? SpanCC[j] : 0u, c, total);
                if (j < nspans) SpanCO[j] = carry + ex;
                carry += total;
            }
            totalC = carry;
        }
        if (tid == 0) {
            unsigned long long base = 0;
            if (totalC) {
                base = atomicAdd(P.comment_used, (unsigned long long)totalC);
                if (base + totalC > P.comment_cap) { c.status = PT_LOG_OVERFLOW; base = 0; }
            }
            c.pool_base = base;
        }
        __syncthreads();
        if (c.status) { if (tid == 0) { pt_log_result r{}; r.status = c.status; *res = r; } __syncthreads(); return; }
        uint32_t* pool = P.comment_pool + c.pool_base;
        for (uint32_t e = tid; e < 2 * Mc; e += BLOCK) {
            uint32_t va = PcA[e], vb = PcB[e];
            if (va >= vb) continue;
            for (uint32_t j = headRank(va), j1 = headRank(vb); j < j1; j++)
                pool[SpanCO[j] + atomicAdd(&SpanCur[j], 1u)] = PcId[e];
        }
        __syncthreads();
        {
            unsigned long long d0 = 0, d1 = 0;
            for (uint32_t w = tid; w < VW; w += BLOCK) {
                uint32_t bits = HeadBits[w], j = HeadPre[w];
                while (bits) {
                    uint32_t b = __ffs(bits) - 1; bits &= bits - 1;
                    uint32_t v = w * 32 + b, s = VisSeg[v];
                    uint32_t cnt = SpanCC[j];
                    uint32_t* lst = pool + SpanCO[j];
                    for (uint32_t x = 1; x < cnt; x++) {                   // insertion sort: ascending comment id
                        uint32_t key = lst[x]; uint32_t y = x;
                        while (y > 0 && lst[y - 1] > key) { lst[y] = lst[y - 1]; y--; }
                        lst[y] = key;
                    }
                    pt_span sp; sp.start = v; sp.flags = SegFlags[s] | (cnt << 8); sp.link_attr = SegLink[s];
                    sp.comment_off = cnt ? (uint32_t)(c.pool_base + SpanCO[j]) : 0u;
                    span_out[j] = sp;
                    for (uint32_t x = 0; x < cnt; x++) digest_add(d0, d1, pt_term_comment(j, x, lst[x]));
                    digest_add(d0, d1, pt_term_span(j, sp.start, sp.flags, sp.link_attr));
                    j++;
                }
            }
            digest_flush<BLOCK>(c, d0, d1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        pt_log_result r;
        r.status = A.overflow ? PT_LOG_OVERFLOW : c.status;
        r.n_elems = N; r.n_visible = nvis; r.n_spans = nspans;
        uint64_t t = pt_term_counts(nvis, nspans);
        r.digest[0] = c.dig0 + t; r.digest[1] = c.dig1 + pt_term_hi(t);
        if (r.status) { r.n_elems = r.n_visible = r.n_spans = 0; r.digest[0] = r.digest[1] = 0; }
        *res = r;
    }
    __syncthreads();
}

// Persistent CTAs pull logs from the bin's work queue.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) merge_logs_kernel(const BatchParams P) {
    extern __shared__ __align__(16) char smem_arena[];
    __shared__ BlockCtx<BLOCK> ctx;
    for (;;) {
        if (threadIdx.x == 0) ctx.work = atomicAdd(P.work_counter, 1u);
        __syncthreads();
        const uint32_t w = ctx.work;
        __syncthreads();
        if (w >= P.n_work) break;
        const uint32_t li = P.order[w];
        const pt_log_desc& L = P.desc[li];
        const bool small = L.n_insdel < 32000u && L.n_mark < 32000u;
        if (small) merge_one_log<uint16_t, BLOCK>(P, li, ctx, smem_arena);
        else merge_one_log<uint32_t, BLOCK>(P, li, ctx, smem_arena);
    }
}

}  // namespace ptk
