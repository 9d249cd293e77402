/* peritext_b200.h — C-ABI of the B200 batch CRDT-merge engine (libperitext_b200.so).
 *
 * Drop-in boundary for ONE hot path of inkandswitch/peritext: replaying op logs and flattening the
 * documents to formatted spans, i.e. what the reference does in
 *     Micromerge.applyChange -> applyOp          (reference src/micromerge.ts:499-608)
 *     applyListInsert / applyListUpdate           (src/micromerge.ts:614-724)
 *     applyAddRemoveMark                          (src/peritext.ts:154-249)
 *     getTextWithFormatting / addCharactersToSpans(src/peritext.ts:337-455)
 * for MANY (document, replica) op logs at once.  The reference has no FFI seam (it is pure TypeScript);
 * these entry points are what an N-API addon behind a `Micromerge` facade binds.
 *
 * Conventions: plain C types only, no C++/torch types; every function returns a pt_status code (never
 * throws); device work is enqueued on the CUDA stream handle given at create time; host buffers passed
 * in are owned by the caller.  PAGEABLE input buffers may be freed as soon as the call returns (the call
 * waits for the copy); buffers in PINNED (page-locked) memory are read asynchronously by the copy engine and
 * must stay valid until the next synchronising call on the handle (pt_batch_sync, pt_batch_download*,
 * pt_batch_last_merge_ms, the next pt_batch_upload*).
 */
#ifndef PERITEXT_B200_H
#define PERITEXT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Packed op records (host and device layout are identical; little endian).
 *
 * One LOG = the ops one replica applied to one text list, in that replica's arrival order
 * (the concatenation of `change.ops` over its applyChange calls, src/micromerge.ts:513).
 * opIds "ctr@actor" (src/micromerge.ts:488) are packed as (ctr, actor_rank) where actor_rank is the rank
 * of the actorId among the log's actors in JS string order (UTF-16 code units), so that
 * compareOpIds (src/micromerge.ts:812-827) == compare (ctr, actor_rank) lexicographically.
 * ---------------------------------------------------------------------------------------------- */

/* Insert (`action:"set", insert:true`, src/micromerge.ts:150-160) or delete (`action:"del"`, :162-168). 16 B. */
typedef struct pt_insdel_rec {
    uint32_t ctr;       /* opId counter (>=1)                                                     */
    uint32_t ref_ctr;   /* insert: reference elemId counter, 0 = HEAD; delete: target elemId ctr */
    uint16_t actor;     /* opId actor rank                                                        */
    uint16_t ref_actor; /* reference / target elemId actor rank                                   */
    uint32_t payload;   /* bits31:30 kind (PT_KIND_*); insert: bit29 = value-pool flag, bits28:0 =
                           Unicode code point (single-character value) or value-pool index        */
} pt_insdel_rec;

#define PT_KIND_INSERT 0u
#define PT_KIND_DELETE 1u
#define PT_PAYLOAD_KIND(p) ((uint32_t)(p) >> 30)
#define PT_PAYLOAD_TOKEN(p) ((uint32_t)(p) & 0x3FFFFFFFu)
#define PT_TOKEN_POOLED 0x20000000u

/* addMark / removeMark (src/peritext.ts:25-65). 32 B. */
typedef struct pt_mark_rec {
    uint32_t ctr;         /* opId counter                                                         */
    uint16_t actor;       /* opId actor rank                                                      */
    uint8_t  kind;        /* bit0: 0 addMark, 1 removeMark; bits2:1 mark type (PT_MARK_*)          */
    uint8_t  bounds;      /* bits1:0 start boundary type, bits3:2 end boundary type (PT_BOUND_*)   */
    uint32_t start_ctr;   /* start.elemId counter (before/after only)                              */
    uint32_t end_ctr;     /* end.elemId counter (before/after only)                                */
    uint16_t start_actor; /* start.elemId actor rank                                               */
    uint16_t end_actor;   /* end.elemId actor rank                                                 */
    uint32_t attr;        /* link: interned url id; comment: comment-id rank (JS string order of
                             the id, per batch); PT_ATTR_NONE otherwise                           */
    uint32_t arrival;     /* number of ins/del records of this log that arrived before this op   */
    uint32_t reserved;    /* 0                                                                     */
} pt_mark_rec;

/* Mark types in ALL_MARKS order (src/schema.ts:125). strong/em: inclusive, single; comment:
 * non-inclusive, allowMultiple; link: non-inclusive, single (src/schema.ts:45-96). */
#define PT_MARK_STRONG 0u
#define PT_MARK_EM 1u
#define PT_MARK_COMMENT 2u
#define PT_MARK_LINK 3u
#define PT_BOUND_BEFORE 0u
#define PT_BOUND_AFTER 1u
#define PT_BOUND_START_OF_TEXT 2u
#define PT_BOUND_END_OF_TEXT 3u
#define PT_ATTR_NONE 0xFFFFFFFFu

/* Per-log descriptor. 32 B. */
typedef struct pt_log_desc {
    uint64_t insdel_off; /* first pt_insdel_rec of the log in the batch's insdel array            */
    uint64_t mark_off;   /* first pt_mark_rec of the log in the batch's mark array                */
    uint32_t n_insdel;
    uint32_t n_mark;
    uint32_t n_actors;   /* actor ranks are < n_actors                                             */
    uint32_t max_ctr;    /* every ctr in the log is in [1, max_ctr]                                */
} pt_log_desc;

/* ------------------------------------------------------------------------------------------------
 * Run-compressed wire format (optional, for the host -> device leg).  The ops of one Change are
 * consecutive (`opId = ++maxOp`, src/micromerge.ts:487) and a typed run chains every insert to the
 * previous one (`elemId = result`, :351-359), so the ins/del stream compresses to RUNS:
 *   insert run: `count` inserts with opIds (ctr0 + k, actor); the first references (ref_ctr, ref_actor),
 *               each following one references its predecessor; values = `count` tokens of the token stream
 *   delete run: `count` deletes with opIds (ctr0 + k, actor) of the elements (ref_ctr + k, ref_actor)
 * pt_batch_upload_runs expands runs to pt_insdel_rec on the device; results are identical.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_run_rec {
    uint32_t ctr0;
    uint32_t ref_ctr;
    uint16_t actor;
    uint16_t ref_actor;
    uint32_t kind_count; /* bits31:30 kind (PT_KIND_*), bits29:0 count (>= 1) */
} pt_run_rec;

typedef struct pt_packed_runs {
    uint32_t n_logs;
    const pt_log_desc* logs;      /* [n_logs] — the EXPANDED layout (insdel_off / n_insdel count records)      */
    const uint64_t* run_off;      /* [n_logs + 1] first run of each log                                         */
    const uint64_t* tok_off;      /* [n_logs + 1] first insert token of each log                                */
    const pt_run_rec* runs;       /* [run_off[n_logs]]                                                          */
    const uint32_t* tokens;       /* [tok_off[n_logs]] PT_PAYLOAD_TOKEN values of the inserts, in record order  */
    const pt_mark_rec* marks;     /* [n_mark_total] (not compressed)                                            */
    uint64_t n_insdel_total;      /* expanded record count                                                      */
    uint64_t n_mark_total;
} pt_packed_runs;

/* ------------------------------------------------------------------------------------------------
 * Change table (optional): what Micromerge.applyChange checks BEFORE it applies a change
 * (src/micromerge.ts:499-511): seq == clock[actor] + 1, and every deps[a] <= clock[a] (a missing or
 * zero clock entry fails).  One pt_change_rec per Change the replica applied, in arrival order; the
 * admission pre-pass (pt_batch_upload_changes) validates every log on the device and reports the
 * first rejected change as the log's status instead of throwing mid-batch; such a log is not merged.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_change_rec {   /* 16 B */
    uint32_t seq;     /* change.seq                                                                */
    uint16_t actor;   /* change.actor, rank among the log's actors (same ranks as the op records)  */
    uint16_t n_deps;  /* entries of change.deps                                                    */
    uint32_t dep_off; /* first pt_dep_rec of the change, relative to the log's dep_off             */
    uint32_t n_ops;   /* ops of the change that target the log's text list (informational)         */
} pt_change_rec;
typedef struct pt_dep_rec {      /* 8 B */
    uint32_t seq;
    uint16_t actor;
    uint16_t reserved;
} pt_dep_rec;
typedef struct pt_change_desc {  /* 24 B */
    uint64_t change_off;
    uint64_t dep_off;
    uint32_t n_changes;
    uint32_t n_deps;
} pt_change_desc;
typedef struct pt_change_table {
    uint32_t n_logs;              /* must equal the uploaded batch's n_logs */
    const pt_change_desc* logs;
    const pt_change_rec* changes;
    uint64_t n_changes_total;
    const pt_dep_rec* deps;
    uint64_t n_deps_total;
} pt_change_table;

/* ------------------------------------------------------------------------------------------------
 * Compact wire format (optional, for the host -> device leg): the same records in half the bytes.
 * Usable for a log with max_ctr < 65536, n_insdel < 65536 and n_actors <= 16 (any document the
 * benchmark shapes produce); value tokens must fit 22 bits (any Unicode code point, or a value-pool
 * index < 2^21).  Records keep their positions (the descriptors are those of the expanded layout);
 * pt_batch_upload_compact expands them to pt_insdel_rec / pt_mark_rec on the device, elementwise.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_insdel_c8 {   /* 8 B */
    uint16_t ctr, ref_ctr;
    uint32_t w;                 /* bits3:0 actor, 7:4 ref_actor, 9:8 kind, 31:10 token (bit21 of the token = value-pool flag) */
} pt_insdel_c8;
typedef struct pt_mark_c16 {    /* 16 B */
    uint16_t ctr, start_ctr, end_ctr, arrival;
    uint32_t attr;
    uint32_t w;                 /* bits3:0 actor, 7:4 start_actor, 11:8 end_actor, 14:12 kind, 18:15 bounds */
} pt_mark_c16;
typedef struct pt_packed_compact {
    uint32_t n_logs;
    const pt_log_desc* logs;       /* the EXPANDED layout */
    const pt_insdel_c8* insdel;    /* [n_insdel_total] */
    uint64_t n_insdel_total;
    const pt_mark_c16* marks;      /* [n_mark_total] */
    uint64_t n_mark_total;
} pt_packed_compact;

/* A host-side batch of logs (SoA). */
typedef struct pt_packed_ops {
    uint32_t n_logs;
    const pt_log_desc* logs;     /* [n_logs]                 */
    const pt_insdel_rec* insdel; /* [sum n_insdel]           */
    uint64_t n_insdel_total;
    const pt_mark_rec* marks;    /* [sum n_mark]             */
    uint64_t n_mark_total;
} pt_packed_ops;

/* ------------------------------------------------------------------------------------------------
 * Results
 * ---------------------------------------------------------------------------------------------- */

/* Per-log status; each maps to a reference throw site. */
#define PT_LOG_OK 0u
#define PT_LOG_ELEM_NOT_FOUND 1u /* "List element not found" src/micromerge.ts:752                 */
#define PT_LOG_BAD_OPID 2u       /* ctr/actor outside the descriptor's bounds, or duplicate opId   */
#define PT_LOG_BAD_KIND 3u       /* record kind not insert/delete                                  */
#define PT_LOG_OVERFLOW 4u       /* an engine capacity (comment pool / scratch) was exceeded       */
#define PT_LOG_CYCLE 5u          /* reference elemId does not causally precede the insert (a peer chose a startOp below
                                    an element it references: the reference would still merge it; this engine's
                                    closed form needs Lamport counters, src/micromerge.ts:487,511)                */
#define PT_LOG_SEQ_GAP 6u        /* admission: "Expected sequence number ..." src/micromerge.ts:501-504; n_elems =
                                    index of the rejected change in the log                                        */
#define PT_LOG_MISSING_DEP 7u    /* admission: "Missing dependency ..." src/micromerge.ts:505-509; n_elems = index of
                                    the rejected change                                                            */

/* Per-log result header. 32 B. */
typedef struct pt_log_result {
    uint32_t status;    /* PT_LOG_*                                                                */
    uint32_t n_elems;   /* list elements incl. tombstones (metadata.length)                        */
    uint32_t n_visible; /* visible elements (text.length, src/micromerge.ts:657)                   */
    uint32_t n_spans;   /* FormatSpanWithText count (src/peritext.ts:361)                          */
    uint64_t digest[2]; /* 128-bit digest of (text tokens, spans, marks) — convergence check       */
} pt_log_result;

/* One formatted span (src/peritext.ts:35-38). 16 B. Text of span j = tokens [start_j, start_{j+1}). */
typedef struct pt_span {
    uint32_t start;       /* index of the span's first visible element                             */
    uint32_t flags;       /* bit0 strong, bit1 em, bit2 link present, bit3 `comment` key present;
                             bits31:8 number of comment ids                                        */
    uint32_t link_attr;   /* url id if bit2, else PT_ATTR_NONE                                     */
    uint32_t comment_off; /* first comment id of the span in the comment pool (ascending ids)      */
} pt_span;

#define PT_SPAN_STRONG 1u
#define PT_SPAN_EM 2u
#define PT_SPAN_LINK 4u
#define PT_SPAN_COMMENT 8u
#define PT_SPAN_NCOMMENTS(f) ((uint32_t)(f) >> 8)

/* Host view of a merged batch; pointers are engine-owned pinned host memory, valid until the next
 * pt_batch_upload/pt_batch_destroy on the handle. The outputs are PACKED on the device before the copy: log i's
 * tokens are text[text_off[i] .. text_off[i+1]), its spans spans[span_off[i] .. span_off[i+1]). */
typedef struct pt_spans_view {
    uint32_t n_logs;
    const pt_log_result* results; /* [n_logs]                                                      */
    const uint64_t* text_off;     /* [n_logs + 1] packed offsets (exclusive scan of n_visible)      */
    const uint64_t* span_off;     /* [n_logs + 1] packed offsets (exclusive scan of n_spans)        */
    const uint32_t* text;         /* visible element value tokens (PT_PAYLOAD_TOKEN)               */
    const pt_span* spans;
    const uint32_t* comment_pool;
    uint64_t comment_pool_used;
    const uint32_t* seq;          /* NULL unless PT_FLAG_EMIT_SEQUENCE: per log (offset seq_off[i], n_elems entries) the element
                                     sequence incl. tombstones (the reference's `metadata` array, src/micromerge.ts:255):
                                     bits29:0 = index of the element's insert record in the log's ins/del records,
                                     bit30 = the element's markOpsAfter slot is defined (src/micromerge.ts:784),
                                     bit31 = deleted                                                                  */
    const uint64_t* seq_off;      /* [n_logs] offsets into seq (capacity layout: running sum of n_insdel); NULL without seq */
    uint64_t comment_pool_needed; /* comment-pool entries the whole batch needs; > the pool's capacity iff some logs
                                     reported PT_LOG_OVERFLOW: call pt_batch_set_comment_pool(needed) and merge again  */
} pt_spans_view;

/* Engine limits / tuning. Zero-initialise for defaults. */
typedef struct pt_limits {
    uint64_t comment_pool_entries; /* 0: 64 x number of mark ops in the batch (+slack)            */
    uint32_t flags;                /* PT_FLAG_*                                                     */
    uint32_t patch_pool_items;     /* 0: 4 x number of op records in the batch (+slack)            */
    uint32_t reserved[4];
} pt_limits;
#define PT_FLAG_EMIT_SEQUENCE 1u   /* also emit the element sequence (needed by op generation / cursors on the host) */
#define PT_FLAG_EMIT_PATCHES 2u    /* also derive the Patch stream of every op on the device (implies EMIT_SEQUENCE) */

/* ------------------------------------------------------------------------------------------------
 * Patch stream (PT_FLAG_EMIT_PATCHES): what Micromerge.applyChange returns for every op of a log, given the
 * log's arrival order (src/micromerge.ts:659-671, 689-703; src/peritext.ts:175-220, 251-281).
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_patch_rec {   /* one per ins/del record, at the record's offset. 16 B */
    uint32_t index;     /* bits30:0: Patch.index (visible elements left of the op's element at apply time);
                           bit31: the op emits a patch (inserts always; a delete only if it is the element's first) */
    uint32_t flags;     /* insert: the `marks` of the patch = marks inherited from the left neighbour at apply time
                           (getActiveMarksAtIndex, src/peritext.ts:328): PT_SPAN_* bits, bits31:8 number of comment ids  */
    uint32_t link_attr; /* url id if PT_SPAN_LINK                                                                 */
    uint32_t reserved;
} pt_patch_rec;
typedef struct pt_patch_item {  /* pool entry, any order. 16 B */
    uint32_t log;       /* log index                                                                             */
    uint32_t tag;       /* bit31 set: mark patch of mark record (tag & 0x7FFFFFFF): a = startIndex, b = endIndex;
                           bit31 clear: comment id `a` in the `marks` of the insert patch of ins/del record `tag`       */
    uint32_t a, b;
} pt_patch_item;
typedef struct pt_patch_view {
    const pt_patch_rec* recs;      /* [n_insdel_total]                                                            */
    const pt_patch_item* items;    /* [n_items]                                                                   */
    uint64_t n_items;
    uint64_t n_items_needed;       /* > the pool's capacity: call pt_batch_set_patch_pool(needed) and merge again */
    const uint32_t* status;        /* [n_logs] 0: computed; 1: not computed (log too large for the device patch kernel or
                                      merge failed) — derive on the host                                          */
} pt_patch_view;

typedef enum pt_status {
    PT_OK = 0,
    PT_ERR_INVALID = 1,   /* bad argument                                                          */
    PT_ERR_CUDA = 2,      /* CUDA runtime error (see pt_last_error)                                */
    PT_ERR_NO_DEVICE = 3, /* no usable sm_100 device                                               */
    PT_ERR_STATE = 4,     /* call out of order (e.g. merge before upload)                          */
    PT_ERR_NOMEM = 5
} pt_status;

typedef struct pt_batch pt_batch; /* opaque; one per (GPU, batch); not thread-safe per handle */

/* Create an engine handle on CUDA device `device`; work is enqueued on `cuda_stream`
 * (a cudaStream_t / CUstream passed as void*, NULL = legacy default stream). */
int pt_batch_create(int device, const pt_limits* limits, void* cuda_stream, pt_batch** out);

/* Copy a packed batch host -> device (asynchronous on the handle's stream for pinned inputs, see the buffer
 * rule above; the descriptors are staged through engine-owned pinned memory). Waits for the handle's previous
 * work first. Replaces any previous batch.
 * This is the H2D leg of Micromerge.applyChange's input (src/micromerge.ts:499). */
int pt_batch_upload(pt_batch*, const pt_packed_ops* host_ops);

/* Same as pt_batch_upload for the run-compressed form: copies runs / tokens / marks host -> device and expands the runs
 * to pt_insdel_rec records on the device (one small kernel).  Typically 2-3x fewer bytes over PCIe. */
int pt_batch_upload_runs(pt_batch*, const pt_packed_runs* host_runs);

/* Host helper: compress a packed batch into runs.  Call once with runs == NULL / tokens == NULL to get the counts
 * (*n_runs, *n_tokens), then with buffers of that size; run_off / tok_off need n_logs + 1 entries. */
int pt_compress_runs(const pt_packed_ops* ops, uint64_t* run_off, uint64_t* tok_off, pt_run_rec* runs, uint32_t* tokens,
                     uint64_t* n_runs, uint64_t* n_tokens);

/* Attach the change table of the uploaded batch (host arrays, copied): the next pt_batch_merge first runs the
 * admission pre-pass on the device; logs with a sequence gap / missing dependency get PT_LOG_SEQ_GAP /
 * PT_LOG_MISSING_DEP and are skipped by the merge (the reference throws before mutating, src/micromerge.ts:501-509).
 * Call after pt_batch_upload*; a new upload drops the table. */
int pt_batch_upload_changes(pt_batch*, const pt_change_table* host_changes);

/* Host helper (multithreaded): convert a packed batch to the compact wire format into caller-provided arrays of
 * n_insdel_total / n_mark_total entries.  PT_ERR_INVALID (nothing useful written) if some log is not representable. */
int pt_compact_ops(const pt_packed_ops* ops, pt_insdel_c8* insdel_out, pt_mark_c16* marks_out, int threads /* 0 = all cores */);

/* Same as pt_batch_upload for the compact form: half the bytes over PCIe, expanded on the device (two elementwise kernels). */
int pt_batch_upload_compact(pt_batch*, const pt_packed_compact* host_compact);

/* Adopt a batch that is ALREADY RESIDENT in device memory (pointers are device pointers owned by the
 * caller, e.g. torch tensors); only the descriptors are read on the host. */
int pt_batch_adopt_device(pt_batch*, const pt_packed_ops* host_desc_device_arrays);

/* Enqueue the merge: op-log apply + flatten for every log of the batch (the replacement for the
 * applyOp loop src/micromerge.ts:513 and getTextWithFormatting src/peritext.ts:337). Asynchronous. */
int pt_batch_merge(pt_batch*);

/* Block until the handle's stream is idle. */
int pt_batch_sync(pt_batch*);

/* Enqueue the device -> host copies of the result arrays without waiting (pinned destinations); a following
 * pt_batch_download only waits.  Lets a caller overlap this handle's download with other handles' work. */
int pt_batch_download_begin(pt_batch*);

/* Copy results device -> host (pinned) and return a view. Synchronises the stream. */
int pt_batch_download(pt_batch*, pt_spans_view* out);

/* PT_FLAG_EMIT_PATCHES: copy the Patch stream of the last merge device -> host and return a view (engine-owned pinned
 * memory, valid until the next upload / destroy). Synchronises. */
int pt_batch_download_patches(pt_batch*, pt_patch_view* out);
int pt_batch_set_patch_pool(pt_batch*, uint64_t items);

/* Batched index -> element resolution on the materialised documents (PT_FLAG_EMIT_SEQUENCE, after a merge): what
 * op generation and cursors need (getListElementId, src/micromerge.ts:762-805; getCursor :465).  Query k asks log
 * `log` for its `index`-th visible element; with PT_QUERY_LOOK_AFTER_TOMBSTONES the answer moves to the LAST following
 * tombstone whose markOpsAfter slot is defined (:775-797, the rule Micromerge.change uses for insert positions).
 * Answers: index of the element's insert record in the log's ins/del records, PT_ELEM_NOT_FOUND if the index is out of
 * bounds ("List index out of bounds", :804).  One warp per query on the device; synchronises. */
typedef struct pt_elem_query { uint32_t log; uint32_t index; uint32_t flags; uint32_t reserved; } pt_elem_query;
#define PT_QUERY_LOOK_AFTER_TOMBSTONES 1u
#define PT_ELEM_NOT_FOUND 0xFFFFFFFFu
int pt_batch_query_elements(pt_batch*, const pt_elem_query* queries, uint32_t n, uint32_t* record_index_out);

/* Copy only the per-log result headers (status, counts, digest). Synchronises the stream. */
int pt_batch_download_results(pt_batch*, pt_log_result* out, uint32_t n_logs);

/* Device pointer to the per-log result headers ([n_logs] pt_log_result) — for the multi-GPU digest
 * all-gather without a host round trip. */
int pt_batch_device_results(pt_batch*, void** dev_ptr, uint32_t* n_logs);

/* Number of kernel launches the handle has enqueued so far (bench `gpu_launches`). */
uint64_t pt_batch_launch_count(const pt_batch*);

/* Diagnostics of the last merge: out[0] = logs materialised entirely in shared memory, out[1] = logs that were
 * restarted on the spill-capable path (working set larger than the bin's shared-memory budget), out[2] = logs deferred on
 * the device to a kernel variant with a larger budget, out[3] = comment-pool entries the batch needs. Synchronises. */
int pt_batch_stats(pt_batch*, uint64_t out[4]);

/* Resize the comment pool (entries of 4 bytes) of the current batch and of later uploads.  Logs that find the pool full
 * report PT_LOG_OVERFLOW (which ones depends on scheduling); pt_spans_view.comment_pool_needed / pt_batch_stats out[3]
 * give the batch's exact demand, so ONE re-merge after pt_batch_set_comment_pool(needed) always succeeds and is
 * deterministic.  Synchronises. */
int pt_batch_set_comment_pool(pt_batch*, uint64_t entries);

/* Time of the last pt_batch_merge on the device, in milliseconds (CUDA events recorded on the
 * handle's stream around the launches); <0 if not available. Synchronises. */
float pt_batch_last_merge_ms(pt_batch*);

void pt_batch_destroy(pt_batch*);

/* ------------------------------------------------------------------------------------------------
 * Native wire-format ingest (host, multithreaded): JSON text of the reference's Change objects
 * (src/micromerge.ts:60-71; the `queues` of the traces/ files) -> packed batch + change table + string pools.
 * logs_json[i] = UTF-8 JSON array of the Change objects replica i applied, in arrival order.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pt_ingest pt_ingest;
#define PT_POOL_VALUES 0        /* multi-character element values, UTF-16LE; token = PT_TOKEN_POOLED | index          */
#define PT_POOL_LINK_ATTRS 1    /* link attrs as canonical JSON (UTF-8); pt_mark_rec.attr = index                      */
#define PT_POOL_COMMENT_IDS 2   /* comment ids, UTF-16LE, in rank order (JS string order); pt_mark_rec.attr = rank     */
#define PT_POOL_COMMENT_ATTRS 3 /* the first-seen attrs object of each comment id as canonical JSON, same order        */
#define PT_POOL_ACTORS 4        /* actor ids UTF-16LE, rank order per log; per_log_first[i] .. per_log_first[i+1]       */
#define PT_POOL_COUNTERS 5      /* dense counter rank -> original counter (u64 each) of logs whose counters were
                                   re-ranked; empty range otherwise; per_log_first as above                            */
int pt_ingest_create(pt_ingest** out);
int pt_ingest_parse(pt_ingest*, const char* const* logs_json, const uint64_t* lens, uint32_t n_logs, int threads /* 0 = all cores */);
/* Views into the parsed batch (valid until the next pt_ingest_parse / pt_ingest_destroy). */
int pt_ingest_packed(pt_ingest*, pt_packed_ops* ops, pt_change_table* changes);
int pt_ingest_pool(pt_ingest*, int kind, const uint8_t** data, const uint64_t** offsets /* [count + 1] */, uint64_t* count,
                   const uint64_t** per_log_first /* may be NULL */);
const char* pt_ingest_error(pt_ingest*);
void pt_ingest_destroy(pt_ingest*);

const char* pt_strerror(int status);
const char* pt_last_error(void); /* thread-local detail string of the last failing call */
const char* pt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PERITEXT_B200_H */
