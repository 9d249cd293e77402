/* Plain-C caller of the drop-in boundary (include/peritext_b200.h): two replicas of "abc" with a concurrent insert and a
 * bold mark, merged on the GPU; checks convergence (digests), text and spans.  Built and run by tests/test_gpu_c_abi.py. */
#include <stdio.h>
#include <string.h>
#include "peritext_b200.h"

#define INS(ctr, actor, rctr, ractor, ch) {ctr, rctr, actor, ractor, (PT_KIND_INSERT << 30) | (uint32_t)(ch)}

int main(void) {
    /* actors: rank 0 = "doc1", rank 1 = "doc2".  doc1: 2,3,4 = a,b,c.  doc1 inserts X after a (5@0); doc2 inserts Y after a (5@1);
     * doc1 bolds [a..b] (6@0: start before 2@0, end before 4@0 -> inclusive mark).  Replica 0 applies doc1's ops first, replica 1 doc2's. */
    pt_insdel_rec r0[] = { INS(2,0,0,0,'a'), INS(3,0,2,0,'b'), INS(4,0,3,0,'c'), INS(5,0,2,0,'X'), INS(5,1,2,0,'Y') };
    pt_insdel_rec r1[] = { INS(2,0,0,0,'a'), INS(3,0,2,0,'b'), INS(4,0,3,0,'c'), INS(5,1,2,0,'Y'), INS(5,0,2,0,'X') };
    pt_insdel_rec ins[10];
    memcpy(ins, r0, sizeof r0); memcpy(ins + 5, r1, sizeof r1);
    pt_mark_rec mk[2];
    memset(mk, 0, sizeof mk);
    for (int k = 0; k < 2; k++) {
        mk[k].ctr = 6; mk[k].actor = 0; mk[k].kind = (PT_MARK_STRONG << 1) | 0;
        mk[k].bounds = PT_BOUND_BEFORE | (PT_BOUND_BEFORE << 2);
        mk[k].start_ctr = 2; mk[k].start_actor = 0; mk[k].end_ctr = 4; mk[k].end_actor = 0;
        mk[k].attr = PT_ATTR_NONE; mk[k].arrival = k == 0 ? 4 : 5;
    }
    pt_log_desc logs[2] = { {0, 0, 5, 1, 2, 6}, {5, 1, 5, 1, 2, 6} };
    pt_packed_ops ops = {2, logs, ins, 10, mk, 2};

    pt_batch* b = NULL;
    if (pt_batch_create(0, NULL, NULL, &b) != PT_OK) { fprintf(stderr, "create: %s\n", pt_last_error()); return 2; }
    pt_spans_view v;
    if (pt_batch_upload(b, &ops) || pt_batch_merge(b) || pt_batch_download(b, &v)) { fprintf(stderr, "run: %s\n", pt_last_error()); return 3; }
    int bad = 0;
    for (int i = 0; i < 2; i++) {
        const pt_log_result* r = &v.results[i];
        printf("log %d: status %u elems %u visible %u spans %u text ", i, r->status, r->n_elems, r->n_visible, r->n_spans);
        for (uint32_t t = 0; t < r->n_visible; t++) putchar((int)v.text[v.text_off[i] + t]);
        for (uint32_t s = 0; s < r->n_spans; s++) printf(" [%u:%s]", v.spans[v.span_off[i] + s].start, (v.spans[v.span_off[i] + s].flags & PT_SPAN_STRONG) ? "strong" : "-");
        printf("\n");
        /* RGA: children of `a` in descending opId: 5@doc2 (Y) before 5@doc1 (X), then b  => aYXbc ; bold covers a..b incl. the inserts */
        if (r->status != PT_LOG_OK || r->n_visible != 5 || r->n_spans != 2) bad = 1;
        if (memcmp(&v.text[v.text_off[i]], (uint32_t[]){'a','Y','X','b','c'}, 20) != 0) bad = 1;
    }
    if (memcmp(v.results[0].digest, v.results[1].digest, 16) != 0) bad = 1;
    printf("launches %llu, merge %.3f ms, %s\n", (unsigned long long)pt_batch_launch_count(b), pt_batch_last_merge_ms(b), bad ? "MISMATCH" : "converged");
    pt_batch_destroy(b);
    return bad;
}
