/* pt_digest.h — definition of the 128-bit per-log digest in pt_log_result.digest.
 *
 * The digest is this engine's stand-in for the fuzz harness's `assert.deepStrictEqual(leftText, rightText)`
 * (reference test/fuzz.ts:278): two replicas converged iff their digests are equal.  Every covered item
 * contributes one position-salted 64-bit term t; digest[0] = SUM of the terms (mod 2^64), digest[1] = XOR of
 * rotl(t, 23).  Both lanes are commutative, so the digest can be reduced in any order (warp shuffles, atomics)
 * and still be deterministic.  Covered: every visible token with its index, every span (index, start,
 * flags incl. comment count, link id) and every comment id with (span index, ordinal).  Pool OFFSETS are
 * not covered (they depend on allocation order).
 */
#ifndef PT_DIGEST_H
#define PT_DIGEST_H
#include <stdint.h>

#if defined(__CUDACC__)
#define PT_HD __host__ __device__ __forceinline__
#else
#define PT_HD static inline
#endif

PT_HD uint64_t pt_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31; return z;
}
PT_HD uint64_t pt_term_text(uint32_t i, uint32_t token) {
    return pt_mix64((((uint64_t)i << 32) | token) + 0x9E3779B97F4A7C15ull);
}
PT_HD uint64_t pt_term_span(uint32_t j, uint32_t start, uint32_t flags, uint32_t link) {
    return pt_mix64(pt_mix64((((uint64_t)j << 32) | start) ^ 0xA5A5A5A55A5A5A5Aull) + (((uint64_t)flags << 32) | link));
}
PT_HD uint64_t pt_term_comment(uint32_t j, uint32_t k, uint32_t id) {
    return pt_mix64(pt_mix64((((uint64_t)j << 32) | k) ^ 0x5BD1E9955BD1E995ull) + id);
}
PT_HD uint64_t pt_term_counts(uint32_t n_visible, uint32_t n_spans) {
    return pt_mix64((((uint64_t)n_visible << 32) | n_spans) ^ 0xC3C3C3C33C3C3C3Cull);
}
/* second lane of a term (XOR-accumulated) */
PT_HD uint64_t pt_term_hi(uint64_t term) { return (term << 23) | (term >> 41); }

#endif
