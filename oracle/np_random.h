/* np_random.h -- TEST INFRASTRUCTURE (CPU oracle).  Not part of the product path.
 *
 * Restatement of the exact arithmetic behind the four `self.np_random.choice(...)` calls of the
 * reference (rware/warehouse.py:781, 788, 797, 916).  `np_random` there is
 * numpy.random.Generator(PCG64(SeedSequence(seed))) (gymnasium/utils/seeding.py: np_random, restated
 * in oracle/refharness/gymnasium_standin/gymnasium/utils/seeding.py).  numpy is a third-party,
 * un-vendored dependency of the reference (setup.py:24-29, unpinned; 2.3.5 in this image), so the
 * algorithm is restated from its published sources:
 *   - SeedSequence:            numpy/random/bit_generator.pyx (pool mixing, generate_state)
 *   - PCG64 (XSL-RR 128/64):   numpy/random/src/pcg64/pcg64.h  (setseq step, next64, buffered next32)
 *   - bounded integers:        numpy/random/src/distributions/distributions.c
 *                              (buffered_bounded_lemire_uint32, random_bounded_uint64[_fill])
 *   - Generator.choice:        numpy/random/_generator.pyx (Floyd sampling + _shuffle_int tail shuffle)
 * Pinned against the real numpy in this container by tests/test_oracle_np_random.py and against
 * reference trajectories by tests/golden/ fixtures.
 */
#ifndef ORC_NP_RANDOM_H
#define ORC_NP_RANDOM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct np_pcg64 {
    uint64_t state_lo, state_hi; /* 128-bit LCG state */
    uint64_t inc_lo, inc_hi;     /* 128-bit stream increment (odd) */
    uint32_t has_uint32;         /* buffered upper half of the last 64-bit draw */
    uint32_t uinteger;
} np_pcg64;

/* Generator(PCG64(SeedSequence(seed))) for a non-negative integer seed < 2^64 */
void np_seed(np_pcg64 *g, uint64_t seed);
uint64_t np_next64(np_pcg64 *g);
uint32_t np_next32(np_pcg64 *g);
/* uniform integer in [0, rng] (inclusive), rng < 2^32-1: buffered_bounded_lemire_uint32 */
uint32_t np_bounded32(np_pcg64 *g, uint32_t rng);
/* Generator.integers(0, high, size=cnt) / choice(a, size=cnt, replace=True) index stream */
void np_integers(np_pcg64 *g, uint32_t high_excl, int cnt, int64_t *out);
/* Generator.choice(pop, size, replace=False) index stream (pop <= 10000 => Floyd + shuffle;
 * larger populations take numpy's tail-shuffle branch, also implemented) */
void np_choice_noreplace(np_pcg64 *g, int64_t pop, int64_t size, int64_t *out);
/* SeedSequence(seed).generate_state(n, uint32) */
void np_seedsequence_state(uint64_t seed, int n_words32, uint32_t *out);
/* 10-word export used by the device-side generator: state[4] (lo..hi), inc[4], has_uint32, uinteger */
void np_export(const np_pcg64 *g, uint32_t out[10]);
void np_import(np_pcg64 *g, const uint32_t in[10]);

#ifdef __cplusplus
}
#endif
#endif
