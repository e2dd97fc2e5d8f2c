/* np_random.c -- TEST INFRASTRUCTURE (CPU oracle); see np_random.h for provenance. */
#include "np_random.h"
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---- SeedSequence (numpy/random/bit_generator.pyx: SeedSequence.mix_entropy / generate_state) ---- */
#define SS_POOL 4
static const uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u;
static const uint32_t INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
static const uint32_t MIX_MULT_L = 0xca01f9ddu, MIX_MULT_R = 0x4973f715u;
#define SS_XSHIFT 16

static uint32_t ss_hashmix(uint32_t value, uint32_t *hash_const) {
    value ^= *hash_const;
    *hash_const *= MULT_A;
    value *= *hash_const;
    value ^= value >> SS_XSHIFT;
    return value;
}
static uint32_t ss_mix(uint32_t x, uint32_t y) {
    uint32_t r = MIX_MULT_L * x - MIX_MULT_R * y;
    r ^= r >> SS_XSHIFT;
    return r;
}

void np_seedsequence_state(uint64_t seed, int n_words32, uint32_t *out) {
    /* _coerce_to_uint32_array(int): little-endian 32-bit words, [0] for zero */
    uint32_t entropy[2];
    int n_ent = 1;
    entropy[0] = (uint32_t)seed;
    entropy[1] = (uint32_t)(seed >> 32);
    if (entropy[1]) n_ent = 2;
    uint32_t pool[SS_POOL];
    uint32_t hc = INIT_A;
    for (int i = 0; i < SS_POOL; i++) pool[i] = ss_hashmix(i < n_ent ? entropy[i] : 0u, &hc);
    for (int s = 0; s < SS_POOL; s++)
        for (int d = 0; d < SS_POOL; d++)
            if (s != d) pool[d] = ss_mix(pool[d], ss_hashmix(pool[s], &hc));
    /* (entropy longer than the pool would be mixed in here; a uint64 seed never is) */
    uint32_t hb = INIT_B;
    for (int i = 0; i < n_words32; i++) {
        uint32_t v = pool[i % SS_POOL];
        v ^= hb;
        hb *= MULT_B;
        v *= hb;
        v ^= v >> SS_XSHIFT;
        out[i] = v;
    }
}

/* ---- PCG64 (numpy/random/src/pcg64/pcg64.h) ---- */
#define PCG_MULT ((((u128)2549297995355413924ULL) << 64) | (u128)4865540595714422341ULL)

static inline u128 g_state(const np_pcg64 *g) { return (((u128)g->state_hi) << 64) | g->state_lo; }
static inline u128 g_inc(const np_pcg64 *g) { return (((u128)g->inc_hi) << 64) | g->inc_lo; }
static inline void s_state(np_pcg64 *g, u128 s) { g->state_lo = (uint64_t)s; g->state_hi = (uint64_t)(s >> 64); }

static inline void pcg_step(np_pcg64 *g) { s_state(g, g_state(g) * PCG_MULT + g_inc(g)); }

void np_seed(np_pcg64 *g, uint64_t seed) {
    /* PCG64.__init__: val = seed_seq.generate_state(4, uint64); pcg64_set_seed(state=val[0:2], inc=val[2:4])
       with PCG_128BIT_CONSTANT(high=val[k], low=val[k+1]) */
    uint32_t w[8];
    np_seedsequence_state(seed, 8, w);
    uint64_t v[4];
    for (int i = 0; i < 4; i++) v[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
    u128 initstate = (((u128)v[0]) << 64) | v[1];
    u128 initseq = (((u128)v[2]) << 64) | v[3];
    u128 inc = (initseq << 1) | 1u;
    g->inc_lo = (uint64_t)inc;
    g->inc_hi = (uint64_t)(inc >> 64);
    s_state(g, 0);
    pcg_step(g);
    s_state(g, g_state(g) + initstate);
    pcg_step(g);
    g->has_uint32 = 0;
    g->uinteger = 0;
}

uint64_t np_next64(np_pcg64 *g) {
    pcg_step(g);
    uint64_t hi = g->state_hi, lo = g->state_lo;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58); /* state >> 122 */
    return (x >> rot) | (x << ((-rot) & 63));
}

uint32_t np_next32(np_pcg64 *g) {
    if (g->has_uint32) {
        g->has_uint32 = 0;
        return g->uinteger;
    }
    uint64_t n = np_next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(n >> 32);
    return (uint32_t)n;
}

uint32_t np_bounded32(np_pcg64 *g, uint32_t rng) {
    /* distributions.c: random_bounded_uint64 -> (rng <= 0xFFFFFFFF, !use_masked) buffered_bounded_lemire_uint32 */
    if (rng == 0) return 0;
    if (rng == 0xFFFFFFFFu) return np_next32(g);
    const uint32_t rng_excl = rng + 1;
    uint64_t m = (uint64_t)np_next32(g) * rng_excl;
    uint32_t leftover = (uint32_t)m;
    if (leftover < rng_excl) {
        const uint32_t threshold = (0xFFFFFFFFu - rng) % rng_excl;
        while (leftover < threshold) {
            m = (uint64_t)np_next32(g) * rng_excl;
            leftover = (uint32_t)m;
        }
    }
    return (uint32_t)(m >> 32);
}

void np_integers(np_pcg64 *g, uint32_t high_excl, int cnt, int64_t *out) {
    for (int i = 0; i < cnt; i++) out[i] = np_bounded32(g, high_excl - 1);
}

static void shuffle_int(np_pcg64 *g, int64_t n, int64_t first, int64_t *data) {
    /* _generator.pyx:_shuffle_int: for i in reversed(range(first, n)): j = bounded[0,i]; swap */
    for (int64_t i = n - 1; i >= first; i--) {
        int64_t j = np_bounded32(g, (uint32_t)i);
        int64_t t = data[i];
        data[i] = data[j];
        data[j] = t;
    }
}

void np_choice_noreplace(np_pcg64 *g, int64_t pop, int64_t size, int64_t *out) {
    if (size <= 0) return;
    if (pop > 10000 && size > pop / 50) {
        /* tail shuffle */
        int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)pop);
        for (int64_t i = 0; i < pop; i++) idx[i] = i;
        int64_t first = pop - size;
        if (first < 1) first = 1;
        shuffle_int(g, pop, first, idx);
        memcpy(out, idx + (pop - size), sizeof(int64_t) * (size_t)size);
        free(idx);
        return;
    }
    /* Floyd's algorithm; numpy keeps the picked values in an open-addressing hash set, which only
       serves as a membership test, so a scan over the values picked so far is equivalent */
    for (int64_t j = pop - size; j < pop; j++) {
        int64_t k = j - pop + size;
        int64_t val = np_bounded32(g, (uint32_t)j);
        int found = 0;
        for (int64_t q = 0; q < k; q++)
            if (out[q] == val) { found = 1; break; }
        out[k] = found ? j : val;
    }
    shuffle_int(g, size, 1, out);
}

void np_export(const np_pcg64 *g, uint32_t o[10]) {
    o[0] = (uint32_t)g->state_lo; o[1] = (uint32_t)(g->state_lo >> 32);
    o[2] = (uint32_t)g->state_hi; o[3] = (uint32_t)(g->state_hi >> 32);
    o[4] = (uint32_t)g->inc_lo;   o[5] = (uint32_t)(g->inc_lo >> 32);
    o[6] = (uint32_t)g->inc_hi;   o[7] = (uint32_t)(g->inc_hi >> 32);
    o[8] = g->has_uint32; o[9] = g->uinteger;
}
void np_import(np_pcg64 *g, const uint32_t i[10]) {
    g->state_lo = i[0] | ((uint64_t)i[1] << 32); g->state_hi = i[2] | ((uint64_t)i[3] << 32);
    g->inc_lo = i[4] | ((uint64_t)i[5] << 32);   g->inc_hi = i[6] | ((uint64_t)i[7] << 32);
    g->has_uint32 = i[8]; g->uinteger = i[9];
}
